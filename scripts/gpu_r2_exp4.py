"""Round-2 experiment 4: folded inter-pass twiddles (KZGB_NTT_FOLD=1, default) vs the round-1 scheme (=0):
bit-equality of forward and inverse transforms at 2^11..2^24 and timings."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kzg_b200 import _ffi
_ffi.init(0)
dev = torch.device("cuda:0")
def run(log_n, fold, inverse):
    os.environ["KZGB_NTT_FOLD"] = fold
    n = 1 << log_n
    buf = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(buf.data_ptr(), 5, n)
    _ffi.ntt_dev(buf.data_ptr(), log_n, inverse)
    torch.cuda.synchronize()
    res = buf.clone()
    for _ in range(3):
        _ffi.ntt_dev(buf.data_ptr(), log_n, inverse)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(30):
        _ffi.ntt_dev(buf.data_ptr(), log_n, inverse)
    e1.record(); torch.cuda.synchronize()
    return res, e0.elapsed_time(e1) / 30
for log_n in (11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24):
    out = {"log_n": log_n}
    for inverse in (False, True):
        a, ta = run(log_n, "0", inverse)
        b, tb = run(log_n, "1", inverse)
        out["inv" if inverse else "fwd"] = {"unfolded_ms": round(ta, 4), "folded_ms": round(tb, 4), "equal": bool(torch.equal(a, b))}
    print(json.dumps(out), flush=True)
