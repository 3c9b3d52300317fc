"""Round-2 experiment 5: warp-owned tile columns in the NTT stages (KZGB_NTT_WCOL=1: __syncwarp between stages) vs the
default block-synchronous stages: bit-equality and timings at 2^14..2^24."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kzg_b200 import _ffi
_ffi.init(0)
dev = torch.device("cuda:0")
def run(log_n, env, inverse=False):
    for k in ("KZGB_NTT_WCOL", "KZGB_NTT_TILE", "KZGB_NTT_PLAN"):
        os.environ.pop(k, None)
    os.environ.update(env)
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
for log_n in (14, 16, 18, 20, 21, 22, 23, 24):
    base, t0 = run(log_n, {})
    out = {"log_n": log_n, "default_ms": round(t0, 4)}
    for name, env in (("wcol", {"KZGB_NTT_WCOL": "1"}), ("wcol_tile10", {"KZGB_NTT_WCOL": "1", "KZGB_NTT_TILE": "10,128"}),
                      ("wcol_tile11_256", {"KZGB_NTT_WCOL": "1", "KZGB_NTT_TILE": "11,256"}), ("wcol_tile10_64", {"KZGB_NTT_WCOL": "1", "KZGB_NTT_TILE": "10,64"})):
        r, t = run(log_n, env)
        out[name] = {"ms": round(t, 4), "equal": bool(torch.equal(r, base))}
    bi, _ = run(log_n, {}, True); ri, ti = run(log_n, {"KZGB_NTT_WCOL": "1"}, True)
    out["inverse_wcol_equal"] = bool(torch.equal(bi, ri))
    print(json.dumps(out), flush=True)
