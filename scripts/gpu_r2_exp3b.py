"""Round-2 experiment 3b: NTT plans for 2^16..2^21 (two vs three passes, 512- vs 1024-element tiles, swizzle on/off)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kzg_b200 import _ffi
_ffi.init(0)
dev = torch.device("cuda:0")
def run(log_n, env):
    for k in ("KZGB_NTT_SWZ", "KZGB_NTT_PLAN", "KZGB_NTT_TILE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    n = 1 << log_n
    buf = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(buf.data_ptr(), 5, n)
    _ffi.ntt_dev(buf.data_ptr(), log_n, False)
    torch.cuda.synchronize()
    res = buf.clone()
    for _ in range(3):
        _ffi.ntt_dev(buf.data_ptr(), log_n, False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(30):
        _ffi.ntt_dev(buf.data_ptr(), log_n, False)
    e1.record(); torch.cuda.synchronize()
    return res, e0.elapsed_time(e1) / 30
for log_n in (16, 17, 18, 19, 20, 21):
    base, ms0 = run(log_n, {"KZGB_NTT_SWZ": "0"})
    print(json.dumps({"log_n": log_n, "variant": "current default", "ms": round(ms0, 4)}), flush=True)
    r = log_n - 16
    plans = ["%d,%d" % ((log_n + 1) // 2, log_n // 2), "8,8,%d" % r if r >= 1 else "8,8", "8,%d,%d" % ((log_n - 8 + 1) // 2, (log_n - 8) // 2), "6,6,%d" % (log_n - 12)]
    for plan in plans:
        if any(int(x) > 11 or int(x) < 1 for x in plan.split(",")) or sum(int(x) for x in plan.split(",")) != log_n:
            continue
        for tile in ("9,128", "10,128", "8,128"):
            for swz in ("0", "1"):
                res, ms = run(log_n, {"KZGB_NTT_SWZ": swz, "KZGB_NTT_PLAN": plan, "KZGB_NTT_TILE": tile})
                print(json.dumps({"log_n": log_n, "plan": plan, "tile": tile, "swz": swz, "ms": round(ms, 4), "same": bool(torch.equal(res, base))}), flush=True)
