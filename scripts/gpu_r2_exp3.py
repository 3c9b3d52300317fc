"""Round-2 experiment 3: NTT shared-memory swizzle (KZGB_NTT_SWZ), pass plans (KZGB_NTT_PLAN) and tile shapes
(KZGB_NTT_TILE) at 2^20 / 2^22 / 2^24; every variant is compared bit for bit with the unswizzled default plan."""
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
    src = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(src.data_ptr(), 5, n)
    buf = src.clone()
    _ffi.ntt_dev(buf.data_ptr(), log_n, False)
    torch.cuda.synchronize()
    res = buf.clone()
    inv = buf.clone()
    _ffi.ntt_dev(inv.data_ptr(), log_n, True)
    torch.cuda.synchronize()
    rt = bool(torch.equal(inv, src))
    for _ in range(3):
        _ffi.ntt_dev(buf.data_ptr(), log_n, False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        _ffi.ntt_dev(buf.data_ptr(), log_n, False)
    e1.record(); torch.cuda.synchronize()
    return res, e0.elapsed_time(e1) / 20, rt


for log_n, plans in ((20, ["", "10,10", "7,7,6", "8,6,6", "8,8,4"]), (22, ["", "11,11", "8,8,6", "8,7,7", "10,6,6"]), (24, ["", "8,8,8", "10,8,6", "9,8,7"])):
    base, ms0, rt0 = run(log_n, {"KZGB_NTT_SWZ": "0"})
    print(json.dumps({"log_n": log_n, "variant": "round-1 layout (no swizzle), default plan", "ms": round(ms0, 4), "roundtrip": rt0}), flush=True)
    for plan in plans:
        for tile in ("", "11,256", "10,256", "9,128"):
            env = {"KZGB_NTT_SWZ": "1"}
            if plan:
                env["KZGB_NTT_PLAN"] = plan
            if tile:
                env["KZGB_NTT_TILE"] = tile
            try:
                res, ms, rt = run(log_n, env)
                same = bool(torch.equal(res, base))
            except Exception as ex:   # noqa: BLE001
                print(json.dumps({"log_n": log_n, "plan": plan or "default", "tile": tile or "default", "error": repr(ex)[:120]}), flush=True)
                continue
            print(json.dumps({"log_n": log_n, "plan": plan or "default", "tile": tile or "default (10,128)", "swizzle": 1, "ms": round(ms, 4),
                              "equals_round1_layout": same, "roundtrip": rt}), flush=True)
