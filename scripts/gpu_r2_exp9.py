"""Round-2 experiment 9: LATENCY of one dependent Fq product for a lone warp per scheduler (what the MSM's latency-bound
tail kernels run at), schoolbook interleaved Montgomery (mode 4) vs Karatsuba + wide reduction (mode 13) vs the dedicated
square (mode 16): ns per dependent product = resident threads / (products per second)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_b200 import _ffi
_ffi.init(0)
SM = 148
for blocks, threads in ((1, 32), (1, 128), (2, 128), (3, 128), (4, 256)):
    row = {"blocks_per_sm": blocks, "threads": threads, "warps_per_scheduler": blocks * threads / 128}
    for name, mode in (("schoolbook", 4), ("karatsuba", 13), ("square", 16), ("fr_schoolbook", 5), ("fr_karatsuba", 14)):
        r = max((_ffi.microbench(mode, blocks, threads, 64) for _ in range(3)), key=lambda x: x["ops_per_s"])
        row[name + "_ns_per_product"] = round(SM * blocks * threads / r["ops_per_s"] * 1e9, 1)
    print(json.dumps(row), flush=True)
