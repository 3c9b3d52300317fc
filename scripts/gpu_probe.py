"""First-contact GPU probe: field parity vs the C oracle + instruction-rate microbenchmarks."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kzg_b200 import _ffi
from oracle import cpu

_ffi.init(0)
n = 4096
a = cpu.fr_stream(1, n); b = cpu.fr_stream(2, n)
ok = {}
ok["fr_mul"] = bool((_ffi.test_fr_op(a, b, 0) == cpu.fr_mul_vec(a, b)).all())
# Fp inputs: take products of stream values reduced by the oracle: use fp_mul_vec of arbitrary < p inputs
rng = np.random.default_rng(5)
fa = rng.integers(0, 2**63, size=(n, 6), dtype=np.uint64); fa[:, 5] &= np.uint64(0x0fffffffffffffff)
fb = rng.integers(0, 2**63, size=(n, 6), dtype=np.uint64); fb[:, 5] &= np.uint64(0x0fffffffffffffff)
ok["fp_mul"] = bool((_ffi.test_fp_op(fa, fb, 0) == cpu.fp_mul_vec(fa, fb)).all())
inv = _ffi.test_fr_op(a[:64], None, 3)
one = cpu.fr_mul_vec(inv, a[:64])
ok["fr_inv"] = bool((one == one[0]).all())
print("PARITY", ok)
res = {}
for mode, name in [(0, "imad_lo"), (1, "imad_hi"), (2, "imad_wide_mix"), (3, "imad_wide"), (6, "iadd"), (7, "wide_plus_iadd"), (4, "fq_mul"), (5, "fr_mul")]:
    for bps, thr in [(2, 256), (4, 256), (8, 256)]:
        it = 64 if mode < 4 or mode > 5 else 8
        r = _ffi.microbench(mode, bps, thr, it)
        res[f"{name}_b{bps}"] = r
        print(name, bps, thr, "%.3e ops/s  %.2f ops/clk/SM  %.3f ms" % (r["ops_per_s"], r["ops_per_clk_sm"], r["ms"]))
r = _ffi.microbench(100, 0, 0, 1024)
print("copy 1GiB: %.1f GB/s" % (r["ops_per_s"] / 1e9))
res["copy"] = r
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"parity": ok, "micro": res}, open("gpurun_out/probe.json", "w"), indent=1)
