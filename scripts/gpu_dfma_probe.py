"""B200 probe: FP64-pipe (DFMA) field product vs the IMAD product -- bit-exactness and chained throughput."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kzg_b200 import _ffi
from oracle import cpu as ocpu, kzg_oracle as O

cpu = ocpu; cpu.build()
_ffi.init(0)
n = 1 << 16
a, b = cpu.fr_stream(11, n), cpu.fr_stream(12, n)
ok_fr = bool((_ffi.test_fr_op(a, b, 8) == cpu.fr_mul_vec(a, b)).all())
rng = np.random.default_rng(7)
fa = rng.integers(0, 2 ** 63, size=(n, 6), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 6), dtype=np.uint64)
fb = rng.integers(0, 2 ** 63, size=(n, 6), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 6), dtype=np.uint64)
fa[:, 5] &= np.uint64(0x0fffffffffffffff)
fb[:, 5] &= np.uint64(0x0fffffffffffffff)
pm1 = np.array(O.fq_to_mont_limbs(O.P_MOD - 1), dtype=np.uint64)
fa[0], fb[0], fa[1], fb[1] = pm1, pm1, 0, pm1
fa[2] = np.uint64(0xffffffffffffffff); fa[2, 5] = np.uint64(0x0fffffffffffffff); fb[2] = fa[2]
ok_fq = bool((_ffi.test_fp_op(fa, fb, 8) == cpu.fp_mul_vec(fa, fb)).all())
print(json.dumps({"mul_dfma_bit_exact_fr": ok_fr, "mul_dfma_bit_exact_fq": ok_fq}))
names = {2: "IMAD.WIDE", 3: "mad.lo.cc+madc.hi", 4: "Fq mul (IMAD)", 5: "Fr mul (IMAD)", 6: "IADD", 8: "DFMA",
         9: "DFMA with IMAD.WIDE 1:1", 10: "Fq mul (DFMA product)", 11: "Fr mul (DFMA product)",
         12: "Fq: IMAD chain + DFMA chain per thread"}
for mode, nm in names.items():
    for bps, thr in ((4, 256), (8, 128), (2, 256), (3, 256)) if mode >= 10 or mode in (4, 5) else ((8, 256),):
        r = max((_ffi.microbench(mode, bps, thr, 32) for _ in range(3)), key=lambda x: x["ops_per_s"])
        print(json.dumps({"mode": mode, "name": nm, "blocks_per_sm": bps, "threads": thr, **r}))
