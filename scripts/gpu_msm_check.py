"""MSM parity vs the C oracle at small sizes + first timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kzg_b200 import _ffi
from oracle import cpu

_ffi.init(0)
s = cpu.fr_stream(69, 1)[0]
N = 4096
t = time.time(); gs = cpu.setup_g1(s, N); print("cpu setup", N, "%.2fs" % (time.time() - t))
aff = cpu.g1_to_affine(gs)
ok = True
for n, wb, mode in [(1, 0, 0), (2, 0, 0), (33, 0, 0), (1024, 0, 0), (1024, 8, 0), (1024, 13, 0), (4096, 0, 0), (4096, 16, 0), (4096, 0, 1), (4096, 11, 1)]:
    srs = _ffi.SRS.from_jacobian(gs[:n], wb)
    assert (srs.read_affine() == aff[:n]).all(), "to_affine mismatch"
    sc = cpu.fr_stream(42, n, mode=mode)
    got = srs.msm(sc)
    exp = cpu.g1_to_affine(cpu.msm_g1(gs[:n], sc, 8))[0]
    good = bool((got == exp).all())
    # jacobian output path
    gj = cpu.g1_to_affine(srs.msm_jacobian(sc))[0]
    good2 = bool((gj == exp).all())
    # partial length (m < n)
    m = max(1, n // 3)
    got3 = srs.msm(sc[:m]); exp3 = cpu.g1_to_affine(cpu.msm_g1(gs[:m], sc[:m], 8))[0]
    good3 = bool((got3 == exp3).all())
    print("n=%d c=%d mode=%d affine:%s jac:%s prefix:%s" % (n, srs.window_bits, mode, good, good2, good3))
    ok &= good and good2 and good3
# edge: zero scalars, repeated points, scalar = r-1, 1
n = 64
srs = _ffi.SRS.from_jacobian(gs[:n])
z = np.zeros((n, 4), dtype=np.uint64)
print("all-zero scalars -> identity:", bool((srs.msm(z) == 0).all()))
one = cpu.fr_from_canonical(np.array([[1, 0, 0, 0]] * n, dtype=np.uint64))
got = srs.msm(one); exp = cpu.g1_to_affine(cpu.msm_g1(gs[:n], one, 1))[0]
print("all-one scalars:", bool((got == exp).all()))
same = np.repeat(gs[:1], n, axis=0)
srs2 = _ffi.SRS.from_jacobian(same)
sc = cpu.fr_stream(3, n)
got = srs2.msm(sc); exp = cpu.g1_to_affine(cpu.msm_g1(same, sc, 1))[0]
print("identical points:", bool((got == exp).all()))
got = srs2.msm(one); exp = cpu.g1_to_affine(cpu.msm_g1(same, one, 1))[0]
print("identical points, identical scalars (doubling path):", bool((got == exp).all()))
print("ALL OK" if ok else "FAILURES")
