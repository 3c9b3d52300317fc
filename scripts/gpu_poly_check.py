"""NTT / scan / eval-form parity vs the C oracle + timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kzg_b200 import _ffi
from oracle import cpu
from oracle import kzg_oracle as O

_ffi.init(0)
ok = True
# NTT parity at many sizes
for lg in list(range(0, 15)) + [16, 20, 21]:
    a = cpu.fr_stream(5, 1 << lg)
    st, f = _ffi.ntt(a, lg, False)
    ref = cpu.fft(a, lg, False, 8)
    good = bool((f == ref).all())
    st, b = _ffi.ntt(f, lg, True)
    good2 = bool((b == a).all())
    ref_i = cpu.fft(a, lg, True, 8)
    st, fi = _ffi.ntt(a, lg, True)
    good3 = bool((fi == ref_i).all())
    print("ntt 2^%d fwd:%s roundtrip:%s inv:%s" % (lg, good, good2, good3)); ok &= good and good2 and good3
# scan parity
x = cpu.fr_stream(7, 1)[0]
for n in [1, 2, 3, 16, 17, 4095, 4096, 4097, 65536 + 5, 1 << 20]:
    p = cpu.fr_stream(42, n)
    y = cpu.poly_eval(p, x)
    q, rem = _ffi.divide_linear(p, x, y)
    good = bool((rem == 0).all())
    if 1 < n <= 70000:
        # oracle long division of (p - y) by (X - x)  (polynomial.rs:193-227 restated in C)
        xi = O.fr_from_mont_limbs(x); yi = O.fr_from_mont_limbs(y)
        pp = p.copy(); pp[0] = np.array(O.fr_to_mont_limbs((O.fr_from_mont_limbs(p[0]) - yi) % O.R_MOD), dtype=np.uint64)
        div = np.array([O.fr_to_mont_limbs((-xi) % O.R_MOD), O.fr_to_mont_limbs(1)], dtype=np.uint64)
        st, qq, rr = cpu.long_division(pp, n - 1, div, 1)
        good &= (st == 0) and bool((qq == q).all())
    q2, rem2 = _ffi.divide_linear(p, x, None)
    good &= bool((rem2 == y).all())
    print("scan n=%d rem0:%s" % (n, good)); ok &= good
print("ALL OK" if ok else "FAILURES")
# NTT timing on device
dev = torch.device("cuda:0")
for lg in (20, 22, 24):
    n = 1 << lg
    buf = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(buf.data_ptr(), 1, n)
    for _ in range(3): _ffi.ntt_dev(buf.data_ptr(), lg, False)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    _ffi.profile_enable(True); _ffi.profile_read()
    for _ in range(10): _ffi.ntt_dev(buf.data_ptr(), lg, False)
    pr = _ffi.profile_read(); _ffi.profile_enable(False)
    ms = pr["ntt"][0] / pr["ntt"][1]
    print("ntt 2^%d: %.3f ms  (%.1f GB/s algorithmic 2-pass bytes)" % (lg, ms, 4 * 32 * n / ms / 1e6))
