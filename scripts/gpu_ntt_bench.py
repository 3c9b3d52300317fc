import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kzg_b200 import _ffi
from oracle import cpu
_ffi.init(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << lg
dev = torch.device("cuda:0")
buf = torch.empty((n, 4), dtype=torch.int64, device=dev)
_ffi.fr_stream_dev(buf.data_ptr(), 5, n)
ref = None
if lg <= 22:
    a = cpu.fr_stream(5, n)
    st, f = _ffi.ntt(a, lg, False)
    k = 12345 % n
    from kzg_b200 import scalar
    from oracle import kzg_oracle as O
    omega = O.compute_omega(n)[2]
    ok = bool((f[k] == cpu.poly_eval(a, scalar.fr(pow(omega, k, O.R_MOD)))).all()) and bool((_ffi.ntt(f, lg, True)[1] == a).all())
else:
    ok = None
for _ in range(3): _ffi.ntt_dev(buf.data_ptr(), lg, False)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): _ffi.ntt_dev(buf.data_ptr(), lg, False)
e1.record(); torch.cuda.synchronize()
print("plan=%s ntt 2^%d: %.3f ms parity=%s" % (os.environ.get("KZGB_NTT_PLAN", "default"), lg, e0.elapsed_time(e1) / 20, ok))
