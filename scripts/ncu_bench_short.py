"""Short stand-in for bench.py under ncu's launch-list pass: 3 warm-up + 2 timed single-call commits at 2^20 (the headline
step), one NTT 2^22, one create_witness and one batched opening -- same entry points and sizes as bench.py, without its
2^24 extras (ncu serialises and replays every launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
dev = torch.device("cuda:0")
n = 1 << 20
s = cpu.fr_stream(69, 1)[0]
srs = _ffi.srs_generate(0, s, n)
sc = torch.empty((n, 4), dtype=torch.int64, device=dev)
_ffi.fr_stream_dev(sc.data_ptr(), 42, n)
out = torch.zeros(12, dtype=torch.int64, device=dev)
for _ in range(5):
    srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
torch.cuda.synchronize()
buf = torch.empty((1 << 22, 4), dtype=torch.int64, device=dev)
_ffi.fr_stream_dev(buf.data_ptr(), 5, 1 << 22)
for _ in range(2):
    _ffi.ntt_dev(buf.data_ptr(), 22, False)
torch.cuda.synchronize()
xy = torch.from_numpy(np.vstack([cpu.fr_stream(7, 1)[0], np.zeros(4, dtype=np.uint64)]).view(np.int64)).to(dev)
work = torch.empty((n, 4), dtype=torch.int64, device=dev)
_ffi.witness_coeff_dev(srs, sc.data_ptr(), n, xy.data_ptr(), work.data_ptr(), out.data_ptr(), 0)
torch.cuda.synchronize()
# one batched opening (256 points on the polynomial)
m = 256
xs = torch.empty((m, 4), dtype=torch.int64, device=dev); _ffi.fr_stream_dev(xs.data_ptr(), 7, m)
ys = torch.empty((m, 4), dtype=torch.int64, device=dev)
import ctypes as C
lib = _ffi.lib()
xy2 = torch.zeros((2, 4), dtype=torch.int64, device=dev)
for i in range(m):
    xy2[0] = xs[i]
    lib.kzgb_divide_linear_dev(C.c_void_p(sc.data_ptr()), C.c_size_t(n), C.c_void_p(xy2.data_ptr()), C.c_int(0), C.c_void_p(work.data_ptr()), C.c_void_p(ys[i].data_ptr()), C.c_void_p(0))
r = torch.zeros((m + 2, 4), dtype=torch.int64, device=dev)
print("batched status", _ffi.witness_batched_dev(srs, sc.data_ptr(), n, xs.data_ptr(), ys.data_ptr(), m, r.data_ptr(), out.data_ptr(), 0))
torch.cuda.synchronize()
