"""Warm per-launch timeline of create_witness_batched (2^20 coefficients, 256 openings), device resident."""
import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
n, m = 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
srs = _ffi.srs_generate(0, cpu.fr_stream(69, 1)[0], n)
p = torch.empty((n, 4), dtype=torch.int64, device=dev); _ffi.fr_stream_dev(p.data_ptr(), 42, n)
xs = torch.empty((m, 4), dtype=torch.int64, device=dev); _ffi.fr_stream_dev(xs.data_ptr(), 7, m)
ys = torch.empty((m, 4), dtype=torch.int64, device=dev)
work = torch.empty((n, 4), dtype=torch.int64, device=dev)
xy = torch.zeros((2, 4), dtype=torch.int64, device=dev)
lib = _ffi.lib()
for i in range(m):
    xy[0] = xs[i]
    lib.kzgb_divide_linear_dev(C.c_void_p(p.data_ptr()), C.c_size_t(n), C.c_void_p(xy.data_ptr()), C.c_int(0), C.c_void_p(work.data_ptr()), C.c_void_p(ys[i].data_ptr()), C.c_void_p(0))
r = torch.zeros((m + 2, 4), dtype=torch.int64, device=dev)
w = torch.zeros(12, dtype=torch.int64, device=dev)
for _ in range(3):
    st = _ffi.witness_batched_dev(srs, p.data_ptr(), n, xs.data_ptr(), ys.data_ptr(), m, r.data_ptr(), w.data_ptr(), 0)
torch.cuda.synchronize()
_ffi.trace_enable(True)
st = _ffi.witness_batched_dev(srs, p.data_ptr(), n, xs.data_ptr(), ys.data_ptr(), m, r.data_ptr(), w.data_ptr(), 0)
rec = _ffi.trace_read()
_ffi.trace_enable(False)
print("status", st, "launches", len(rec))
tot = 0; msm = 0
for i, (ln, ms) in enumerate(rec):
    tot += ms
    print("%3d line %4d %8.1f us" % (i, ln, ms * 1e3))
print("total %.3f ms" % tot)
