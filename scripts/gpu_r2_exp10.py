"""Round-2 experiment 10: pointwise_div_kernel (coset division of the batched opening, 2^20 elements) with the
block-level batched inversion, per build variant (KZGB_LIB): per-thread batch x threads per block x min blocks per SM.
Prints the kernel's in-stream time (kzgb_trace mark of its launch), the whole batched witness, and the witness bytes."""
import hashlib, json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
n, m = 1 << 20, 256
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
def call():
    return _ffi.witness_batched_dev(srs, p.data_ptr(), n, xs.data_ptr(), ys.data_ptr(), m, r.data_ptr(), w.data_ptr(), 0)
for _ in range(3): st = call()
torch.cuda.synchronize()
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kzg_b200", "csrc", "batched.cu")).read().split("\n")
pd_line = next(i for i, l in enumerate(src) if "pointwise_div_kernel<<<" in l) + 1
acc = []
for _ in range(5):
    _ffi.trace_enable(True); call(); rec = _ffi.trace_read(); _ffi.trace_enable(False)
    acc.append([ms for ln, ms in rec if abs(ln - pd_line) <= 2 and ms > 0.03][0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): call()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(_ffi.LIB_PATH), "status": st, "pointwise_div_us": round(1e3 * sum(acc) / len(acc), 1),
                  "witness_batched_ms": round(e0.elapsed_time(e1) / 10, 3),
                  "digest": hashlib.sha256(w.cpu().numpy().tobytes() + r.cpu().numpy().tobytes()).hexdigest()[:16]}), flush=True)
