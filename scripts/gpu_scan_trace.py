"""Per-launch timeline of one linear quotient (kzgb_divide_linear_dev) at 2^20 and 2^24, warm, via kzgb_trace_*."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
lib = _ffi.lib()
dev = torch.device("cuda:0")
for log_n in (20, 24):
    n = 1 << log_n
    p = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(p.data_ptr(), 42, n)
    q = torch.empty((n, 4), dtype=torch.int64, device=dev)
    rem = torch.zeros(4, dtype=torch.int64, device=dev)
    xy = torch.from_numpy(np.vstack([cpu.fr_stream(7, 1)[0], np.zeros(4, dtype=np.uint64)]).view(np.int64)).to(dev)
    def call():
        _ffi.check(lib.kzgb_divide_linear_dev(C.c_void_p(p.data_ptr()), C.c_size_t(n), C.c_void_p(xy.data_ptr()), C.c_int(0),
                                              C.c_void_p(q.data_ptr()), C.c_void_p(rem.data_ptr()), C.c_void_p(0)))
    for _ in range(5): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): call()
    e1.record(); torch.cuda.synchronize()
    import hashlib
    dig = hashlib.sha256(q[: n - 1].cpu().numpy().tobytes() + rem.cpu().numpy().tobytes()).hexdigest()[:16]
    print("%s 2^%d: %.1f us per division, digest %s" % (os.path.basename(_ffi.LIB_PATH), log_n, e0.elapsed_time(e1) / 50 * 1e3, dig))
    acc = {}
    for rep in range(5):
        _ffi.trace_enable(True); call(); rec = _ffi.trace_read(); _ffi.trace_enable(False)
        for i, (ln, ms) in enumerate(rec): acc.setdefault(i, [ln, 0.0])[1] += ms / 5
    for i in sorted(acc): print("   mark %d (scan.cu:%d) +%.1f us" % (i, acc[i][0], acc[i][1] * 1e3))
