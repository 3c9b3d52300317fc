"""Device-resident timings of BASELINE configs C2-C4 (scan, witness, eval-form witness, batched witness)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from kzg_b200 import _ffi, scalar
from oracle import cpu
_ffi.init(0)
dev = torch.device("cuda:0"); stream = 0
n = 1 << 20
s = cpu.fr_stream(69, 1)[0]
srs = _ffi.srs_generate(0, s, n)
lag = _ffi.srs_generate(1, s, n)
p = torch.empty((n, 4), dtype=torch.int64, device=dev); _ffi.fr_stream_dev(p.data_ptr(), 42, n)
lib = _ffi.lib()
def timeit(f, reps=10, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3
res = {}
# scan alone
xy = torch.from_numpy(np.vstack([cpu.fr_stream(7, 1)[0], np.zeros(4, dtype=np.uint64)]).view(np.int64)).to(dev)
q = torch.empty((n, 4), dtype=torch.int64, device=dev); rem = torch.zeros(4, dtype=torch.int64, device=dev)
def scan(): _ffi.check(lib.kzgb_divide_linear_dev(C.c_void_p(p.data_ptr()), C.c_size_t(n), C.c_void_p(xy.data_ptr()), C.c_int(0), C.c_void_p(q.data_ptr()), C.c_void_p(rem.data_ptr()), C.c_void_p(stream)))
res["scan_2p20_ms"] = timeit(scan)
scan(); torch.cuda.synchronize(); xy[1] = rem
out = torch.zeros(12, dtype=torch.int64, device=dev)
res["commit_2p20_ms"] = timeit(lambda: srs.msm_dev(p.data_ptr(), n, out.data_ptr(), 0))
res["witness_2p20_ms"] = timeit(lambda: _ffi.witness_coeff_dev(srs, p.data_ptr(), n, xy.data_ptr(), q.data_ptr(), out.data_ptr()))
res["eval_commit_2p20_ms"] = timeit(lambda: lag.msm_dev(p.data_ptr(), n, out.data_ptr(), 0))
res["eval_witness_2p20_ms"] = timeit(lambda: _ffi.witness_eval_dev(lag, p.data_ptr(), 20, 424242, q.data_ptr(), out.data_ptr()))
res["intt_2p20_ms"] = timeit(lambda: _ffi.ntt_dev(q.data_ptr(), 20, True))
# batched witness m = 256
m = 256
xs_h = cpu.fr_stream(7, m)
ys_h = np.vstack([_ffi.divide_linear(p.cpu().numpy().view(np.uint64), xs_h[i], None)[1] for i in range(m)])
xs = torch.from_numpy(xs_h.view(np.int64)).to(dev); ys = torch.from_numpy(ys_h.view(np.int64)).to(dev)
r = torch.zeros((m + 2, 4), dtype=torch.int64, device=dev)
st = _ffi.witness_batched_dev(srs, p.data_ptr(), n, xs.data_ptr(), ys.data_ptr(), m, r.data_ptr(), out.data_ptr())
res["batched_status"] = st
res["witness_batched_2p20_m256_ms"] = timeit(lambda: _ffi.witness_batched_dev(srs, p.data_ptr(), n, xs.data_ptr(), ys.data_ptr(), m, r.data_ptr(), out.data_ptr()), reps=5)
_ffi.profile_enable(True); _ffi.profile_read()
_ffi.witness_batched_dev(srs, p.data_ptr(), n, xs.data_ptr(), ys.data_ptr(), m, r.data_ptr(), out.data_ptr())
pr = _ffi.profile_read(); _ffi.profile_enable(False)
res["batched_stages_ms"] = {k: round(v[0], 3) for k, v in pr.items()}
print(json.dumps(res, indent=1))
