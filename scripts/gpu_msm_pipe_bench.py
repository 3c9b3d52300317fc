"""Pipelined batch of 2^20 commits vs the same commits one call at a time (device-resident and host-pointer paths)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
LOGN = int(sys.argv[1]) if len(sys.argv) > 1 else 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = 1 << LOGN
s = cpu.fr_stream(69, 1)[0]
dev = torch.device("cuda:0")
srs = _ffi.srs_generate(0, s, n)
NB = 8
polys = []
for b in range(NB):
    t = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(t.data_ptr(), 42 + b, n)
    polys.append(t)
st = torch.cuda.current_stream().cuda_stream
out = torch.zeros((K, 12), dtype=torch.int64, device=dev)
ref = torch.zeros((K, 12), dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def seq():
    for i in range(K): srs.msm_dev(polys[i % NB].data_ptr(), n, ref[i].data_ptr(), 0, st)
def bat():
    srs.msm_batch_dev([polys[i % NB].data_ptr() for i in range(K)], [n] * K, out.data_ptr(), 0, st)
for f in (seq, bat, seq, bat):
    f(); torch.cuda.synchronize()
print("batch == sequential:", bool(torch.equal(out, ref)))
for name, f in (("sequential msm_dev", seq), ("pipelined batch_dev", bat), ("sequential msm_dev", seq), ("pipelined batch_dev", bat)):
    torch.cuda.synchronize(); e0.record(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("%-22s %.3f ms per commit  (%.3e points/s)" % (name, ms, n / ms * 1e3))
# host-pointer path
hp = [polys[b].cpu().pin_memory() for b in range(2)]
harr = [h.numpy().view(np.uint64) for h in hp]
for name in ("host sequential kzgb_msm_g1", "host pipelined kzgb_msm_g1_batch") * 2:
    t0 = time.perf_counter()
    if "sequential" in name:
        r = np.stack([srs.msm(harr[i % 2]) for i in range(K)])
    else:
        r2 = srs.msm_batch([harr[i % 2] for i in range(K)])
    dt = (time.perf_counter() - t0) / K
    print("%-34s %.3f ms per commit  (%.3e points/s)" % (name, dt * 1e3, n / dt))
print("host batch == host sequential:", bool((r == r2).all()))
# multi-opening: witnesses of one polynomial at K points, pipelined vs one call at a time
xs = cpu.fr_stream(7, K)
p_host = harr[0]
ys = np.vstack([cpu.poly_eval(p_host, xs[i]) for i in range(4)] + [np.zeros((K - 4, 4), dtype=np.uint64)])
for name in ("witness sequential kzgb_witness_coeff", "witness pipelined kzgb_witness_coeff_many") * 2:
    t0 = time.perf_counter()
    if "sequential" in name:
        rs = [_ffi.witness_coeff(srs, p_host, xs[i], ys[i]) for i in range(K)]
    else:
        st_many, w_many = _ffi.witness_coeff_many(srs, p_host, xs, ys)
    dt = (time.perf_counter() - t0) / K
    print("%-44s %.3f ms per witness" % (name, dt * 1e3))
print("witness many == sequential:", all(int(st_many[i]) == rs[i][0] and (rs[i][0] != 0 or (w_many[i] == rs[i][1]).all()) for i in range(K)),
      "statuses", [int(v) for v in st_many[:6]])
