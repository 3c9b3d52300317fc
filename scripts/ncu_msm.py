"""Short MSM workload for ncu captures (2 warm-up MSMs + a few profiled ones, n = 2^20)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kzg_b200 import _ffi
from oracle import cpu
_ffi.init(0)
n = 1 << 20
srs = _ffi.srs_generate(0, cpu.fr_stream(69, 1)[0], n)
sc = torch.empty((n, 4), dtype=torch.int64, device="cuda:0")
_ffi.fr_stream_dev(sc.data_ptr(), 42, n)
out = torch.zeros(12, dtype=torch.int64, device="cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
torch.cuda.synchronize()
if len(sys.argv) > 2:   # also an NTT 2^22 and a witness for the launch list
    buf = torch.empty((1 << 22, 4), dtype=torch.int64, device="cuda:0")
    _ffi.fr_stream_dev(buf.data_ptr(), 5, 1 << 22)
    for _ in range(2):
        _ffi.ntt_dev(buf.data_ptr(), 22, False)
    torch.cuda.synchronize()
