"""Round-2 experiment 8: occupancy of msm_accumulate_kernel -- more resident warps at a lower register cap
(KZG_MSM_ACC_THREADS x KZG_MSM_ACC_BLOCKS at __maxnreg__ KZG_MSM_ACC_MAXREG; libraries under build_var2/, selected with
KZGB_LIB) against the shipped 128 threads x 3 blocks at 168 registers.  Prints the commit time, the accumulate stage time
and the result's bytes (every variant must print the same digest)."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
for _ in range(4):
    srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
e1.record(); torch.cuda.synchronize()
_ffi.trace_enable(True)
srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
rec = _ffi.trace_read()
_ffi.trace_enable(False)
print(json.dumps({"lib": os.path.basename(_ffi.LIB_PATH), "commit_ms": round(e0.elapsed_time(e1) / 20, 4),
                  "largest_stage_ms": round(max(ms for _, ms in rec), 4),
                  "digest": hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]}), flush=True)
