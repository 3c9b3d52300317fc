import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
dev = torch.device("cuda:0")
s = cpu.fr_stream(69, 1)[0]
log_n = 21
n = 1 << log_n
srs = _ffi.srs_generate(0, s, n)
sc = torch.empty((n, 4), dtype=torch.int64, device=dev)
_ffi.fr_stream_dev(sc.data_ptr(), 42, n)
out = torch.zeros(12, dtype=torch.int64, device=dev)
for chunk in (96, 112, 128, 160, 192):
    for fix in (8, 12):
        os.environ["KZGB_MSM_CHUNK"] = str(chunk); os.environ["KZGB_MSM_FIX_F"] = str(fix)
        for _ in range(2): srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"log_n": log_n, "chunk": chunk, "fix_f": fix, "ms": round(e0.elapsed_time(e1) / 10, 3)}), flush=True)
