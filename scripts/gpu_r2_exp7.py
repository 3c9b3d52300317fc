"""Round-2 experiment 7: accumulate chunk length (KZGB_MSM_CHUNK) and fix-up fan-in (KZGB_MSM_FIX_F) at the two large
shapes: one GPU's 2^24-point commit (c = 22, 2.0e8 sorted entries) and an 8-GPU shard of it (2^21 points, c = 20)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
dev = torch.device("cuda:0")
s = cpu.fr_stream(69, 1)[0]
for log_n, chunks in ((21, (40, 48, 64, 80, 96)), (24, (64, 80, 96, 128, 160, 192))):
    n = 1 << log_n
    srs = _ffi.srs_generate(0, s, n)
    sc = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(sc.data_ptr(), 42, n)
    out = torch.zeros(12, dtype=torch.int64, device=dev)
    ref = None
    for chunk in chunks:
        for fix in ((8,) if log_n == 21 else (8, 12, 16)):
            os.environ["KZGB_MSM_CHUNK"] = str(chunk); os.environ["KZGB_MSM_FIX_F"] = str(fix)
            for _ in range(2):
                srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
            torch.cuda.synchronize()
            res = out.clone()
            if ref is None: ref = res
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10 if log_n == 21 else 4
            e0.record()
            for _ in range(reps):
                srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
            e1.record(); torch.cuda.synchronize()
            print(json.dumps({"log_n": log_n, "window_bits": srs.window_bits, "chunk": chunk, "fix_f": fix,
                              "ms": round(e0.elapsed_time(e1) / reps, 3), "equal": bool(torch.equal(res, ref))}), flush=True)
    del srs, sc
    torch.cuda.empty_cache()
