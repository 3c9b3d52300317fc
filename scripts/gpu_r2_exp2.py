"""Round-2 experiment 2: MSM tail knobs at n = 2^20 / 2^21 -- fix-up level-0 slots per thread (KZGB_MSM_FIX_F) and the
accumulate chunk length (KZGB_MSM_CHUNK); per-stage CUDA-event timings, parity with p(s)*G for every setting."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kzg_b200 import _ffi
from oracle import cpu

cpu.build()
_ffi.init(0)
dev = torch.device("cuda:0")
s = cpu.fr_stream(69, 1)[0]
for logn in (20, 21):
    n = 1 << logn
    srs = _ffi.srs_generate(0, s, n, window_bits=20)
    sc = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(sc.data_ptr(), 42, n)
    torch.cuda.synchronize()
    expect = cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), cpu.poly_eval(sc.cpu().numpy().view(np.uint64), s)))[0]
    out = torch.zeros(12, dtype=torch.int64, device=dev)
    for chunk in ("48", "60", "80"):
        for fixf in ("", "16", "8", "12", "20"):
            os.environ["KZGB_MSM_CHUNK"] = chunk
            if fixf:
                os.environ["KZGB_MSM_FIX_F"] = fixf
            else:
                os.environ.pop("KZGB_MSM_FIX_F", None)
            for _ in range(3):
                srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
            torch.cuda.synchronize()
            ok = bool((out.cpu().numpy().view(np.uint64) == expect).all())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            _ffi.profile_enable(True); _ffi.profile_read()
            for _ in range(5):
                srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
            prof = _ffi.profile_read(); _ffi.profile_enable(False)
            print(json.dumps({"log_n": logn, "chunk": chunk, "fix_f": fixf or "auto", "ms": round(ms, 4), "parity": ok,
                              "stage_ms": {k.replace("msm_", ""): round(v[0] / v[1], 3) for k, v in prof.items()}}), flush=True)
    # reference-style u64 scalars (one giant carry bucket): the fix-up must stay logarithmic
    os.environ["KZGB_MSM_CHUNK"] = "48"; os.environ.pop("KZGB_MSM_FIX_F", None)
    sc2 = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(sc2.data_ptr(), 42, n, mode=1)
    for _ in range(2):
        srs.msm_dev(sc2.data_ptr(), n, out.data_ptr(), 0)
    torch.cuda.synchronize()
    exp2 = cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), cpu.poly_eval(sc2.cpu().numpy().view(np.uint64), s)))[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        srs.msm_dev(sc2.data_ptr(), n, out.data_ptr(), 0)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"log_n": logn, "u64_scalars_ms": round(e0.elapsed_time(e1) / 5, 4),
                      "parity": bool((out.cpu().numpy().view(np.uint64) == exp2).all())}), flush=True)
    srs.close()
