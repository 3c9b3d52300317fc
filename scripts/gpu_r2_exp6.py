"""Round-2 experiment 6: the host-pointer commit (kzgb_msm_g1) from PAGEABLE caller memory (a Rust Vec<Fr> is pageable)
against pinned memory, with upload_async's threaded pinned relay on (KZGB_UPLOAD_THREADS = 1..8) and off (= 0: the
driver's own pageable path).  Run once per setting: the knob is read once per process."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
n = 1 << 20
s = cpu.fr_stream(69, 1)[0]
srs = _ffi.srs_generate(0, s, n)
bufs = [cpu.fr_stream(42 + i, n) for i in range(4)]                       # pageable numpy
pins = [torch.from_numpy(b.view(np.int64)).pin_memory() for b in bufs]
pin_views = [p.numpy().view(np.uint64) for p in pins]
def timed(views, reps=20):
    for i in range(3):
        srs.msm(views[i % 4])
    t0 = time.perf_counter()
    for i in range(reps):
        srs.msm(views[i % 4])
    return (time.perf_counter() - t0) / reps * 1e3
same = all((srs.msm(bufs[i]) == srs.msm(pin_views[i])).all() for i in range(4))
print(json.dumps({"upload_threads": os.environ.get("KZGB_UPLOAD_THREADS", "default(4)"), "n": n, "results_equal": bool(same),
                  "pageable_ms": round(timed(bufs), 3), "pinned_ms": round(timed(pin_views), 3),
                  "pageable_ms_again": round(timed(bufs), 3)}), flush=True)
