"""Warm in-stream per-kernel timeline of one 2^20 MSM (kzgb_trace_*): where the tail's time goes."""
import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
LOGN = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << LOGN
s = cpu.fr_stream(69, 1)[0]
dev = torch.device("cuda:0")
sc = torch.empty((n, 4), dtype=torch.int64, device=dev)
_ffi.fr_stream_dev(sc.data_ptr(), 42, n)
out = torch.zeros(12, dtype=torch.int64, device=dev)
srs = _ffi.srs_generate(0, s, n)
for _ in range(4): srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
torch.cuda.synchronize()
src = open(os.path.join(os.path.dirname(_ffi.LIB_PATH), "csrc", "msm.cu")).read().split("\n")
acc = {}
REP = 5
for rep in range(REP):
    _ffi.trace_enable(True)
    srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
    rec = _ffi.trace_read()
    _ffi.trace_enable(False)
    for i, (ln, ms) in enumerate(rec):
        acc.setdefault(i, [ln, 0.0])[1] += ms / REP
tot = 0
for i in sorted(acc):
    ln, ms = acc[i]
    name = ""
    for back in range(1, 6):
        m = re.search(r"(msm_\w+)", src[ln - 1 - back]) if ln - 1 - back >= 0 else None
        if m: name = m.group(1); break
    tot += ms
    print("%3d line %4d %-34s %8.1f us" % (i, ln, name, ms * 1e3))
print("total after first mark: %.3f ms" % tot)
