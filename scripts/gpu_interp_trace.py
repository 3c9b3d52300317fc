"""Per-launch timeline of the Lagrange interpolation behind create_witness_batched (kzgb_interpolate, m points)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
for m in (256, 1024, 4096):
    xs, ys = cpu.fr_stream(7, m), cpu.fr_stream(8, m)
    for _ in range(3): _ffi.interpolate(xs, ys)
    acc = {}
    for _ in range(3):
        _ffi.trace_enable(True); _ffi.interpolate(xs, ys); rec = _ffi.trace_read(); _ffi.trace_enable(False)
        for i, (ln, ms) in enumerate(rec): acc.setdefault(i, [ln, 0.0])[1] += ms / 3
    print("m =", m, " ".join("L%d:%.1fus" % (acc[i][0], acc[i][1] * 1e3) for i in sorted(acc)), "total %.1f us" % (1e3 * sum(v[1] for v in acc.values())))
