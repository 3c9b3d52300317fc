"""Round-2 experiment 11: div_by_omega_kernel (eval-form quotient, d = 2^20) with the block-level batched inversion, per
build variant (KZGB_LIB): per-thread batch x min blocks per SM.  In-stream kernel time via kzgb_trace, whole eval-form
create_witness, and the witness bytes."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
log_d = 20; n = 1 << log_d
dev = torch.device("cuda:0")
srs = _ffi.srs_generate(0, cpu.fr_stream(69, 1)[0], n)
f = torch.empty((n, 4), dtype=torch.int64, device=dev); _ffi.fr_stream_dev(f.data_ptr(), 42, n)
work = torch.empty((n, 4), dtype=torch.int64, device=dev)
w = torch.zeros(12, dtype=torch.int64, device=dev)
def call():
    return _ffi.witness_eval_dev(srs, f.data_ptr(), log_d, 12345, work.data_ptr(), w.data_ptr(), 0)
for _ in range(3): call()
torch.cuda.synchronize()
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kzg_b200", "csrc", "evalform.cu")).read().split("\n")
line = next(i for i, l in enumerate(src) if "div_by_omega_kernel<<<" in l) + 1
acc = []
for _ in range(5):
    _ffi.trace_enable(True); call(); rec = _ffi.trace_read(); _ffi.trace_enable(False)
    acc.append([ms for ln, ms in rec if abs(ln - line) <= 2][0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): call()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(_ffi.LIB_PATH), "div_by_omega_us": round(1e3 * sum(acc) / len(acc), 1),
                  "witness_eval_ms": round(e0.elapsed_time(e1) / 10, 3),
                  "digest": hashlib.sha256(w.cpu().numpy().tobytes() + work.cpu().numpy().tobytes()).hexdigest()[:16]}), flush=True)
