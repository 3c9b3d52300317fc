"""torchrun script (config C5): one G1 MSM of 2^LOG points sharded over the ranks (strong scaling) -- parity with the
closed form p(s)*G and time per commit, one call at a time and through the pipelined batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from kzg_b200 import _ffi, distributed as kd
from oracle import cpu

rank, world, local = kd.init_process_group_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
_ffi.init(local)
LOG = int(sys.argv[1]) if len(sys.argv) > 1 else 24
total = 1 << LOG
per = total // world
s = cpu.fr_stream(69, 1)[0]
t0 = time.time()
srs = _ffi.srs_generate(0, s, total, first=rank * per, count=per)
torch.cuda.synchronize()
t_gen = time.time() - t0
st = torch.cuda.current_stream().cuda_stream
sc = torch.empty((per, 4), dtype=torch.int64, device=dev)
_ffi.fr_stream_dev(sc.data_ptr(), 42, per, start=rank * per, stream=st)
out = kd.sharded_msm(srs, sc)
torch.cuda.synchronize()
if rank == 0:
    coeffs = cpu.fr_stream(42, total)
    expect = cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), cpu.poly_eval(coeffs, s)))[0]
    print("MSM 2^%d over %d rank(s) (%d points each, window %d, table %.1f GB/rank, built in %.1f s): parity with p(s)*G: %s"
          % (LOG, world, per, srs.window_bits, srs.device_bytes / 1e9, t_gen, bool((out.cpu().numpy().view(np.uint64) == expect).all())), flush=True)
    del coeffs
def timed(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
ms_seq = timed(lambda: kd.sharded_msm(srs, sc), 5)
K = 6
ms_pipe = timed(lambda: kd.sharded_msm_many(srs, [sc] * K), 2) / K
if rank == 0:
    print("MSM 2^%d on %d GPU(s): %.3f ms per commit one call at a time (%.3e points/s), %.3f ms pipelined (%.3e points/s), max over ranks"
          % (LOG, world, ms_seq, total / ms_seq * 1e3, ms_pipe, total / ms_pipe * 1e3), flush=True)
if world > 1: dist.destroy_process_group()
