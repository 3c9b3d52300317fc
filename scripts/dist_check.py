"""torchrun script: sharded MSM parity (closed form) and four-step NTT 2^24 parity + timing across ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from kzg_b200 import _ffi, distributed as kd, scalar
from oracle import cpu

rank, world, local = kd.init_process_group_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
_ffi.init(local)
LOG = int(sys.argv[1]) if len(sys.argv) > 1 else 24
LOG_N1 = int(sys.argv[2]) if len(sys.argv) > 2 else None   # four-step split (default: four_step_shape's choice)
# ---- sharded MSM over world * 2^18 points ----
per = 1 << 18
total = per * world
s = cpu.fr_stream(69, 1)[0]
srs = _ffi.srs_generate(0, s, total, first=rank * per, count=per)
sc = torch.empty((per, 4), dtype=torch.int64, device=dev)
_ffi.fr_stream_dev(sc.data_ptr(), 42, per, start=rank * per, stream=torch.cuda.current_stream().cuda_stream)
out = kd.sharded_msm(srs, sc)
torch.cuda.synchronize()
if rank == 0:
    coeffs = cpu.fr_stream(42, total)
    expect = cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), cpu.poly_eval(coeffs, s)))[0]
    print("sharded MSM parity (%d ranks, %d points):" % (world, total), bool((out.cpu().numpy().view(np.uint64) == expect).all()), flush=True)
# ---- K polynomials at once (pipelined batch + one all-gather + batched combine), device and host scalars ----
scs = []
for b in range(3):
    t = torch.empty((per, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(t.data_ptr(), 42 + b, per, start=rank * per, stream=torch.cuda.current_stream().cuda_stream)
    scs.append(t)
many = kd.sharded_msm_many(srs, scs)
many_h = kd.sharded_msm_many(srs, [t.cpu().numpy().view(np.uint64) for t in scs], device=dev)
torch.cuda.synchronize()
if rank == 0:
    oks = []
    for b in range(3):
        cf = cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), cpu.poly_eval(cpu.fr_stream(42 + b, total), s)))[0]
        oks.append(bool((many[b].cpu().numpy().view(np.uint64) == cf).all()) and bool((many_h[b].cpu().numpy().view(np.uint64) == cf).all()))
    print("sharded_msm_many parity, device and host scalars (%d ranks, 3 polynomials):" % world, oks, flush=True)
# ---- four-step NTT 2^LOG ----
n = 1 << LOG
l1, l2, lc = kd.four_step_shape(LOG, world, LOG_N1)
n1, c = 1 << l1, 1 << lc
# generate this rank's slab directly: slab[i1][cc] = a[i1*N2 + rank*C + cc] ; build from the full stream on host (rank-local slices)
full = cpu.fr_stream(5, n) if rank == 0 else None
slab = torch.empty((n1, c, 4), dtype=torch.int64, device=dev)
for i1 in range(n1):   # device-side generation row by row (stream positions are contiguous per row segment)
    _ffi.fr_stream_dev(slab[i1].data_ptr(), 5, c, start=i1 * (1 << l2) + rank * c, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
work = slab.clone()
rows = kd.ntt_four_step(work, LOG, False, log_n1=LOG_N1)
torch.cuda.synchronize()
gathered = [torch.empty_like(rows) for _ in range(world)]
dist.all_gather(gathered, rows) if world > 1 else gathered.__setitem__(0, rows)
if rank == 0:
    got = kd.four_step_gather([g.cpu().numpy().view(np.uint64) for g in gathered], LOG, world, LOG_N1)
    ref_t = torch.from_numpy(full.view(np.int64)).to(dev)
    tmp = torch.empty_like(ref_t)
    _ffi.ntt_dev(ref_t.data_ptr(), LOG, False, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("four-step NTT 2^%d parity vs single-GPU:" % LOG, bool((got == ref_t.cpu().numpy().view(np.uint64)).all()), flush=True)
    del ref_t, tmp
# timing (max over ranks)
for _ in range(2):
    work.copy_(slab); kd.ntt_four_step(work, LOG, False, log_n1=LOG_N1)
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
tot = 0.0
for _ in range(K):
    work.copy_(slab)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0.record(); kd.ntt_four_step(work, LOG, False, log_n1=LOG_N1); e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
t = torch.tensor([tot / K], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("four-step NTT 2^%d on %d GPU(s): %.3f ms (max over ranks)" % (LOG, world, t.item()), flush=True)
# ---- fused exchange: column pass stores into the peers' row matrices (CUDA IPC over NVLink) ----
plan = kd.FourStepPeerPlan(LOG, log_n1=LOG_N1)
ok = True
for it in range(4):           # alternates the two row matrices; inverse on odd calls
    inv = bool(it & 1)
    work.copy_(slab)
    ref = kd.ntt_four_step(work, LOG, inv, log_n1=LOG_N1)
    fused = kd.ntt_four_step_fused(slab, plan, inv)
    ok = ok and bool(torch.equal(ref, fused))
torch.cuda.synchronize()
okt = torch.tensor([1 if ok else 0], device=dev)
if world > 1: dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print("fused four-step NTT 2^%d == all-to-all version on every rank (4 calls, fwd/inv):" % LOG, bool(okt.item()), flush=True)
for _ in range(3): kd.ntt_four_step_fused(slab, plan, False)
torch.cuda.synchronize()
if world > 1: dist.barrier()
K = 10
tot = 0.0
for _ in range(K):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0.record(); kd.ntt_four_step_fused(slab, plan, False); e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
t = torch.tensor([tot / K], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
# back-to-back (no host sync between calls): K calls in one timed region
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0.record()
for _ in range(K): kd.ntt_four_step_fused(slab, plan, False)
e1.record()
torch.cuda.synchronize()
t2 = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t2, op=dist.ReduceOp.MAX)
if rank == 0:
    print("fused four-step NTT 2^%d on %d GPU(s): %.3f ms per isolated call, %.3f ms back to back (max over ranks)"
          % (LOG, world, t.item(), t2.item()), flush=True)
plan.close()
if world > 1: dist.destroy_process_group()
