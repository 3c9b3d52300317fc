"""Multi-GPU paths (one process per GPU, torch.distributed for the plumbing).

* MSM: shards naturally.  Rank g permanently holds SRS points [g*n/G, (g+1)*n/G) (and their window-multiple
  table); per call it runs the full single-GPU Pippenger on its slice of the scalars and produces ONE partial sum
  (raw XYZZ, 192 bytes).  The only exchange is an all-gather of G x 192 bytes followed by G-1 point additions
  and the affine conversion on every rank (there is no elliptic-curve ncclRedOp).
* NTT >= 2^24: four-step split with ONE all-to-all (see `ntt_four_step`).

The local compute is injected through a small `ops` object so that the orchestration (index ranges, layouts,
exchange order) is testable with the gloo backend on CPU; the default ops are the CUDA kernels behind the C ABI
and there is no CPU implementation in this package."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous index range [lo, hi) of rank `rank` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_process_group_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises torch.distributed when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


# ---------------------------------------------------------------------------
# C-ABI communicator (NCCL inside libkzg_b200.so): the product path on GPUs.  torch.distributed only carries the
# 128-byte id during bring-up; the collectives of the data path are issued by the library itself
# (kzgb_msm_g1_sharded*, kzgb_ntt_fr_sharded*), so a Rust / C++ caller reaches the same code with no Python.
# ---------------------------------------------------------------------------
_comm_ready = False


def comm_init(group=None) -> Tuple[int, int]:
    """Initialises the library's communicator for the ranks of `group` (default: the world).  Returns (rank, world)."""
    global _comm_ready
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if not _comm_ready:
        uid = [_ffi.comm_unique_id() if (rank == 0 and world > 1) else None]
        if world > 1:
            dist.broadcast_object_list(uid, src=0, group=group)
        _ffi.comm_init(rank, world, uid[0])
        _comm_ready = True
    return rank, world


def comm_destroy() -> None:
    global _comm_ready
    if _comm_ready:
        _ffi.comm_destroy()
        _comm_ready = False


def commit_sharded(srs_shard, scalars_shard) -> np.ndarray:
    """KZGProver::commit (coeff_form.rs:59-64) over a point-sharded SRS, host scalars in, affine point out (u64[12])
    on every rank: kzgb_msm_g1_sharded."""
    return _ffi.msm_sharded(srs_shard, scalars_shard)


def commit_sharded_dev(srs_shard, scalars_shard: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device-resident variant: kzgb_msm_g1_sharded_dev on the current stream -> int64[12] affine."""
    if out is None:
        out = torch.zeros(12, dtype=torch.int64, device=scalars_shard.device)
    _ffi.msm_sharded_dev(srs_shard, scalars_shard.data_ptr(), scalars_shard.shape[0], out.data_ptr(), 0,
                         torch.cuda.current_stream().cuda_stream)
    return out


def commit_sharded_many_dev(srs_shard, scalars_shards, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K sharded commits in one call (pipelined local MSMs, one all-gather): kzgb_msm_g1_sharded_batch_dev."""
    k = len(scalars_shards)
    if out is None:
        out = torch.zeros((k, 12), dtype=torch.int64, device=scalars_shards[0].device)
    _ffi.msm_sharded_batch_dev(srs_shard, [t.data_ptr() for t in scalars_shards], [t.shape[0] for t in scalars_shards],
                               out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    return out


class CudaOps:
    """Local kernels behind the C ABI, operating on torch CUDA tensors (int64 views of the blst layouts)."""

    device_type = "cuda"

    def msm_partial_many(self, srs: _ffi.SRS, scalars_list, out_xyzz: torch.Tensor) -> None:
        srs.msm_batch_dev([t.data_ptr() for t in scalars_list], [t.shape[0] for t in scalars_list], out_xyzz.data_ptr(), 2,
                          torch.cuda.current_stream().cuda_stream)

    def msm_partial_many_host(self, srs: _ffi.SRS, host_scalars_list, out_xyzz: torch.Tensor) -> None:
        srs.msm_batch_h2d(host_scalars_list, out_xyzz.data_ptr(), 2, torch.cuda.current_stream().cuda_stream)

    def msm_affine_many(self, srs: _ffi.SRS, scalars_list, out_affine: torch.Tensor) -> None:
        srs.msm_batch_dev([t.data_ptr() for t in scalars_list], [t.shape[0] for t in scalars_list], out_affine.data_ptr(), 0,
                          torch.cuda.current_stream().cuda_stream)

    def msm_partial(self, srs: _ffi.SRS, scalars: torch.Tensor, out_xyzz: torch.Tensor) -> None:
        srs.msm_dev(scalars.data_ptr(), scalars.shape[0], out_xyzz.data_ptr(), 2,
                    torch.cuda.current_stream().cuda_stream)

    def g1_combine(self, partials: torch.Tensor, count: int, out_affine: torch.Tensor) -> None:
        _ffi.g1_combine_dev(partials.data_ptr(), count, out_affine.data_ptr(), 0,
                            torch.cuda.current_stream().cuda_stream)

    def g1_combine_many(self, gathered: torch.Tensor, count: int, items: int, out_affine: torch.Tensor) -> None:
        """gathered: [count ranks][items][24] -> out_affine [items][12], one launch pair for all items."""
        _ffi.g1_combine_batch_dev(gathered.data_ptr(), count, items, out_affine.data_ptr(), 0,
                                  torch.cuda.current_stream().cuda_stream)


def sharded_msm(srs_shard, scalars_shard: torch.Tensor, ops=None, group=None) -> torch.Tensor:
    """commit over a point-sharded SRS: every rank passes ITS slice of the scalars (aligned with its SRS shard)
    and receives the affine result (int64[12]).  Exchange: all-gather of one 192-byte partial per rank."""
    ops = ops or CudaOps()
    dev = scalars_shard.device
    partial = torch.zeros(24, dtype=torch.int64, device=dev)
    ops.msm_partial(srs_shard, scalars_shard, partial)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        gathered = torch.empty(world * 24, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, partial, group=group)
    else:
        gathered = partial
    out = torch.zeros(12, dtype=torch.int64, device=dev)
    ops.g1_combine(gathered, world, out)
    return out


def sharded_msm_many(srs_shard, scalars_shards, ops=None, group=None, device=None) -> torch.Tensor:
    """`sharded_msm` for K polynomials at once: the K local MSMs run through the pipelined batch entry point, the K
    partials travel in ONE all-gather (G x K x 192 bytes) and are combined on every rank -> int64[K, 12].
    `scalars_shards` are device tensors, or (with `device` given) host numpy arrays [m, 4] u64 whose copies are
    pipelined with the MSMs."""
    ops = ops or CudaOps()
    k = len(scalars_shards)
    from_host = device is not None
    dev = device if from_host else scalars_shards[0].device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = torch.zeros((k, 12), dtype=torch.int64, device=dev)
    if world == 1 and hasattr(ops, "msm_affine_many") and not from_host:
        ops.msm_affine_many(srs_shard, scalars_shards, out)
        return out
    partials = torch.zeros((k, 24), dtype=torch.int64, device=dev)
    if from_host:
        ops.msm_partial_many_host(srs_shard, scalars_shards, partials)
    elif hasattr(ops, "msm_partial_many"):
        ops.msm_partial_many(srs_shard, scalars_shards, partials)
    else:
        for i, sc in enumerate(scalars_shards):
            ops.msm_partial(srs_shard, sc, partials[i])
    if world > 1:
        gathered = torch.empty((world, k, 24), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered.view(-1), partials.view(-1), group=group)
    else:
        gathered = partials.view(1, k, 24)
    if hasattr(ops, "g1_combine_many"):
        ops.g1_combine_many(gathered, world, k, out)
    else:
        per_item = gathered.transpose(0, 1).contiguous()      # [k][world][24]
        for i in range(k):
            ops.g1_combine(per_item[i].reshape(-1), world, out[i])
    return out


# ---------------------------------------------------------------------------
# Four-step NTT across GPUs (N >= 2^24; one all-to-all)
# ---------------------------------------------------------------------------
def four_step_shape(log_n: int, world: int, log_n1: Optional[int] = None) -> Tuple[int, int, int]:
    """(log_n1, log_n2, log_c): the vector is an N1 x N2 matrix (index i1*N2 + i2); every rank holds C = N2/world
    columns.  N1 <= 2^10 keeps the column transform in one shared-memory pass with 64-byte row segments."""
    lw = world.bit_length() - 1
    assert 1 << lw == world, "world size must be a power of two"
    if log_n1 is None:
        log_n1 = min(10, log_n // 2)
    log_n1 = max(log_n1, lw)          # at least one row block per rank
    log_n2 = log_n - log_n1
    assert log_n2 >= lw, "transform too small for this many ranks"
    return log_n1, log_n2, log_n2 - lw


def four_step_scatter(vec: np.ndarray, log_n: int, world: int, log_n1: Optional[int] = None):
    """Natural-order vector [N, 4] -> per-rank input slabs [N1, C, 4] (slab_g[i1][c] = a[i1*N2 + g*C + c])."""
    l1, l2, lc = four_step_shape(log_n, world, log_n1)
    m = np.ascontiguousarray(vec).reshape(1 << l1, world, 1 << lc, 4)
    return [np.ascontiguousarray(m[:, g]) for g in range(world)]


def four_step_gather(rows_per_rank, log_n: int, world: int, log_n1: Optional[int] = None) -> np.ndarray:
    """Per-rank outputs [N1/G, N2, 4] (rows_h[r][k2] = X[(h*N1/G + r) + N1*k2]) -> natural-order vector."""
    l1, l2, lc = four_step_shape(log_n, world, log_n1)
    full = np.concatenate([np.asarray(r).reshape(-1, 1 << l2, 4) for r in rows_per_rank], axis=0)  # [k1][k2]
    return np.ascontiguousarray(full.transpose(1, 0, 2)).reshape(1 << log_n, 4)                   # index k1 + N1*k2


class CudaNttOps:
    """Local steps of the four-step NTT behind the C ABI (torch CUDA tensors, int64 views of Fr)."""

    def columns(self, slab: torch.Tensor, log_n, log_n1, log_c, col_offset, inverse):
        _ffi.ntt4_columns_dev(slab.data_ptr(), log_n, log_n1, log_c, col_offset, inverse,
                              torch.cuda.current_stream().cuda_stream)

    def unpack(self, recv: torch.Tensor, world, nrows, cols) -> torch.Tensor:
        rows = torch.empty((nrows, world * cols, 4), dtype=torch.int64, device=recv.device)
        _ffi.ntt4_unpack_dev(recv.data_ptr(), rows.data_ptr(), world, nrows, cols, torch.cuda.current_stream().cuda_stream)
        return rows

    def rows(self, rows: torch.Tensor, log_n2, inverse, scale):
        tmp = torch.empty_like(rows) if log_n2 > 10 else rows
        _ffi.ntt_batch_dev(rows.data_ptr(), tmp.data_ptr(), log_n2, rows.shape[0], inverse, scale,
                           torch.cuda.current_stream().cuda_stream)


def ntt_four_step(slab: torch.Tensor, log_n: int, inverse: bool = False, ops=None, group=None,
                  log_n1: Optional[int] = None) -> torch.Tensor:
    """Distributed NTT.  `slab`: this rank's [N1, C, 4] input slab (see four_step_scatter); it is overwritten.
    Returns this rank's [N1/G, N2, 4] block of the result (see four_step_gather).  Same values as the single-GPU
    transform / the reference's fft, ifft (ft.rs:111-140).  One exchange: an all-to-all of N*32/G^2 bytes per pair."""
    ops = ops or CudaNttOps()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    l1, l2, lc = four_step_shape(log_n, world, log_n1)
    n1, c = 1 << l1, 1 << lc
    assert tuple(slab.shape) == (n1, c, 4), f"slab shape {tuple(slab.shape)} != {(n1, c, 4)}"
    ops.columns(slab, log_n, l1, lc, rank * c, inverse)
    if world > 1:
        recv = torch.empty_like(slab)
        dist.all_to_all_single(recv.view(-1), slab.view(-1), group=group)   # row block h of every rank -> rank h
    else:
        recv = slab
    rows = ops.unpack(recv, world, n1 // world, c)
    from .scalar import R_MOD, fr
    scale = fr(pow(1 << log_n, -1, R_MOD)) if inverse else None
    ops.rows(rows, l2, inverse, scale)
    return rows


# ---------------------------------------------------------------------------
# Fused exchange: the column pass stores straight into the peers' row matrices over NVLink
# ---------------------------------------------------------------------------
class _DeviceBuffer:
    """Wraps a raw device allocation so torch can view it (no copy) through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8", "data": (ptr, False), "version": 3}


class CudaPeerOps:
    """Peer memory (cudaMalloc + CUDA IPC) and the two kernels of the fused four-step NTT, behind the C ABI."""

    def peer_alloc(self, nbytes: int):
        return _ffi.peer_alloc(nbytes)

    def peer_open(self, handle: bytes) -> int:
        return _ffi.peer_open(handle)

    def peer_close(self, ptr: int) -> None:
        _ffi.peer_close(ptr)

    def peer_free(self, ptr: int) -> None:
        _ffi.peer_free(ptr)

    def wrap(self, ptr: int, nbytes: int) -> torch.Tensor:
        return torch.as_tensor(_DeviceBuffer(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device()))

    def token(self) -> torch.Tensor:
        return torch.zeros(1, dtype=torch.int32, device="cuda")

    def drain(self) -> None:
        torch.cuda.synchronize()

    def columns_scatter(self, slab, peers, rank, log_n, log_n1, log_c, col_offset, inverse):
        _ffi.ntt4_columns_scatter_dev(slab.data_ptr(), peers, rank, log_n, log_n1, log_c, col_offset, inverse,
                                      torch.cuda.current_stream().cuda_stream)

    def rows(self, rows: torch.Tensor, log_n2, inverse, scale):
        CudaNttOps().rows(rows, log_n2, inverse, scale)


class FourStepPeerPlan:
    """Peer-memory plan of the fused four-step NTT for one (log_n, world): every rank owns TWO row matrices
    [N1/G, N2] allocated with cudaMalloc and mapped into every other rank through CUDA IPC (one process per GPU,
    one node; NVLink/NVSwitch carries the stores).  Calls alternate between the two matrices, so a single barrier
    per call (after the column pass) is enough: a rank can only be writing matrix k%2 of call k after every rank has
    passed the barrier of call k-1, i.e. after every rank finished the row pass of call k-2 that last read it."""

    def __init__(self, log_n: int, group=None, log_n1: Optional[int] = None, ops=None):
        self.ops = ops or CudaPeerOps()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.log_n = log_n
        self.l1, self.l2, self.lc = four_step_shape(log_n, self.world, log_n1)
        self.nrows = (1 << self.l1) // self.world
        self.nbytes = self.nrows * (1 << self.l2) * 32
        self.calls = 0
        self.local = []          # my two matrices (pointers)
        handles = []
        for _ in range(2):
            p, h = self.ops.peer_alloc(self.nbytes)
            self.local.append(p)
            handles.append(h)
        gathered = [handles]
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, handles, group=group)
        self._opened = []
        self.peers = [[], []]    # peers[k][g] = matrix k of rank g as seen from this rank
        for k in range(2):
            for g in range(self.world):
                if g == self.rank:
                    self.peers[k].append(self.local[k])
                else:
                    q = self.ops.peer_open(gathered[g][k])
                    self._opened.append(q)
                    self.peers[k].append(q)
        self._token = None

    def rows_tensor(self, k: int) -> torch.Tensor:
        return self.ops.wrap(self.local[k], self.nbytes).view(self.nrows, 1 << self.l2, 4)

    def barrier(self):
        """Stream-ordered barrier (a 1-element all-reduce): no host synchronisation on the CUDA path."""
        if self.world > 1:
            if self._token is None:
                self._token = self.ops.token()
            dist.all_reduce(self._token, group=self.group)

    def close(self):
        if not self.local:
            return
        self.ops.drain()
        if self.world > 1:
            dist.barrier(group=self.group)     # nobody is still storing into a matrix that is about to go away
        for q in self._opened:
            self.ops.peer_close(q)
        if self.world > 1:
            dist.barrier(group=self.group)
        for p in self.local:
            self.ops.peer_free(p)
        self.local, self._opened = [], []


def ntt_four_step_fused(slab: torch.Tensor, plan: FourStepPeerPlan, inverse: bool = False) -> torch.Tensor:
    """Distributed NTT with the exchange fused into the column pass: same input slab and same result block as
    `ntt_four_step`, but the column kernel's epilogue stores each element directly into the row matrix of the rank
    that owns its row (peer memory over NVLink), already transposed -- no all-to-all, no unpack pass; one
    stream-ordered barrier separates the two passes.  The slab is only read.  Returns a view of the plan's row
    matrix (valid until the plan's next-but-one call)."""
    n1, c = 1 << plan.l1, 1 << plan.lc
    assert tuple(slab.shape) == (n1, c, 4), f"slab shape {tuple(slab.shape)} != {(n1, c, 4)}"
    k = plan.calls & 1
    plan.calls += 1
    plan.ops.columns_scatter(slab, plan.peers[k], plan.rank, plan.log_n, plan.l1, plan.lc, plan.rank * c, inverse)
    plan.barrier()
    rows = plan.rows_tensor(k)
    from .scalar import R_MOD, fr
    scale = fr(pow(1 << plan.log_n, -1, R_MOD)) if inverse else None
    plan.ops.rows(rows, plan.l2, inverse, scale)
    return rows
