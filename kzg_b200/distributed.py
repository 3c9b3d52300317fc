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
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


class CudaOps:
    """Local kernels behind the C ABI, operating on torch CUDA tensors (int64 views of the blst layouts)."""

    device_type = "cuda"

    def msm_partial(self, srs: _ffi.SRS, scalars: torch.Tensor, out_xyzz: torch.Tensor) -> None:
        srs.msm_dev(scalars.data_ptr(), scalars.shape[0], out_xyzz.data_ptr(), 2,
                    torch.cuda.current_stream().cuda_stream)

    def g1_combine(self, partials: torch.Tensor, count: int, out_affine: torch.Tensor) -> None:
        _ffi.g1_combine_dev(partials.data_ptr(), count, out_affine.data_ptr(), 0,
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
