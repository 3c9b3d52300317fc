"""CPU suite, part 3: the N > 1 orchestration (index ranges, layouts, exchange order) with world_size = 2 over
gloo.  The local kernels are replaced by oracle-backed stand-ins defined HERE (test infrastructure); the package
itself has no CPU implementation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kzg_b200 import distributed as kd
from oracle import kzg_oracle as O

R = O.R_MOD


def _to_np(t):
    return t.numpy().view(np.uint64)


def _from_np(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64))


class OracleMsmOps:
    """msm_partial / g1_combine with the C oracle: partial = Jacobian point padded to the 192-byte XYZZ slot."""

    def __init__(self, cpu):
        self.cpu = cpu

    def msm_partial(self, srs_shard, scalars, out_xyzz):
        p = self.cpu.msm_g1(srs_shard, _to_np(scalars), 1)
        out = np.zeros(24, dtype=np.uint64)
        out[:18] = p
        out_xyzz.copy_(_from_np(out))

    def g1_combine(self, partials, count, out_affine):
        acc = None
        arr = _to_np(partials).reshape(count, 24)
        for k in range(count):
            acc = arr[k, :18].copy() if acc is None else self.cpu.g1_add(acc, arr[k, :18])
        out_affine.copy_(_from_np(self.cpu.g1_to_affine(acc)[0]))


class OracleNttOps:
    def __init__(self, cpu):
        self.cpu = cpu

    def columns(self, slab, log_n, log_n1, log_c, col_offset, inverse):
        a = _to_np(slab).reshape(1 << log_n1, 1 << log_c, 4)
        omega = O.compute_omega(1 << log_n)[2]
        if inverse:
            omega = pow(omega, -1, R)
        for c in range(1 << log_c):
            col = np.ascontiguousarray(a[:, c])
            # unscaled transform with omega^(+-N/N1): forward fft, or inverse fft times N1
            f = self.cpu.fft(col, log_n1, inverse, 1)
            vals = [O.fr_from_mont_limbs(r) for r in f]
            for k1 in range(1 << log_n1):
                v = vals[k1] * ((1 << log_n1) if inverse else 1) % R
                v = v * pow(omega, k1 * (col_offset + c), R) % R
                a[k1, c] = np.array(O.fr_to_mont_limbs(v), dtype=np.uint64)
        slab.copy_(_from_np(a.reshape(slab.shape)))

    def unpack(self, recv, world, nrows, cols):
        r = _to_np(recv).reshape(world, nrows, cols, 4)
        return _from_np(np.ascontiguousarray(r.transpose(1, 0, 2, 3)).reshape(nrows, world * cols, 4))

    def rows(self, rows, log_n2, inverse, scale):
        a = _to_np(rows)
        for r in range(a.shape[0]):
            f = self.cpu.fft(np.ascontiguousarray(a[r]), log_n2, inverse, 1)
            if inverse:   # cpu.fft already multiplied by N2^-1: replace it by the requested scale
                sc = O.fr_from_mont_limbs(scale) * (1 << log_n2) % R
                f = np.array([O.fr_to_mont_limbs(O.fr_from_mont_limbs(x) * sc % R) for x in f], dtype=np.uint64)
            a[r] = f
        rows.copy_(_from_np(a))


class OraclePeerOps(OracleNttOps):
    """Stand-in for CudaPeerOps on CPU: "peer memory" is a named POSIX shared-memory segment (the name plays the
    role of the CUDA IPC handle), pointers are indices into a local registry."""

    def __init__(self, cpu):
        super().__init__(cpu)
        self.segs = {}

    def peer_alloc(self, nbytes):
        from multiprocessing import shared_memory
        seg = shared_memory.SharedMemory(create=True, size=nbytes)
        ptr = len(self.segs) + 1
        self.segs[ptr] = seg
        return ptr, seg.name.encode()

    def peer_open(self, handle):
        from multiprocessing import shared_memory
        seg = shared_memory.SharedMemory(name=handle.decode())
        ptr = len(self.segs) + 1
        self.segs[ptr] = seg
        return ptr

    def peer_close(self, ptr):
        self.segs.pop(ptr).close()

    def peer_free(self, ptr):
        seg = self.segs.pop(ptr)
        seg.close()
        seg.unlink()

    def _view(self, ptr, nbytes=None):
        seg = self.segs[ptr]
        return np.ndarray(((nbytes or seg.size) // 8,), dtype=np.uint64, buffer=seg.buf)

    def wrap(self, ptr, nbytes):
        return torch.from_numpy(self._view(ptr, nbytes).view(np.int64))

    def token(self):
        return torch.zeros(1, dtype=torch.int32)

    def drain(self):
        pass

    def columns_scatter(self, slab, peers, rank, log_n, log_n1, log_c, col_offset, inverse):
        tmp = slab.clone()
        self.columns(tmp, log_n, log_n1, log_c, col_offset, inverse)      # the unfused column pass ...
        a = _to_np(tmp).reshape(1 << log_n1, 1 << log_c, 4)
        world, C = len(peers), 1 << log_c
        rpr = (1 << log_n1) // world
        n2 = world * C
        for k1 in range(1 << log_n1):                                    # ... stored where the fused kernel stores it
            h, r = divmod(k1, rpr)
            dst = self._view(peers[h]).reshape(rpr, n2, 4)
            dst[r, rank * C:(rank + 1) * C] = a[k1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from oracle import cpu
    r, w, _ = kd.init_process_group_from_env("gloo")
    assert (r, w) == (rank, world)
    try:
        # ---- sharded MSM: rank g holds points [lo, hi) and the matching scalars ----
        n = 96
        s = cpu.fr_stream(69, 1)[0]
        gs = cpu.setup_g1(s, n)
        sc = cpu.fr_stream(42, n)
        lo, hi = kd.shard_range(n, world, rank)
        out = kd.sharded_msm(gs[lo:hi], _from_np(sc[lo:hi]), ops=OracleMsmOps(cpu))
        expect = cpu.g1_to_affine(cpu.msm_g1(gs, sc, 1))[0]
        ok_msm = bool((_to_np(out) == expect).all())
        # ---- K polynomials at once: one all-gather of K partials ----
        sc2 = cpu.fr_stream(43, n)
        outs = kd.sharded_msm_many(gs[lo:hi], [_from_np(sc[lo:hi]), _from_np(sc2[lo:hi])], ops=OracleMsmOps(cpu))
        ok_msm = ok_msm and bool((_to_np(outs[0]) == expect).all())
        ok_msm = ok_msm and bool((_to_np(outs[1]) == cpu.g1_to_affine(cpu.msm_g1(gs, sc2, 1))[0]).all())
        # ---- four-step NTT, forward and inverse ----
        log_n = 8
        vec = cpu.fr_stream(5, 1 << log_n)
        oks = []
        for inverse in (False, True):
            slabs = kd.four_step_scatter(vec, log_n, world, log_n1=3)
            rows = kd.ntt_four_step(_from_np(slabs[rank]).clone(), log_n, inverse, ops=OracleNttOps(cpu), log_n1=3)
            gathered = [torch.empty_like(rows) for _ in range(world)]
            dist.all_gather(gathered, rows)
            full = kd.four_step_gather([_to_np(g) for g in gathered], log_n, world, log_n1=3)
            oks.append(bool((full == cpu.fft(vec, log_n, inverse, 1)).all()))
        # ---- fused four-step (peer-memory stores instead of all-to-all + unpack): same blocks, slab untouched ----
        plan = kd.FourStepPeerPlan(log_n, log_n1=3, ops=OraclePeerOps(cpu))
        for inverse in (False, True, False):   # three calls: both row matrices and the reuse of the first
            slabs = kd.four_step_scatter(vec, log_n, world, log_n1=3)
            mine = _from_np(slabs[rank]).clone()
            rows = kd.ntt_four_step_fused(mine, plan, inverse).clone()
            assert (_to_np(mine) == slabs[rank]).all()
            gathered = [torch.empty_like(rows) for _ in range(world)]
            dist.all_gather(gathered, rows)
            full = kd.four_step_gather([_to_np(g) for g in gathered], log_n, world, log_n1=3)
            oks.append(bool((full == cpu.fft(vec, log_n, inverse, 1)).all()))
        plan.close()
        q.put((rank, ok_msm, oks))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n, w in ((10, 3), (1 << 20, 8), (5, 8), (7, 1)):
        spans = [kd.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_four_step_layout_roundtrip():
    log_n, world = 6, 4
    l1, l2, lc = kd.four_step_shape(log_n, world, log_n1=3)
    assert (l1, l2, lc) == (3, 3, 1)
    vec = np.arange((1 << log_n) * 4, dtype=np.uint64).reshape(-1, 4)
    slabs = kd.four_step_scatter(vec, log_n, world, log_n1=3)
    assert slabs[2].shape == (8, 2, 4)
    # slab_g[i1][c] = a[i1*N2 + g*C + c]
    assert (slabs[2][5, 1] == vec[5 * 8 + 2 * 2 + 1]).all()
    # output blocks: rows_h[r][k2] = X[(h*N1/G + r) + N1*k2]
    X = vec
    blocks = [np.stack([[X[(h * 2 + r) + 8 * k2] for k2 in range(8)] for r in range(2)]) for h in range(world)]
    assert (kd.four_step_gather(blocks, log_n, world, log_n1=3) == X).all()


@pytest.mark.parametrize("log_n,world,log_n1", [(6, 1, 3), (6, 2, 2), (7, 4, 3), (8, 8, 4), (8, 4, 6), (9, 2, None)])
def test_four_step_layouts_any_world_single_process(log_n, world, log_n1):
    """The index algebra of both four-step variants for several world sizes and splits, every rank emulated in this
    process with the oracle-backed stand-ins: all-to-all variant (columns -> row blocks -> unpack -> rows) and fused
    variant (columns_scatter straight into the owners' row matrices) must both equal the plain transform."""
    from oracle import cpu
    cpu.build()
    vec = cpu.fr_stream(5, 1 << log_n)
    l1, l2, lc = kd.four_step_shape(log_n, world, log_n1)
    n1, c = 1 << l1, 1 << lc
    nrows = n1 // world
    ops = OraclePeerOps(cpu)
    for inverse in (False, True):
        expect = cpu.fft(vec, log_n, inverse, 1)
        scale = np.array(O.fr_to_mont_limbs(pow(1 << log_n, -1, R)), dtype=np.uint64) if inverse else None
        slabs = [_from_np(s).clone() for s in kd.four_step_scatter(vec, log_n, world, log_n1)]
        # all-to-all variant
        done = [t.clone() for t in slabs]
        for g in range(world):
            ops.columns(done[g], log_n, l1, lc, g * c, inverse)
        blocks = []
        for h in range(world):
            recv = torch.stack([done[g].view(world, nrows, c, 4)[h] for g in range(world)]).contiguous()
            rows = ops.unpack(recv, world, nrows, c)
            ops.rows(rows, l2, inverse, scale)
            blocks.append(_to_np(rows))
        assert (kd.four_step_gather(blocks, log_n, world, log_n1) == expect).all()
        # fused variant: every rank stores into every owner's matrix
        allocs = [ops.peer_alloc(nrows * (1 << l2) * 32) for _ in range(world)]
        peers = [p for p, _ in allocs]
        for g in range(world):
            ops.columns_scatter(slabs[g], peers, g, log_n, l1, lc, g * c, inverse)
        blocks = []
        for h in range(world):
            rows = ops.wrap(peers[h], nrows * (1 << l2) * 32).view(nrows, 1 << l2, 4)
            ops.rows(rows, l2, inverse, scale)
            blocks.append(_to_np(rows).copy())
        del rows
        assert (kd.four_step_gather(blocks, log_n, world, log_n1) == expect).all()
        for p in peers:
            ops.peer_free(p)


def test_world_size_2_gloo_msm_and_ntt():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_msm, oks in results:
        assert ok_msm, f"rank {rank}: sharded MSM mismatch"
        assert oks == [True] * 5, f"rank {rank}: four-step NTT mismatch {oks} (unfused fwd/inv, fused fwd/inv/fwd)"
