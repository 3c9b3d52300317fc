#!/usr/bin/env python
"""bench.py -- headline benchmark of the KZG prover hot path on B200 (contract: see DESIGN.md "Measurement").

metric : BLS12-381 G1 MSM points/sec (commit)   [BASELINE.json]; Fr NTT 2^22 ms and the other configs are in `extra`.
N = 1  : config C2.  A step = ONE KZGProver::commit (coeff_form.rs:59-64) of a degree-(2^20 - 1) polynomial = one call of
         the drop-in entry point: `value` = kzgb_msm_g1_dev (scalars resident in HBM, CUDA events on the launching
         stream), `e2e` = kzgb_msm_g1 (pinned host scalars in, 96-byte affine point out, both copies inside the timed
         region).  One call at a time -- the pipelined batch entry points are reported in `extra.batch`, not here.
N > 1  : config C5, STRONG scaling.  A step = one commit of a degree-(2^24 - 1) polynomial whose points and scalars are
         split evenly over the N ranks: kzgb_msm_g1_sharded_dev / kzgb_msm_g1_sharded (local Pippenger, ncclAllGather of
         N x 192 B inside the library, combine on every rank).  Rank 0 also times the SAME 2^24 commit on its GPU alone
         (`strong_scaling_base`), so the line carries its own denominator; `extra.ntt_2p24_four_step` is the four-step
         NTT 2^24 over the same ranks (fused peer-store exchange through the C ABI plan, and the NCCL all-to-all variant).
roofline: the dominant kernel (msm_accumulate_kernel), per launch, from CUDA events around the kernel on its stream,
         against the IMAD.WIDE rate measured live by kzgb_microbench (north_star defines the MSM roofline as INT32 IMAD).
cpu_baseline / --impl reference: the CPU restatement of multi_exp (oracle/kzg_cpu.c, all host threads) on the SAME
         points (s^i*G) and scalars as the GPU arm; a bounded sample of the workload only when the full size would not
         finish within minutes (then `same_config` is false and the sample is stated).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 20
N_POINTS = 1 << LOG_N
LOG_C5 = 24                             # config C5: MSM 2^24 / NTT 2^24 over 2/4/8 GPUs
IMAD_PER_MADD = 3000                    # XYZZ mixed add, 8M+2S at 300 IMAD per Fq product
IMAD_PER_MADD_ISSUED = 6 * 300 + 2 * 222 + 444   # six products, two squarings, Y3 as one fused dual product
NTT_LOG = 22
SECRET_SEED, POLY_SEED = 69, 42
METRIC = "BLS12-381 G1 MSM points/sec (commit, n=2^20)"
DTYPE = "u32 limbs (Fq 381-bit / Fr 255-bit Montgomery, integer)"


def canonical_imad_per_point(n):
    """MSM(n, c = 16) = n*W*3000 + W*2^c*4200 + W*(c*2700 + 4200) IMAD, per point."""
    W, c = 16, 16
    return (n * W * 3000 + W * (1 << c) * 4200 + W * (c * 2700 + 4200)) / n


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """Clocks / throttle reasons DURING the timed region: ONE `nvidia-smi -lms` poller per rank, started before the
    warm-up (so its start-up cost -- NVML initialisation touches every GPU of the box -- is outside the timed region)
    and killed at the end, as the profiling recipe prescribes; samples are time-stamped and those inside
    [begin(), end()] are summarised."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period_ms=50):
        self.rows = []          # (monotonic time, [fields])
        self.t0 = self.t1 = None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            if os.environ.get("KZGB_BENCH_NO_POLLER") == "1":   # diagnostic: is the poller perturbing host-synchronous loops?
                raise RuntimeError("poller disabled")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        if not self.proc:
            return
        for line in self.proc.stdout:
            f = [c.strip() for c in line.split(",")]
            if len(f) >= 6:
                self.rows.append((time.monotonic(), f))

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def close(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        inside = [f for (t, f) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        note = None
        if not inside and self.rows and self.t0 is not None:   # region shorter than the polling period: nearest sample
            inside = [min(self.rows, key=lambda r: abs(r[0] - (self.t0 + (self.t1 or self.t0)) / 2))[1]]
            note = "timed region shorter than the polling period: nearest sample"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(int(r[0]) for r in inside if r[0].isdigit())
        mx = [int(r[1]) for r in inside if r[1].isdigit()]
        reasons = [n for k, n in enumerate(self.NAMES) if any(r[2 + k] == "Active" for r in inside)]
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
               "samples": len(inside)}
        if note:
            out["note"] = note
        return out


def cpu_msm_time(gs_jac, scalars, reps=1):
    """Seconds per call of the CPU restatement of multi_exp (oracle/kzg_cpu.c kzgo_msm_g1, all host threads)."""
    from oracle import cpu
    cores = os.cpu_count() or 1
    cpu.msm_g1(gs_jac[: 1 << 12], scalars[: 1 << 12], cores)       # warm the thread pool / page the library in
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu.msm_g1(gs_jac, scalars, cores)
    return (time.perf_counter() - t0) / reps, cores


def run_reference(args):
    """--impl reference: the reference's CPU path (C restatement of multi_exp; the Rust crate cannot be built in this
    image) on the GPU arm's config: the same s^i*G points and seed-42 scalars.  N = 1: the 2^20 commit; N > 1: the 2^24
    commit (same total work as the sharded GPU arm).  The per-step sample is the full workload when K steps fit in
    ~3 minutes, else the largest power-of-two prefix that does."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu
    cpu.build()
    cores = os.cpu_count() or 1
    steps = max(1, args.steps)
    full_log = LOG_N if args.gpus <= 1 else LOG_C5
    secret = cpu.fr_stream(SECRET_SEED, 1)[0]
    cal_n = 1 << 14
    t_cal, _ = cpu_msm_time(cpu.setup_g1_fast(secret, cal_n), cpu.fr_stream(POLY_SEED, cal_n))
    budget_s = float(os.environ.get("KZGB_REF_BUDGET_S", "150"))
    sample_log = 14
    # cost grows ~linearly in n (a little slower: the window widens): extrapolate conservatively with n
    while sample_log < full_log and t_cal * (1 << (sample_log + 1 - 14)) * (steps + 1) <= budget_s:
        sample_log += 1
    m = 1 << sample_log
    t0 = time.perf_counter()
    gs = cpu.setup_g1_fast(secret, m)           # gs[i] = s^i*G, the GPU arm's SRS (threaded fixed-base generator)
    sc = cpu.fr_stream(POLY_SEED, m)
    setup_s = time.perf_counter() - t0
    for _ in range(min(args.warmup, 1)):
        cpu.msm_g1(gs, sc, cores)
    dt, cores = cpu_msm_time(gs, sc, steps)
    pps = m / dt
    same = sample_log == full_log
    sample = ("the full workload: MSM of 2^%d points s^i*G x seed-42 scalars" % full_log) if same else \
             ("MSM over the first 2^%d points/scalars of the 2^%d workload (the full size would take ~%.0f s per step)"
              % (sample_log, full_log, dt * (1 << (full_log - sample_log))))
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": "points/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak" if args.gpus <= 1 else "strong",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": ("C2 commit: G1 MSM 2^20" if args.gpus <= 1 else "C5 commit: G1 MSM 2^24 (total work of the sharded GPU arm)")
                   + ", CPU restatement of multi_exp (coeff_form.rs:61) on the host cores", "sample_log2": sample_log, "same_config": same,
                   "srs_setup_s": round(setup_s, 2)},
        "cpu_baseline": {"value": pps, "unit": "points/s", "cores": cores, "kind": "port", "same_config": same,
                         "sample": sample + "; %d step(s); bucket method, threads over windows/ranges; portable C, not blst assembly" % steps},
        "e2e": {"value": pps, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Timer:
    """CUDA-event timing on torch's current stream (the stream every library call below is given), max over ranks."""

    def __init__(self, torch, dist, dev, world):
        self.torch, self.dist, self.dev, self.world = torch, dist, dev, world
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup=0):
        """ms per step of fn(i), i = 0..steps-1, device-timed between two barriers; max over ranks."""
        for i in range(warmup):
            fn(i)
        self.barrier()
        self.e0.record()
        for i in range(steps):
            fn(i)
        self.e1.record()
        self.barrier()
        return self.max_over_ranks(self.e0.elapsed_time(self.e1)) / steps

    def wall(self, fn, steps, warmup=0):
        """ms per step by the host clock (for host-synchronous calls: copies and the final sync are inside)."""
        for i in range(warmup):
            fn(i)
        self.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        self.barrier()
        return self.max_over_ranks((time.perf_counter() - t0) * 1e3) / steps


def fail(msg):
    print(json.dumps({"error": msg}), flush=True)
    sys.exit(1)


def accumulate_traffic():
    """ncu DRAM bytes per launch of the dominant kernel -- from the committed capture, NOT measured in this run."""
    for name in ("r02_msm_accumulate_summary.json", "r01f_msm_accumulate_summary.json"):
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", name))).get("dram_bytes_per_launch")
            if v:
                return v, "profiles/%s (ncu --set full capture of this kernel at n = 2^20; not measured in this run)" % name
        except Exception:
            pass
    return None, None


def run_gpu(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from kzg_b200 import _ffi, distributed as kd, scalar

    rank, world, local = kd.init_process_group_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _ffi.init(local)
    kd.comm_init()                    # the library's own NCCL communicator (torch.distributed only ships the id)
    lib = _ffi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    clk = ClockSampler(local)         # one poller, started before the warm-up
    T = Timer(torch, dist, dev, world)
    from oracle import cpu as ocpu    # ONLY for the seeded secret, the parity gates and the cpu_baseline leg
    ocpu.build()
    secret = ocpu.fr_stream(SECRET_SEED, 1)[0]
    K, W = args.steps, args.warmup
    peaks = measured_peaks()
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)

    def device_scalars(count, seed, start):
        t = torch.empty((count, 4), dtype=torch.int64, device=dev)
        _ffi.fr_stream_dev(t.data_ptr(), seed, count, start=start, stream=stream)
        return t

    def closed_form_commit(coeffs_u64):
        return ocpu.g1_to_affine(ocpu.g1_mul(ocpu.g1_generator(), ocpu.poly_eval(coeffs_u64, secret)))[0]

    # on-box peak for the roofline denominator: dense independent IMAD.WIDE chains
    imad_peak = max(_ffi.microbench(3, 8, 256, 64)["ops_per_s"] for _ in range(3))
    extra = {}

    def stage_pass(fn, steps):
        """Runs fn(i) `steps` times with the library's stage timers on: {stage: ms per launch}."""
        _ffi.profile_enable(True)
        _ffi.profile_read()
        for i in range(steps):
            fn(i)
        prof = _ffi.profile_read()
        _ffi.profile_enable(False)
        return {k: v[0] / v[1] for k, v in prof.items()}

    def msm_roofline(stage_ms, points_per_launch, window_bits, ms_step):
        acc_ms = stage_ms.get("msm_accumulate")
        if not acc_ms:
            return None
        nwin = -(-256 // window_bits)
        achieved = 16 * IMAD_PER_MADD * points_per_launch / (acc_ms * 1e-3)
        traffic, tsrc = accumulate_traffic()
        return {"bound": "int32_imad", "kernel": "msm_accumulate_kernel", "achieved": achieved / 1e12, "peak": imad_peak / 1e12,
                "unit": "TIMAD/s (32x32->64 multiply-adds)", "frac": achieved / imad_peak,
                "traffic": traffic, "traffic_source": tsrc,
                "peak_source": "measured live: kzgb_microbench mode 3 (dense independent IMAD.WIDE.U32 chains)",
                "algorithmic_work": "16 windows x 3000 IMAD per point (SURVEY 8d, canonical c = 16) x %d points per launch" % points_per_launch,
                "executed_work": "%d windows x (6 x 300 + 2 x 222 + 444) IMAD actually issued per point at c = %d "
                                 "(Y3 = R(Q-X3) - Y1*PPP is one fused dual-product Montgomery pass)" % (nwin, window_bits),
                "frac_executed": nwin * IMAD_PER_MADD_ISSUED * points_per_launch / (acc_ms * 1e-3) / imad_peak,
                "kernel_ms": acc_ms, "share_of_step": acc_ms / ms_step,
                "timed_in": "a second pass of the same single-call steps with CUDA events around each stage on its stream"}

    if world == 1:
        line = bench_single(args, locals())
    else:
        line = bench_sharded(args, locals())
    clk.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    kd.comm_destroy()
    if world > 1:
        dist.destroy_process_group()


def bench_single(args, env):
    """N = 1: config C2 (+ the other single-GPU configs as extras)."""
    import ctypes as C
    import numpy as np
    torch, _ffi, ocpu, scalar, T, dev, stream, lib, clk = (env[k] for k in ("torch", "_ffi", "ocpu", "scalar", "T", "dev", "stream", "lib", "clk"))
    K, W, secret, extra, imad_peak, hbm_peak, peaks = (env[k] for k in ("K", "W", "secret", "extra", "imad_peak", "hbm_peak", "peaks"))
    srs = _ffi.srs_generate(0, secret, N_POINTS)
    NBUF = 8   # 8 distinct 32 MiB polynomials (256 MiB > 126 MB L2) rotated between steps
    polys = [env["device_scalars"](N_POINTS, POLY_SEED + b, 0) for b in range(NBUF)]
    out = torch.zeros(12, dtype=torch.int64, device=dev)

    def step(i):     # the drop-in commit, device-resident scalars: ONE call of kzgb_msm_g1_dev
        srs.msm_dev(polys[i % NBUF].data_ptr(), N_POINTS, out.data_ptr(), 0, stream)

    # ---- parity gate: commit == p(s)*G (closed form), compared as compressed G1 ----
    step(0)
    torch.cuda.synchronize()
    coeffs0 = polys[0].cpu().numpy().view(np.uint64)
    commit0 = out.cpu().numpy().view(np.uint64).copy()
    if scalar.g1_affine_compress(commit0) != ocpu.g1_affine_compress(env["closed_form_commit"](coeffs0)):
        fail("parity check failed: commit != p(s)*G")

    # ---- headline: K single calls, device-resident ----
    T.timed(step, W)
    launches0 = _ffi.launch_count()
    clk.begin()
    ms_step = T.timed(step, K)
    clk.end()
    launches = _ffi.launch_count() - launches0
    value = N_POINTS / (ms_step * 1e-3)
    stage_ms = env["stage_pass"](step, K)
    roofline = env["msm_roofline"](stage_ms, N_POINTS, srs.window_bits, ms_step)

    # ---- e2e: the host-pointer drop-in call, pinned host scalars, copies inside the timed region ----
    host_polys = [polys[b].cpu().pin_memory() for b in range(2)]
    host_views = [hp.numpy().view(np.uint64) for hp in host_polys]
    out_host = np.zeros(12, dtype=np.uint64)

    def e2e_step(i):
        _ffi.check(lib.kzgb_msm_g1(srs._h, C.c_void_p(host_polys[i % 2].data_ptr()), C.c_size_t(N_POINTS), _ffi.ptr(out_host)))
    e2e_ms = T.wall(e2e_step, K, warmup=max(W, 3))
    e2e_step(0)
    if not (out_host == commit0).all():
        fail("host-pointer commit differs from the device-resident commit")

    extra["stage_ms"] = {k: round(v, 4) for k, v in stage_ms.items()}
    extra["window_bits"] = srs.window_bits
    extra["srs_device_bytes"] = srs.device_bytes
    extra["msm_whole_frac_of_imad_peak"] = canonical_imad_per_point(N_POINTS) * N_POINTS / (ms_step * 1e-3) / imad_peak
    extra["issue_model"] = ("IMAD.WIDE issues at 32 lanes/clk/SM and IADD3/LOP3 at 128, but on this chip they share one issue "
                            "port per scheduler (microbench mode 7: a 1:1 IMAD.WIDE/IADD mix runs at the serial sum; the "
                            "Karatsuba product with 12 % fewer IMAD.WIDE and ~150 more ALU ops is 11 % slower): see DESIGN.md")

    def secondary():
        # ---- the same host-pointer call from ordinary pageable memory (what a Rust Vec<Fr> is): upload_async's relay ----
        pageable = [hp.numpy().view(np.uint64).copy() for hp in host_polys]

        def pageable_step(i):
            _ffi.check(lib.kzgb_msm_g1(srs._h, _ffi.ptr(pageable[i % 2]), C.c_size_t(N_POINTS), _ffi.ptr(out_host)))
        pg_ms = T.wall(pageable_step, K, warmup=3)
        extra["e2e_pageable_host"] = {"ms_per_step": pg_ms, "value": N_POINTS / (pg_ms * 1e-3),
                                      "note": "kzgb_msm_g1 from unpinned numpy memory: 32 MiB relayed through the lane's pinned ring "
                                              "by 4 host threads (KZGB_UPLOAD_THREADS); profiles/r02s_pageable_upload_relay.log"}
        # ---- pipelined batch entry points (NOT the headline): K commits per call ----
        d_out = torch.zeros((K, 12), dtype=torch.int64, device=dev)

        def batch_dev(_):
            srs.msm_batch_dev([polys[i % NBUF].data_ptr() for i in range(K)], [N_POINTS] * K, d_out.data_ptr(), 0, stream)
        b_ms = T.timed(batch_dev, 2, warmup=2) / K
        if not (d_out[0].cpu().numpy().view(np.uint64) == commit0).all():
            fail("pipelined batch result differs from the single call")
        srs.msm_batch([host_views[i % 2] for i in range(K)])      # warm-up: staging and pinned result buffers get allocated
        t0 = time.perf_counter()
        srs.msm_batch([host_views[i % 2] for i in range(K)])
        srs.msm_batch([host_views[i % 2] for i in range(K)])
        bh_ms = (time.perf_counter() - t0) / 2 / K * 1e3
        extra["batch"] = {"ms_per_commit_device_resident": b_ms, "value": N_POINTS / (b_ms * 1e-3),
                          "ms_per_commit_from_pinned_host": bh_ms, "e2e_value": N_POINTS / (bh_ms * 1e-3),
                          "note": "kzgb_msm_g1_batch{_dev}: %d commits per call, tail of commit i overlaps commit i+1 (an API the reference does not have)" % K}
        # ---- Fr NTT 2^22 (the second half of the BASELINE metric) ----
        nbuf = env["device_scalars"](1 << NTT_LOG, 5, 0)

        def ntt_step(_):
            _ffi.ntt_dev(nbuf.data_ptr(), NTT_LOG, False, stream)
        ntt_ms = T.timed(ntt_step, 10, warmup=3)
        ntt_stage = env["stage_pass"](ntt_step, 3)
        n = 1 << NTT_LOG
        algo_bytes = 2 * 2 * 32 * n                       # 2 passes x (read + write) x 32 B x N
        algo_imad = (n // 2) * NTT_LOG * 136              # butterflies x 136 IMAD
        extra["ntt_2p22_ms"] = ntt_ms
        extra["roofline_ntt"] = {"kernel": "ntt_pass_kernel x3", "ms": ntt_ms,
                                 "frac_hbm": algo_bytes / (ntt_ms * 1e-3) / 1e9 / hbm_peak, "achieved_GBps": algo_bytes / (ntt_ms * 1e-3) / 1e9,
                                 "peak_GBps": hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                                 "frac_imad": algo_imad / (ntt_ms * 1e-3) / imad_peak, "achieved_TIMADps": algo_imad / (ntt_ms * 1e-3) / 1e12,
                                 "peak_TIMADps": imad_peak / 1e12, "binding": "int32 issue slots (IMAD.WIDE), not HBM: SURVEY 8d 'honest note', DESIGN.md 6",
                                 "stage_ms": ntt_stage.get("ntt")}
        del nbuf
        # ---- linear quotient scan alone (the HBM-bound kernel of the path) ----
        xy = torch.from_numpy(np.vstack([ocpu.fr_stream(7, 1)[0], np.zeros(4, dtype=np.uint64)]).view(np.int64)).to(dev)
        work = torch.empty((N_POINTS, 4), dtype=torch.int64, device=dev)
        rem = torch.zeros(4, dtype=torch.int64, device=dev)

        def scan_step(i):
            lib.kzgb_divide_linear_dev(C.c_void_p(polys[i % NBUF].data_ptr()), C.c_size_t(N_POINTS), C.c_void_p(xy.data_ptr()), C.c_int(0),
                                       C.c_void_p(work.data_ptr()), C.c_void_p(rem.data_ptr()), C.c_void_p(stream))
        scan_ms = T.timed(scan_step, 20, warmup=3)
        extra["roofline_scan"] = {"kernel": "scan_divide_linear (3 phases)", "ms": scan_ms, "algorithmic_bytes": 64 * N_POINTS,
                                  "achieved_GBps": 64 * N_POINTS / (scan_ms * 1e-3) / 1e9, "peak_GBps": hbm_peak,
                                  "frac_hbm": 64 * N_POINTS / (scan_ms * 1e-3) / 1e9 / hbm_peak,
                                  "frac_imad": 2.6 * 136 * N_POINTS / (scan_ms * 1e-3) / imad_peak,
                                  "binding": "int32 issue slots: a parallel Horner scan needs >= 2 Fr products (136 IMAD each) per "
                                             "coefficient, this one 2.6 -- 41 us at the IMAD peak against 10 us of traffic (DESIGN.md 4.4)"}
        # ---- C2 witness: scan + MSM(2^20 - 1), device resident ----
        scan_step(0)
        torch.cuda.synchronize()
        xy[1] = rem  # y = p(x)
        wout = torch.zeros(12, dtype=torch.int64, device=dev)

        def wit(_):
            extra["witness_status"] = _ffi.witness_coeff_dev(srs, polys[0].data_ptr(), N_POINTS, xy.data_ptr(), work.data_ptr(), wout.data_ptr(), stream)
        extra["witness_2p20_ms"] = T.wall(wit, 5, warmup=2)
        # ---- f2: one polynomial uploaded once, then commit + 8 witnesses on the resident copy ----
        xs8 = ocpu.fr_stream(7, 8)
        ys8 = np.vstack([ocpu.poly_eval(host_views[0], xs8[i]) for i in range(8)])

        def resident(_):
            h = _ffi.DevicePoly(host_views[0])            # the only 32 MiB H2D
            h.commit(srs)
            st = [h.witness(srs, N_POINTS, xs8[i], ys8[i])[0] for i in range(8)]
            h.close()
            return st
        resident(0)
        t0 = time.perf_counter()
        st8 = resident(0)
        extra["resident_poly_commit_plus_8_witnesses_ms"] = (time.perf_counter() - t0) * 1e3
        extra["resident_poly_note"] = "kzgb_poly_upload once (32 MiB H2D), then kzgb_poly_commit + 8 x kzgb_poly_witness: 9 MSMs, 1 upload; statuses %s" % st8

        def hostptr(_):
            srs.msm(host_views[0])
            for i in range(8):
                _ffi.witness_coeff(srs, host_views[0], xs8[i], ys8[i])
        hostptr(0)
        t0 = time.perf_counter()
        hostptr(0)
        extra["hostptr_commit_plus_8_witnesses_ms"] = (time.perf_counter() - t0) * 1e3
        # multi-opening through the pipelined call
        _ffi.witness_coeff_many(srs, host_views[0], xs8, ys8)
        t0 = time.perf_counter()
        st8b, _w8 = _ffi.witness_coeff_many(srs, host_views[0], xs8, ys8)
        extra["witness_many_2p20_ms_per_point"] = (time.perf_counter() - t0) / 8 * 1e3
        # ---- C4: create_witness_batched, 256 openings, device resident ----
        m = 256
        xs_t = env["device_scalars"](m, 7, 0)
        ys_t = torch.empty((m, 4), dtype=torch.int64, device=dev)
        xy1 = torch.zeros((2, 4), dtype=torch.int64, device=dev)
        for i in range(m):  # y_i = p(x_i): remainders of linear scans
            xy1[0] = xs_t[i]
            lib.kzgb_divide_linear_dev(C.c_void_p(polys[0].data_ptr()), C.c_size_t(N_POINTS), C.c_void_p(xy1.data_ptr()), C.c_int(0),
                                       C.c_void_p(work.data_ptr()), C.c_void_p(ys_t[i].data_ptr()), C.c_void_p(stream))
        r_t = torch.zeros((m + 2, 4), dtype=torch.int64, device=dev)

        def c4(_):
            extra["witness_batched_status"] = _ffi.witness_batched_dev(srs, polys[0].data_ptr(), N_POINTS, xs_t.data_ptr(), ys_t.data_ptr(), m,
                                                                       r_t.data_ptr(), wout.data_ptr(), stream)
        extra["witness_batched_2p20_m256_ms"] = T.wall(c4, 3, warmup=2)
        # ---- C3: eval-form commit (Lagrange SRS) + iNTT 2^20 ----
        lag = _ffi.srs_generate(1, secret, N_POINTS)

        def c3(i):
            lag.msm_dev(polys[i % NBUF].data_ptr(), N_POINTS, out.data_ptr(), 0, stream)
        extra["evalform_commit_2p20_ms"] = T.timed(c3, 5, warmup=2)
        nb20 = polys[1].clone()

        def intt20(_):
            _ffi.ntt_dev(nb20.data_ptr(), 20, True, stream)
        extra["intt_2p20_ms"] = T.timed(intt20, 10, warmup=2)
        lag.close()

    def c5_single_gpu():
        """The C5 workloads on ONE GPU: the denominators of the strong-scaling runs (bench.py --gpus 2/4/8)."""
        n24 = 1 << LOG_C5
        t0 = time.perf_counter()
        big = _ffi.srs_generate(0, secret, n24)
        torch.cuda.synchronize()
        gen_s = time.perf_counter() - t0
        sc = env["device_scalars"](n24, POLY_SEED, 0)

        def big_step(_):
            big.msm_dev(sc.data_ptr(), n24, out.data_ptr(), 0, stream)
        ms = T.timed(big_step, 5, warmup=2)
        extra["c5_single_gpu"] = {"msm_2p24_ms": ms, "msm_2p24_value": n24 / (ms * 1e-3), "window_bits": big.window_bits,
                                  "srs_device_bytes": big.device_bytes, "srs_build_s": round(gen_s, 2)}
        big.close()

        def ntt24(_):
            _ffi.ntt_dev(sc.data_ptr(), LOG_C5, False, stream)
        extra["c5_single_gpu"]["ntt_2p24_ms"] = T.timed(ntt24, 5, warmup=2)

    for fn in (secondary, c5_single_gpu):
        try:   # secondary measurements never take the headline line down
            fn()
        except SystemExit:
            raise
        except Exception as ex:   # noqa: BLE001
            extra[fn.__name__ + "_error"] = repr(ex)[:300]

    # ---- CPU baseline: the same 2^20 points and scalars, all host threads, one run ----
    aff = srs.read_affine(0, N_POINTS)
    gs = np.zeros((N_POINTS, 18), dtype=np.uint64)
    gs[:, :12] = aff
    gs[:, 12:] = np.array([0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745,
                           0x5c071a97a256ec6d, 0x15f65ec3fa80e493], dtype=np.uint64)  # Montgomery one
    del aff
    dt, cores = cpu_msm_time(gs, coeffs0, 1)
    cpu_baseline = {"value": N_POINTS / dt, "unit": "points/s", "cores": cores, "kind": "port", "same_config": True, "ms": dt * 1e3,
                    "sample": "the full workload: MSM over the same 2^20 points (read back from the device SRS) and seed-42 scalars, 1 run, "
                              "all host threads; C restatement of multi_exp (not blst assembly)"}
    return {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": "C2 coeff-form commit: ONE call of the drop-in entry point per step (kzgb_msm_g1_dev / kzgb_msm_g1), "
                               "G1 MSM of a degree-(2^20-1) polynomial against a 2^20-point SRS",
                   "points": N_POINTS, "scalars": "full-width uniform Fr (splitmix64 seeds 42..49)", "srs": "s^i*G, secret seed 69",
                   "cache": "8 rotating 32 MiB scalar vectors (256 MiB) + a 1.3 GB point table: inputs larger than L2"},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": N_POINTS / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": N_POINTS * 32, "d2h_bytes_per_step": 96,
                "ms_per_step": e2e_ms, "call": "kzgb_msm_g1 (host pointers, pinned), one call per step"},
        "gpu_launches": launches, "clocks": clk.summary(), "extra": extra,
    }


def bench_sharded(args, env):
    """N > 1: config C5, strong scaling -- one 2^24 commit and one 2^24 NTT split over the ranks."""
    import ctypes as C
    import numpy as np
    torch, dist, _ffi, kd, ocpu, scalar, T, dev, stream, lib, clk = (env[k] for k in ("torch", "dist", "_ffi", "kd", "ocpu", "scalar", "T", "dev", "stream", "lib", "clk"))
    K, W, secret, extra, imad_peak, rank, world = (env[k] for k in ("K", "W", "secret", "extra", "imad_peak", "rank", "world"))
    total = 1 << LOG_C5
    per = total // world
    t0 = time.perf_counter()
    srs = _ffi.srs_generate(0, secret, total, first=rank * per, count=per)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    NBUF = max(2, min(8, (1 << 28) // (per * 32) + 1))    # rotate at least 256 MiB of scalars per rank
    polys = [env["device_scalars"](per, POLY_SEED + b, rank * per) for b in range(NBUF)]
    out = torch.zeros(12, dtype=torch.int64, device=dev)

    def step(i):     # the sharded drop-in commit: kzgb_msm_g1_sharded_dev (NCCL all-gather inside the library)
        _ffi.msm_sharded_dev(srs, polys[i % NBUF].data_ptr(), per, out.data_ptr(), 0, stream)

    step(0)
    torch.cuda.synchronize()
    commit0 = out.cpu().numpy().view(np.uint64).copy()
    if rank == 0:    # parity gate: the sharded commit == p(s)*G over all 2^24 coefficients
        if scalar.g1_affine_compress(commit0) != ocpu.g1_affine_compress(env["closed_form_commit"](ocpu.fr_stream(POLY_SEED, total))):
            fail("parity check failed: sharded commit != p(s)*G")
    T.timed(step, W)
    launches0 = _ffi.launch_count()
    clk.begin()
    ms_step = T.timed(step, K)
    clk.end()
    launches = _ffi.launch_count() - launches0
    value = total / (ms_step * 1e-3)
    stage_ms = env["stage_pass"](step, K)
    roofline = env["msm_roofline"](stage_ms, per, srs.window_bits, ms_step)
    # per-rank times of the same steps (load balance)
    my = torch.tensor([T.e0.elapsed_time(T.e1) / K], dtype=torch.float64, device=dev)
    allr = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allr, my)
    extra["ms_per_step_by_rank"] = [round(float(x.item()), 4) for x in allr]
    extra["stage_ms"] = {k: round(v, 4) for k, v in stage_ms.items()}
    extra["window_bits"] = srs.window_bits
    extra["srs_device_bytes_per_rank"] = srs.device_bytes
    extra["srs_build_s"] = round(gen_s, 2)

    # ---- e2e: pinned host scalars of this rank's shard -> kzgb_msm_g1_sharded ----
    host_polys = [polys[b].cpu().pin_memory() for b in range(2)]
    out_host = np.zeros(12, dtype=np.uint64)

    def e2e_step(i):
        _ffi.check(lib.kzgb_msm_g1_sharded(srs._h, C.c_void_p(host_polys[i % 2].data_ptr()), C.c_size_t(per), _ffi.ptr(out_host)))
    e2e_ms = T.wall(e2e_step, K, warmup=max(W, 3))
    e2e_step(0)
    if not (out_host == commit0).all():
        fail("host-pointer sharded commit differs from the device-resident one")

    def batch_extra():
        d_out = torch.zeros((K, 12), dtype=torch.int64, device=dev)

        def batch(_):
            _ffi.msm_sharded_batch_dev(srs, [polys[i % NBUF].data_ptr() for i in range(K)], [per] * K, d_out.data_ptr(), 0, stream)
        b_ms = T.timed(batch, 2, warmup=1) / K
        extra["batch"] = {"ms_per_commit_device_resident": b_ms, "value": total / (b_ms * 1e-3),
                          "note": "kzgb_msm_g1_sharded_batch_dev: %d commits per call, one all-gather of all partials" % K}

    def ntt_extra():
        plan = _ffi.Ntt4Plan(LOG_C5)
        slab_elems = (1 << plan.l1) << plan.lc
        slab = env["device_scalars"](slab_elems, 5, rank * slab_elems)

        def fused(_):
            plan.run_dev(slab.data_ptr(), False, stream)
        res = {"log_n": LOG_C5, "shape": [plan.l1, plan.l2, plan.lc], "fused_peer_store_ms": T.timed(fused, 10, warmup=3)}
        work4 = slab.view(1 << plan.l1, 1 << plan.lc, 4).clone()

        def a2a(_):
            kd.ntt_four_step(work4, LOG_C5, False)
        res["nccl_all_to_all_ms"] = T.timed(a2a, 10, warmup=3)
        # host vectors in natural order (scatter of the slab + transposed gather of this rank's block inside the call)
        if rank == 0:
            full = ocpu.fr_stream(5, 1 << LOG_C5)
        else:
            full = np.zeros((1 << LOG_C5, 4), dtype=np.uint64)
        ft_dev = torch.from_numpy(full.view(np.int64)).to(dev)
        dist.broadcast(ft_dev, src=0)
        full_pinned = ft_dev.cpu().pin_memory()
        del ft_dev, full
        outv = torch.zeros_like(full_pinned).pin_memory()

        def host_call(_):
            _ffi.check(lib.kzgb_ntt_fr_sharded(plan._h, C.c_void_p(full_pinned.data_ptr()), C.c_void_p(outv.data_ptr()), C.c_int(0), C.c_int(0)))
        res["host_vector_own_block_ms"] = T.wall(host_call, 3, warmup=1)
        res["host_vector_bytes_per_rank"] = {"h2d": slab_elems * 32, "d2h": slab_elems * 32}
        res["note"] = ("strong scaling: ONE 2^24 transform split over the ranks, through kzgb_ntt4_plan (C ABI); parity with the "
                       "single-GPU transform: tests/cpp/multi_gpu_check.cpp, profiles/r02b_mg_rank0.log")
        extra["ntt_2p24_four_step"] = res
        plan.close()

    def base_extra():
        """Rank 0 alone: the SAME 2^24 commit / transform on one GPU -- the denominator of this line's strong scaling."""
        if os.environ.get("KZGB_BENCH_SKIP_BASE") == "1":
            return
        base = {}
        if rank == 0:
            big = _ffi.srs_generate(0, secret, total)
            sc = env["device_scalars"](total, POLY_SEED, 0)
            o2 = torch.zeros(12, dtype=torch.int64, device=dev)
            for _ in range(2):
                big.msm_dev(sc.data_ptr(), total, o2.data_ptr(), 0, stream)
            torch.cuda.synchronize()
            T.e0.record()
            for _ in range(5):
                big.msm_dev(sc.data_ptr(), total, o2.data_ptr(), 0, stream)
            T.e1.record()
            torch.cuda.synchronize()
            ms1 = T.e0.elapsed_time(T.e1) / 5
            same = bool((o2.cpu().numpy().view(np.uint64) == commit0).all())
            big.close()
            for _ in range(2):
                _ffi.ntt_dev(sc.data_ptr(), LOG_C5, False, stream)
            torch.cuda.synchronize()
            T.e0.record()
            for _ in range(5):
                _ffi.ntt_dev(sc.data_ptr(), LOG_C5, False, stream)
            T.e1.record()
            torch.cuda.synchronize()
            base = {"workload": "the same 2^24 commit on ONE GPU (rank 0 alone, same run)", "ms_per_step": ms1, "value": total / (ms1 * 1e-3),
                    "result_equals_sharded": same, "ntt_2p24_ms": T.e0.elapsed_time(T.e1) / 5}
            del sc
        T.barrier()
        extra["strong_scaling_base"] = base

    for fn in (batch_extra, ntt_extra, base_extra):
        try:
            fn()
        except SystemExit:
            raise
        except Exception as ex:   # noqa: BLE001
            extra[fn.__name__ + "_error"] = repr(ex)[:300]
            T.barrier()
    base = extra.get("strong_scaling_base") or {}
    return {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": "C5 coeff-form commit, STRONG scaling: one G1 MSM of 2^24 points per step, points and scalars split evenly over "
                               "%d GPUs (kzgb_msm_g1_sharded_dev: local Pippenger, ncclAllGather of %d x 192 B, combine)" % (world, world),
                   "points_total": total, "points_per_gpu": per, "scalars": "full-width uniform Fr (splitmix64)", "srs": "s^i*G, secret seed 69",
                   "parallelism": "points sharded over %d GPUs" % world,
                   "cache": "%d rotating scalar vectors of %d MiB per rank + the point table: inputs larger than L2" % (NBUF, per * 32 >> 20),
                   "note": "N = 1 of this bench is config C2 (one 2^20 commit per step); the single-GPU time of THIS workload is strong_scaling_base"},
        "strong_scaling_base": base or None,
        "roofline": roofline, "cpu_baseline": None,
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": per * 32 * world, "d2h_bytes_per_step": 96 * world,
                "ms_per_step": e2e_ms, "call": "kzgb_msm_g1_sharded (host pointers, pinned), one call per step on every rank"},
        "gpu_launches": launches, "clocks": clk.summary(), "extra": extra,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kzg_b200", choices=["kzg_b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
