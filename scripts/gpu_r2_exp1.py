"""Round-2 experiment 1 (one B200): (a) bit-exactness of the Karatsuba / lazy-reduction field products on the device,
(b) chained-multiply throughput of every schedule + the IMAD.WIDE/IADD co-issue probe, (c) MSM 2^20 stage timings,
(d) where the single-call e2e time goes (H2D, poller on/off).  Select the library build with KZGB_LIB."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kzg_b200 import _ffi
from oracle import cpu, kzg_oracle as O

cpu.build()
_ffi.init(0)
tag = os.environ.get("KZGB_LIB", "default").split("/")[-1]
what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")

if what in ("all", "field"):
    n = 1 << 16
    a, b = cpu.fr_stream(11, n), cpu.fr_stream(12, n)
    R, P = O.R_MOD, O.P_MOD
    ok = {}
    ok["fr_kara"] = bool((_ffi.test_fr_op(a, b, 9) == cpu.fr_mul_vec(a, b)).all())
    ab = cpu.fr_mul_vec(a, b)
    apb = _ffi.test_fr_op(a, b, 1)
    ok["fr_lazy10"] = bool((_ffi.test_fr_op(a, b, 10) == _ffi.test_fr_op(ab, cpu.fr_mul_vec(apb, b), 2)).all())
    rng = np.random.default_rng(7)
    fa = rng.integers(0, 2 ** 63, size=(n, 6), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 6), dtype=np.uint64)
    fb = rng.integers(0, 2 ** 63, size=(n, 6), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 6), dtype=np.uint64)
    fa[:, 5] &= np.uint64(0x0fffffffffffffff)
    fb[:, 5] &= np.uint64(0x0fffffffffffffff)
    pm1 = np.array(O.fq_to_mont_limbs(P - 1), dtype=np.uint64)
    fa[0], fb[0], fa[1], fb[1] = pm1, pm1, 0, pm1
    fa[2] = np.uint64(0xffffffffffffffff); fa[2, 5] = np.uint64(0x0fffffffffffffff); fb[2] = fa[2]
    fa[3, 3:] = fa[3, :3]          # equal halves: zero Karatsuba cross difference
    fab = cpu.fp_mul_vec(fa, fb)
    ok["fq_mul"] = bool((_ffi.test_fp_op(fa, fb, 0) == fab).all())
    ok["fq_kara"] = bool((_ffi.test_fp_op(fa, fb, 9) == fab).all())
    ok["fq_sqr"] = bool((_ffi.test_fp_op(fa, None, 6) == cpu.fp_mul_vec(fa, fa)).all())
    fapb = _ffi.test_fp_op(fa, fb, 1)
    ok["fq_lazy10"] = bool((_ffi.test_fp_op(fa, fb, 10) == _ffi.test_fp_op(fab, cpu.fp_mul_vec(fapb, fb), 2)).all())
    faa = cpu.fp_mul_vec(fa, fa)
    ok["fq_lazy11"] = bool((_ffi.test_fp_op(fa, fb, 11) == _ffi.test_fp_op(fab, cpu.fp_mul_vec(fb, faa), 2)).all())
    print(json.dumps({"lib": tag, "bit_exact": ok}), flush=True)

if what in ("all", "micro"):
    names = {2: "IMAD.WIDE", 3: "mad.lo.cc+madc.hi", 6: "IADD", 7: "IMAD.WIDE + IADD 1:1 (counts IMAD.WIDE)", 4: "Fq mul (interleaved Montgomery)",
             13: "Fq mul (Karatsuba + wide reduction)", 15: "Fq a*b-c*d lazy (2 products each)", 16: "Fq sqr (+add) chain",
             5: "Fr mul (interleaved)", 14: "Fr mul (Karatsuba)"}
    for mode, nm in names.items():
        cfgs = ((4, 256), (3, 128), (2, 256), (6, 128)) if mode in (4, 5, 13, 14, 15, 16) else ((8, 256),)
        for bps, thr in cfgs:
            r = max((_ffi.microbench(mode, bps, thr, 32) for _ in range(3)), key=lambda x: x["ops_per_s"])
            print(json.dumps({"lib": tag, "mode": mode, "name": nm, "blocks_per_sm": bps, "threads": thr, "ops_per_s": r["ops_per_s"], "ms": r["ms"]}), flush=True)

if what in ("all", "msm"):
    n = 1 << 20
    s = cpu.fr_stream(69, 1)[0]
    srs = _ffi.srs_generate(0, s, n)
    sc = torch.empty((n, 4), dtype=torch.int64, device=dev)
    _ffi.fr_stream_dev(sc.data_ptr(), 42, n)
    torch.cuda.synchronize()
    coeffs = sc.cpu().numpy().view(np.uint64)
    expect = cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), cpu.poly_eval(coeffs, s)))[0]
    out = torch.zeros(12, dtype=torch.int64, device=dev)
    for _ in range(3):
        srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
    torch.cuda.synchronize()
    parity = bool((out.cpu().numpy().view(np.uint64) == expect).all())
    _ffi.profile_enable(True); _ffi.profile_read()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0)
    e1.record(); torch.cuda.synchronize()
    prof = _ffi.profile_read(); _ffi.profile_enable(False)
    print(json.dumps({"lib": tag, "msm_2p20_ms": e0.elapsed_time(e1) / K, "parity": parity, "c": srs.window_bits,
                      "stage_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items()}}), flush=True)
    # ---- e2e through the host-pointer call: where does the time go? ----
    if what == "all" or os.environ.get("E2E") == "1":
        import ctypes as C
        lib = _ffi.lib()
        hp = sc.cpu().pin_memory()
        hv = hp.numpy().view(np.uint64)
        out_host = np.zeros(12, dtype=np.uint64)
        def one():
            _ffi.check(lib.kzgb_msm_g1(srs._h, C.c_void_p(hp.data_ptr()), C.c_size_t(n), _ffi.ptr(out_host)))
        for _ in range(3):
            one()
        def timed(K=20):
            t0 = time.perf_counter()
            for _ in range(K):
                one()
            return (time.perf_counter() - t0) / K * 1e3
        res = {"lib": tag, "e2e_single_call_ms_no_poller": [round(timed(), 3) for _ in range(3)]}
        # raw H2D of the same pinned buffer, torch stream
        d = torch.empty_like(sc)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            d.copy_(hp, non_blocking=True)
        torch.cuda.synchronize(); res["h2d_32MiB_ms"] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
        # pageable source
        pg = np.array(hv)
        def one_pg():
            _ffi.check(lib.kzgb_msm_g1(srs._h, _ffi.ptr(pg), C.c_size_t(n), _ffi.ptr(out_host)))
        one_pg(); t0 = time.perf_counter()
        for _ in range(5):
            one_pg()
        res["e2e_single_call_ms_pageable"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        for period in (50, 200):
            proc = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", str(period)],
                                    stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            time.sleep(1.0)
            res["e2e_single_call_ms_poller_%dms" % period] = [round(timed(), 3) for _ in range(3)]
            # device-resident calls under the poller, host-synchronised each call
            t0 = time.perf_counter()
            for _ in range(20):
                srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0); torch.cuda.synchronize()
            res["dev_call_sync_ms_poller_%dms" % period] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
            proc.terminate(); proc.wait()
        t0 = time.perf_counter()
        for _ in range(20):
            srs.msm_dev(sc.data_ptr(), n, out.data_ptr(), 0); torch.cuda.synchronize()
        res["dev_call_sync_ms_no_poller"] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
        print(json.dumps(res), flush=True)
    srs.close()
