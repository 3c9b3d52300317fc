"""GPU suite, round 2: device-resident polynomial handles, the in-memory layout pinned by DOCUMENTED blst / zkcrypto
constants (not only struct sizes), the multi-GPU C ABI (world 1 through Python; world >= 2 and host threads through the
C++ checks of tests/cpp/), per-call device selection from foreign threads."""
import os
import subprocess
import threading

import numpy as np
import pytest

from kzg_b200 import scalar
from oracle import kzg_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = O.R_MOD

# ---- constants as PUBLISHED by blst (src/consts.c, e2.c) / zkcrypto bls12_381 (g1.rs, scalar.rs): raw in-memory limbs
BLS12_381_Rx_P = [0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745, 0x5c071a97a256ec6d,
                  0x15f65ec3fa80e493]                                        # blst BLS12_381_Rx.p = 2^384 mod p
BLS12_381_G1_X = [0x5cb38790fd530c16, 0x7817fc679976fff5, 0x154f95c7143ba1c1, 0xf0ae6acdf3d0e747, 0xedce6ecc21dbf440,
                  0x120177419e0bfb75]                                        # blst BLS12_381_G1.X (Montgomery)
BLS12_381_G1_Y = [0xbaac93d50ce72271, 0x8c22631a7918fd8e, 0xdd595f13570725ce, 0x51ac582950405194, 0x0e1c8c3fad0059c0,
                  0x0bbc3efc5008a26a]
FR_R = [0x00000001fffffffe, 0x5884b7fa00034802, 0x998c4fefecbc4ff5, 0x1824b159acc5056f]      # Scalar::one() = 2^256 mod r
FR_ROOT_OF_UNITY = [0xb9b58d8c5f0e466a, 0x5b1b4c801819d7ec, 0x0af53ae352a31e64, 0x5bf3adda19e9b27b]   # Scalar::root_of_unity()
ROOT_OF_UNITY_CANONICAL = 0x16a2a19edfe81f20d09b681922c813b4b63683508c2280b93829971f439f0d2b
G1_X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1_Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
G1_COMPRESSED = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"


def _limbs(v, n):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def test_blst_memory_layout_pinned_by_published_constants(gpu):
    """Limb ORDER and Montgomery RADIX of kzgb_fr / kzgb_fp / kzgb_p1 are asserted with the raw limb images that blst
    and zkcrypto publish for known values -- pushed through the device as opaque blst_fr / blst_fp / blst_p1 buffers."""
    u = lambda l: np.array(l, dtype=np.uint64).reshape(1, -1)
    # from_mont(R) = 1, from_mont(root_of_unity limbs) = the canonical 2^32-th root of unity
    assert gpu.test_fr_op(u(FR_R), None, 5)[0].tolist() == [1, 0, 0, 0]
    assert gpu.test_fr_op(u(FR_ROOT_OF_UNITY), None, 5)[0].tolist() == _limbs(ROOT_OF_UNITY_CANONICAL, 4)
    assert gpu.test_fp_op(u(BLS12_381_Rx_P), None, 5)[0].tolist() == [1, 0, 0, 0, 0, 0]
    assert gpu.test_fp_op(u(BLS12_381_G1_X), None, 5)[0].tolist() == _limbs(G1_X, 6)
    assert gpu.test_fp_op(u(BLS12_381_G1_Y), None, 5)[0].tolist() == _limbs(G1_Y, 6)
    # the published root of unity has order exactly 2^32: 31 squarings give -1 = r - 1, one more gives R (= one)
    x = u(FR_ROOT_OF_UNITY)
    for _ in range(31):
        x = gpu.test_fr_op(x, None, 6)
    assert scalar.fr_int(x[0]) == R - 1
    assert gpu.test_fr_op(x, None, 6)[0].tolist() == FR_R
    # the published generator, as a raw blst_p1 {X, Y, Z = Rx}, times Scalar::one() through the MSM: same affine limbs,
    # and the 48 bytes every BLS12-381 library prints for the generator
    p1 = np.array(BLS12_381_G1_X + BLS12_381_G1_Y + BLS12_381_Rx_P, dtype=np.uint64).reshape(1, 18)
    srs = gpu.SRS.from_jacobian(p1)
    aff = srs.msm(u(FR_R))
    assert aff.tolist() == BLS12_381_G1_X + BLS12_381_G1_Y
    assert gpu.g1_compress(aff).hex() == G1_COMPRESSED
    jac = srs.msm_jacobian(u(FR_R))          # blst_p1 out: X/Z^2 must be the published x again
    zi = pow(O.fq_from_mont_limbs(jac[12:18]), -1, O.P_MOD)
    assert O.fq_from_mont_limbs(jac[0:6]) * zi * zi % O.P_MOD == G1_X
    srs.close()


def test_device_resident_polynomial_handles(gpu, cpu):
    """f2: commit -> create_witness -> create_witness_batched on ONE uploaded polynomial == the host-pointer calls."""
    from kzg_b200 import KZGParams, KZGProver, Polynomial
    from kzg_b200.params import PointNotOnPolynomial
    n = 3000
    s = cpu.fr_stream(69, 1)[0]
    params = KZGParams(gpu.srs_generate(0, s, 4096))
    prover = KZGProver(params)
    coeffs = cpu.fr_stream(42, n)
    host = Polynomial.new(coeffs)
    dev = Polynomial.new(coeffs).to_device()
    assert dev.device is not None and (dev.device.download() == coeffs).all()
    assert (prover.commit(dev) == prover.commit(host)).all()
    x = cpu.fr_stream(7, 1)[0]
    y = host.eval(x)
    assert (prover.create_witness(dev, (x, y)) == prover.create_witness(host, (x, y))).all()
    with pytest.raises(PointNotOnPolynomial):
        prover.create_witness(dev, (x, scalar.fr(5)))
    xs = cpu.fr_stream(8, 9)
    ys = np.vstack([host.eval(xi) for xi in xs])
    a, b = prover.create_witness_batched(dev, xs, ys), prover.create_witness_batched(host, xs, ys)
    assert a == b
    ys[3] = scalar.fr(1)
    with pytest.raises(PointNotOnPolynomial):
        prover.create_witness_batched(dev, xs, ys)
    # declared degree below the uploaded length: only coeffs[..degree+1] take part (polynomial.rs:148-154)
    short = Polynomial(coeffs, 99).to_device()
    assert (prover.commit(short) == prover.commit(Polynomial(coeffs[:100], 99))).all()
    dev.release_device()
    assert dev.device is None


def test_device_resident_evaluation_domain(gpu, cpu):
    from kzg_b200 import EvaluationDomain, KZGParams, KZGProverEvalForm
    from kzg_b200.eval_form import compute_lagrange_basis_known_secret
    d = 1 << 10
    s = cpu.fr_stream(69, 1)[0]
    params = KZGParams(gpu.srs_generate(0, s, d))
    prover = KZGProverEvalForm(params, compute_lagrange_basis_known_secret(s, d))
    evals = EvaluationDomain.from_coeffs(cpu.fr_stream(3, d))
    ref_c, ref_w = prover.commit(evals), prover.create_witness(evals, 77)
    evals.to_device()
    assert evals.device is not None
    assert (prover.commit(evals) == ref_c).all() and (prover.create_witness(evals, 77) == ref_w).all()
    # handle-level transform == host transform; a host-side mutation drops the stale resident copy
    h = gpu.DevicePoly(evals.coeffs)
    h.ntt(10, False)
    st, want = gpu.ntt(evals.coeffs, 10, False)
    assert (h.download() == want).all()
    h.close()
    evals.fft()
    assert evals.device is None and (evals.coeffs == want).all()


def test_entry_points_from_foreign_threads(gpu, cpu):
    """Entry points select the library's device per call: results from worker threads equal the main thread's."""
    n = 1 << 12
    srs = gpu.srs_generate(0, cpu.fr_stream(69, 1)[0], n)
    p = cpu.fr_stream(42, n)
    want_c = srs.msm(p)
    want_f = gpu.ntt(p, 12, False)[1]
    got = {}

    def work(k):
        got[k] = (srs.msm(p), gpu.ntt(p, 12, False)[1])

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for k in range(4):
        assert (got[k][0] == want_c).all() and (got[k][1] == want_f).all()
    srs.close()


def test_sharded_entry_points_degenerate_to_single_gpu(gpu, cpu):
    """World of one (no communicator needed): kzgb_msm_g1_sharded* and the four-step plan give the single-GPU bytes."""
    n = 1 << 12
    srs = gpu.srs_generate(0, cpu.fr_stream(69, 1)[0], n)
    p, q = cpu.fr_stream(42, n), cpu.fr_stream(43, n - 7)
    assert (gpu.msm_sharded(srs, p) == srs.msm(p)).all()
    many = gpu.msm_sharded_batch(srs, [p, q, p[:0]])
    assert (many[0] == srs.msm(p)).all() and (many[1] == srs.msm(q)).all() and not many[2].any()
    for log_n in (12, 16):
        a = cpu.fr_stream(5, 1 << log_n)
        plan = gpu.Ntt4Plan(log_n)
        X = plan.run_host(a, False, True)
        assert (X == gpu.ntt(a, log_n, False)[1]).all()
        assert (plan.run_host(X, True, False) == a).all()
        plan.close()
    srs.close()


def _run(exe, *args, timeout=600):
    path = os.path.join(ROOT, "build", exe)
    if not os.path.exists(path):
        pytest.skip(f"build/{exe} not built (make)")
    return subprocess.run([path, *map(str, args)], capture_output=True, text=True, timeout=timeout)


def test_cpp_threads_check():
    """4 host threads x 2 SRS handles through the C ABI (no Python in the loop): bit-equal results."""
    r = _run("threads_check", 14, 4)
    assert r.returncode == 0 and "threads_check: PASS" in r.stdout, r.stdout + r.stderr


def test_cpp_multi_gpu_check(tmp_path):
    """C++ multi-process parity of the sharded MSM and the fused four-step NTT: one process per visible GPU (2 when
    the box has them; the world-of-one run still goes through every entry point)."""
    import torch
    world = 2 if torch.cuda.device_count() >= 2 else 1
    path = os.path.join(ROOT, "build", "multi_gpu_check")
    if not os.path.exists(path):
        pytest.skip("build/multi_gpu_check not built (make)")
    idf = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([path, str(r), str(world), idf, "14", "16"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "multi_gpu_check: PASS" in outs[0]


def test_srs_import_rejects_points_outside_the_subgroup(gpu, cpu):
    """blstrs `from_compressed` semantics on the SRS import path: a point ON the curve but OUTSIDE G1 is refused
    (the cofactor of E(Fq) is ~2^126, so almost every curve point is such a point); the unchecked variant takes it."""
    P = O.P_MOD
    x = 1
    while True:       # first small x with x^3 + 4 a square: that curve point is not in G1 (checked below by [r]P != O)
        rhs = (x ** 3 + 4) % P
        y = pow(rhs, (P + 1) // 4, P)
        if y * y % P == rhs:
            break
        x += 1
    enc = bytearray(x.to_bytes(48, "big"))
    enc[0] |= 0x80 | (0x20 if y > (P - 1) // 2 else 0)
    good = gpu.srs_export_compressed(gpu.srs_generate(0, cpu.fr_stream(69, 1)[0], 4))
    blob = good[:96] + bytes(enc) + good[96:]
    with pytest.raises(gpu.KZGBackendError, match="subgroup"):
        gpu.srs_from_compressed(blob)
    srs = gpu.srs_from_compressed(blob, subgroup_check=False)
    pts = srs.read_affine()
    assert O.fq_from_mont_limbs(pts[2][:6]) == x
    assert gpu.g1_subgroup_check(pts[:2]) and gpu.g1_subgroup_check(pts[3:]) and not gpu.g1_subgroup_check(pts)
    assert gpu.srs_from_compressed(good) is not None            # honest input still loads


def test_interpolant_shape_follows_the_reference_at_the_fft_mul_threshold(gpu, cpu):
    """m >= 258: the reference's top-level best_mul takes the fft_mul path (polynomial.rs:167-191) and r comes back in a
    vector padded to 2^k >= m + 1 with a re-measured degree; below, exactly m coefficients (oracle: tree recursion)."""
    from kzg_b200 import Polynomial
    for m in (257, 258, 300):
        xs, ys = cpu.fr_stream(21, m), cpu.fr_stream(22, m)
        r = Polynomial.lagrange_interpolation(xs, ys)
        tree = O.SubProductTree.new_from_points(scalar.fr_vec_int(xs))
        ref = O.Polynomial.lagrange_interpolation_with_tree(scalar.fr_vec_int(xs), scalar.fr_vec_int(ys), tree)
        assert (r.degree, len(r.coeffs)) == (ref.degree, len(ref.coeffs))
        assert scalar.fr_vec_int(r.coeffs) == [c % R for c in ref.coeffs]


def test_batched_witness_beyond_4096_points(gpu, cpu):
    """The round-1 cap (m <= 4096) is lifted to 16384: m = 4500 openings of a degree-5999 polynomial, checked through
    the closed form under the known secret, w = ((p(s) - r(s)) / Z(s)) * G, and r(x_i) = y_i on a sample."""
    n, m = 6000, 4500
    s = cpu.fr_stream(69, 1)[0]
    srs = gpu.srs_generate(0, s, n)
    p = cpu.fr_stream(42, n)
    xs = cpu.fr_stream(7, m)
    ys = np.vstack([cpu.poly_eval(p, x) for x in xs])
    st, r, w = gpu.witness_batched(srs, p, xs, ys)
    assert st == 0 and len(r) == m
    for i in (0, 1, m // 2, m - 1):
        assert (cpu.poly_eval(r, xs[i]) == ys[i]).all()
    si = scalar.fr_int(s)
    zs = 1
    for x in scalar.fr_vec_int(xs):
        zs = zs * (si - x) % R
    k = (scalar.fr_int(cpu.poly_eval(p, s)) - scalar.fr_int(cpu.poly_eval(r, s))) * pow(zs, -1, R) % R
    assert (w == cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), scalar.fr(k)))[0]).all()
    ys[17] = scalar.fr(3)
    assert gpu.witness_batched(srs, p, xs, ys)[0] == gpu.KZGB_POINT_NOT_ON_POLYNOMIAL
    srs.close()


def test_pageable_host_inputs_take_the_pinned_relay(gpu, cpu):
    """Host-pointer entry points accept ordinary (pageable) caller memory -- what a Rust Vec<Fr> is: large uploads go
    through upload_async's threaded pinned relay.  Ragged sizes (not a multiple of the 2 MiB slice, more slices than ring
    slots) must arrive byte-exact, and a commit from pageable memory must equal the one from pinned memory."""
    import torch
    for n in ((1 << 17) + 3, (1 << 20) + 12345, 5 * (1 << 16) * 13 + 1):
        v = cpu.fr_stream(1000 + (n & 0xff), n)
        h = gpu.DevicePoly(v)
        assert (h.download() == v).all()
        del h
    # host NTT entry point, pageable in and out (both relays): round trip, and == the device-resident transform
    v = cpu.fr_stream(77, 1 << 18)
    st, fwd = gpu.ntt(v, 18)
    h = gpu.DevicePoly(v)
    assert st == 0 and h.ntt(18) == 0 and (h.download() == fwd).all()
    st, back = gpu.ntt(fwd, 18, True)
    assert st == 0 and (back == v).all()
    n = 1 << 19
    s = cpu.fr_stream(69, 1)[0]
    srs = gpu.srs_generate(0, s, n)
    sc = cpu.fr_stream(4242, n)
    pinned = torch.from_numpy(sc.view(np.int64)).pin_memory()
    a = srs.msm(sc)
    b = srs.msm(pinned.numpy().view(np.uint64))
    assert (a == b).all()
    # twice in a row from the same lane: the ring's slots are reused across calls
    sc2 = cpu.fr_stream(4243, n)
    c = srs.msm(sc2)
    d = srs.msm(torch.from_numpy(sc2.view(np.int64)).pin_memory().numpy().view(np.uint64))
    assert (c == d).all() and not (a == c).all()
    # the pipelined batch entry point uploads on its own copy stream: pageable items above the relay threshold, mixed
    # with a short one, must give the single-call results
    items = [sc, sc2[: (1 << 17) + 5], sc[:100], sc2]
    singles = np.stack([srs.msm(x) for x in items])
    assert (srs.msm_batch(items) == singles).all()


@pytest.mark.parametrize("m", [2, 5, 9, 31, 33, 100, 129])
def test_interpolation_with_roots_of_unity_among_the_points(gpu, cpu, m):
    """Z(X) is built from its values on the 2^k-point domain, 2^k > m, and one inverse NTT: points that ARE roots of
    unity of that domain (Z vanishes at some evaluation points), of the next larger one, and 0, must give the same
    interpolant as the reference's tree recursion; m on both sides of powers of two."""
    from kzg_b200 import scalar as S
    N, _, omega = O.compute_omega(m + 1)
    _, _, omega2 = O.compute_omega(2 * N)
    special = [1, 0, omega, pow(omega, N - 1, R), omega2, pow(omega2, 3, R), R - 1]
    xs_i = (special + [x for x in O.fr_stream(7, m) if x not in special])[:m]
    assert len(set(xs_i)) == m
    ys_i = O.fr_stream(8, m)
    r = gpu.interpolate(S.fr_vec(xs_i), S.fr_vec(ys_i))
    tree = O.SubProductTree.new_from_points(xs_i)
    ref = O.Polynomial.lagrange_interpolation_with_tree(xs_i, ys_i, tree)
    assert len(r) == ref.degree + 1 and S.fr_vec_int(r) == ref.coeffs[: ref.degree + 1]


def test_inversion_three_ways(gpu, cpu):
    """inv() (Kaliski's almost-inverse), the Fermat chain (op 7) and the binary GCD on word-size approximations (op 12: an
    experiment kept in field.cuh, measured slower) must agree with each other and with the integers, on random values and
    on edge values: 0, 1, MOD - 1, single bits, values just below the modulus, tiny values."""
    from kzg_b200 import scalar as S
    n = 1024
    edge_r = [0, 1, 2, 3, R - 1, R - 2, (R - 1) // 2, (R + 1) // 2] + [1 << k for k in range(0, 255, 7)] + [R - (1 << k) for k in range(0, 254, 9)]
    a = np.vstack([S.fr_vec(edge_r), cpu.fr_stream(31, n)])
    got = gpu.test_fr_op(a, None, 3)
    assert (got == gpu.test_fr_op(a, None, 12)).all() and (got == gpu.test_fr_op(a, None, 7)).all()
    vals = S.fr_vec_int(a)
    assert S.fr_vec_int(got) == [pow(x, -1, R) if x else 0 for x in vals]
    P = O.P_MOD
    edge_p = [0, 1, 2, P - 1, P - 2, (P - 1) // 2] + [1 << k for k in range(0, 381, 11)] + [P - (1 << k) for k in range(0, 380, 13)]
    rng = np.random.default_rng(77)
    fa = rng.integers(0, 2 ** 63, size=(n, 6), dtype=np.uint64)
    fa[:, 5] &= np.uint64(0x0fffffffffffffff)
    fa = np.vstack([np.array([O.fq_to_mont_limbs(x) for x in edge_p], dtype=np.uint64), fa])
    g = gpu.test_fp_op(fa, None, 3)
    assert (g == gpu.test_fp_op(fa, None, 12)).all() and (g == gpu.test_fp_op(fa, None, 7)).all()
    for k in list(range(len(edge_p))) + [len(edge_p) + 5, len(edge_p) + 77]:
        x = O.fq_from_mont_limbs(fa[k])
        assert O.fq_from_mont_limbs(g[k]) == (pow(x, -1, P) if x else 0)


@pytest.mark.parametrize("case", ["root_coset_blocked", "root_and_first_power_blocked", "free"])
def test_batched_witness_coset_choice(gpu, cpu, case):
    """The exact division by Z runs on a coset s*H that must avoid the opening points.  The first candidate is the
    2N-th root of unity (s^N = -1: every scaling factor is a twiddle-table entry); an x_i with x_i^N = -1 forces the
    fall-back to powers of 7, an additional x_j on 7*H forces 49*H.  The witness and the interpolant must not change."""
    from kzg_b200 import scalar as S
    s_i = O.fr_stream(69, 1)[0]
    n = 256
    srs = gpu.srs_generate(0, S.fr(s_i), n)
    p_i = O.fr_stream(42, n)
    omega = O.compute_omega(n)[2]
    w2n = O.compute_omega(2 * n)[2]
    assert pow(w2n, n, R) == R - 1
    xs_i = O.fr_stream(7, 6)
    if case != "free":
        xs_i = [w2n * pow(omega, 3, R) % R] + xs_i          # on the coset w_2N * H: x^N = -1
    if case == "root_and_first_power_blocked":
        xs_i = [7 * pow(omega, 5, R) % R] + xs_i             # on 7 * H as well
    poly = O.Polynomial.new(p_i)
    ys_i = [poly.eval(x) for x in xs_i]
    st, r, w = gpu.witness_batched(srs, S.fr_vec(p_i), S.fr_vec(xs_i), S.fr_vec(ys_i))
    tree = O.SubProductTree.new_from_points(xs_i)
    ref = O.Polynomial.lagrange_interpolation_with_tree(xs_i, ys_i, tree)
    z = 1
    for x in xs_i:
        z = z * (s_i - x) % R
    k = (poly.eval(s_i) - ref.eval(s_i)) * pow(z, -1, R) % R
    assert st == 0 and S.fr_vec_int(r) == ref.coeffs[: ref.degree + 1]
    assert (w == cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), S.fr(k)))[0]).all()     # closed form ((p - r)/Z)(s) * G
    # a wrong y must still be detected on every coset
    ys_bad = list(ys_i)
    ys_bad[1] = (ys_bad[1] + 1) % R
    assert gpu.witness_batched(srs, S.fr_vec(p_i), S.fr_vec(xs_i), S.fr_vec(ys_bad))[0] == gpu.KZGB_POINT_NOT_ON_POLYNOMIAL


def test_shutdown_releases_handles_and_allows_reinit():
    """kzgb_shutdown with live handles (own process): their device memory is released, a later destroy only frees the
    host side, and a re-initialised library works again (per-device kernel attributes and caches are rebuilt)."""
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from kzg_b200 import _ffi
from oracle import cpu
cpu.build()
_ffi.init(0)
s = cpu.fr_stream(69, 1)[0]
p = cpu.fr_stream(42, 2048)
srs = _ffi.srs_generate(0, s, 2048)
want_c = srs.msm(p)
want_f = _ffi.ntt(p, 11, False)[1]
h = _ffi.DevicePoly(p)
lib = _ffi.load()
assert lib.kzgb_shutdown() == 0
import ctypes
assert lib.kzgb_ntt_fr(ctypes.c_void_p(p.ctypes.data), ctypes.c_uint32(11), ctypes.c_int(0)) < 0   # not initialised: loud failure
assert lib.kzgb_srs_destroy(srs._h) == 0                    # handle outlived the shutdown: host side only
srs._h = None
h._h = None                                                  # (its buffers went with the shutdown's cache / context)
assert lib.kzgb_init(0) == 0
srs2 = _ffi.srs_generate(0, s, 2048)
assert (srs2.msm(p) == want_c).all()
assert (_ffi.ntt(p, 11, False)[1] == want_f).all()
print("REINIT_OK")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "REINIT_OK" in r.stdout, r.stdout + r.stderr
