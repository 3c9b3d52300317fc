"""GPU suite: the CUDA path (through the C ABI) vs the oracle on the same seeded inputs -- bit-exact.

Small sizes: against oracle/ (C restatement, itself pinned to the Python restatement in test_oracle.py) and
the committed golden vectors.  Full BASELINE sizes (2^20): against closed forms under the known SRS secret
(commit = p(s)G, ...), NTT round trips / linearity, and the committed full-size golden commitment."""
import numpy as np
import pytest

from kzg_b200 import scalar
from oracle import kzg_oracle as O

pytestmark = pytest.mark.gpu
R = O.R_MOD


def hexes(arr):
    return [scalar.fr_to_bytes_le(r).hex() for r in np.asarray(arr).reshape(-1, 4)]


def closed_form_point(cpu, k_limbs):
    """k*G through the oracle's C scalar multiplication -> affine limbs."""
    return cpu.g1_to_affine(cpu.g1_mul(cpu.g1_generator(), k_limbs))[0]


# ------------------------------------------------------------------ K0 field arithmetic
def test_field_ops_bit_exact(gpu, cpu):
    n = 2048
    a, b = cpu.fr_stream(1, n), cpu.fr_stream(2, n)
    edge = scalar.fr_vec([0, 1, R - 1, R - 2, 2, (R + 1) // 2, O.ROOT_OF_UNITY, 7])
    a[: len(edge)] = edge
    b[: len(edge)] = edge[::-1]
    assert (gpu.test_fr_op(a, b, 0) == cpu.fr_mul_vec(a, b)).all()
    ai, bi = scalar.fr_vec_int(a[:64]), scalar.fr_vec_int(b[:64])
    assert scalar.fr_vec_int(gpu.test_fr_op(a[:64], b[:64], 1)) == [(x + y) % R for x, y in zip(ai, bi)]
    assert scalar.fr_vec_int(gpu.test_fr_op(a[:64], b[:64], 2)) == [(x - y) % R for x, y in zip(ai, bi)]
    assert scalar.fr_vec_int(gpu.test_fr_op(a[:64], None, 3)) == [pow(x, -1, R) if x else 0 for x in ai]
    assert scalar.fr_vec_int(gpu.test_fr_op(a[:64], None, 6)) == [x * x % R for x in ai]
    # canonical <-> Montgomery
    canon = cpu.fr_to_canonical(a)
    assert (gpu.test_fr_op(a, None, 5) == canon).all() and (gpu.test_fr_op(canon, None, 4) == a).all()
    rng = np.random.default_rng(5)
    fa = rng.integers(0, 2 ** 63, size=(n, 6), dtype=np.uint64)
    fb = rng.integers(0, 2 ** 63, size=(n, 6), dtype=np.uint64)
    fa[:, 5] &= np.uint64(0x0fffffffffffffff)
    fb[:, 5] &= np.uint64(0x0fffffffffffffff)
    pm1 = np.array(O.fq_to_mont_limbs(O.P_MOD - 1), dtype=np.uint64)
    fa[0], fb[0], fa[1], fb[1] = pm1, pm1, 0, pm1
    assert (gpu.test_fp_op(fa, fb, 0) == cpu.fp_mul_vec(fa, fb)).all()
    assert (gpu.test_fp_op(fa, None, 6) == cpu.fp_mul_vec(fa, fa)).all()      # dedicated squaring path
    for k in range(8):
        x, y = O.fq_from_mont_limbs(fa[k]), O.fq_from_mont_limbs(fb[k])
        assert O.fq_from_mont_limbs(gpu.test_fp_op(fa[k:k + 1], fb[k:k + 1], 1)[0]) == (x + y) % O.P_MOD
        assert O.fq_from_mont_limbs(gpu.test_fp_op(fa[k:k + 1], fb[k:k + 1], 2)[0]) == (x - y) % O.P_MOD
    xi = O.fq_from_mont_limbs(fa[3])
    assert O.fq_from_mont_limbs(gpu.test_fp_op(fa[3:4], None, 3)[0]) == pow(xi, -1, O.P_MOD)


# ------------------------------------------------------------------ SRS
def test_srs_upload_and_generators(gpu, cpu, golden):
    s = cpu.fr_stream(69, 1)[0]
    n = 300
    gs = cpu.setup_g1(s, n)                      # lib.rs:38-46 literally
    srs = gpu.SRS.from_jacobian(gs)
    aff = cpu.g1_to_affine(gs)
    assert (srs.read_affine() == aff).all()
    gen = gpu.srs_generate(0, s, n)
    assert (gen.read_affine() == aff).all()      # device generator == sequential setup
    for i, hx in golden["constants"]["srs_points"].items():
        if int(i) < n:
            assert scalar.g1_affine_compress(gen.read_affine(int(i), 1)[0]).hex() == hx
    # sharded generation: points [first, first+count) of a larger SRS
    part = gpu.srs_generate(0, s, n, first=100, count=50)
    assert (part.read_affine() == aff[100:150]).all()
    # identity points survive (blst_p1 with z == 0)
    gs2 = gs[:8].copy()
    gs2[3] = 0
    assert not gpu.SRS.from_jacobian(gs2).read_affine(3, 1).any()
    # Lagrange SRS against the closed form
    d = 32
    lag = gpu.srs_generate(1, s, d)
    sc = O.lagrange_scalars_known_secret(O.fr_stream(69, 1)[0], d)
    for i in (0, 1, 17, 31):
        assert (lag.read_affine(i, 1)[0] == closed_form_point(cpu, scalar.fr(sc[i]))).all()


# ------------------------------------------------------------------ K2-K6 MSM
@pytest.mark.parametrize("n,window,mode", [(1, 0, 0), (2, 0, 0), (33, 0, 0), (1000, 0, 0), (1024, 8, 0), (1024, 13, 0),
                                           (4096, 0, 0), (4096, 16, 0), (4096, 0, 1), (4096, 11, 1)])
def test_msm_matches_oracle(gpu, cpu, n, window, mode):
    s = cpu.fr_stream(69, 1)[0]
    srs = gpu.srs_generate(0, s, n, window_bits=window)
    gs_aff = srs.read_affine()
    gs = np.zeros((n, 18), dtype=np.uint64)
    gs[:, :12] = gs_aff
    gs[:, 12:] = np.array(O.fq_to_mont_limbs(1), dtype=np.uint64)
    sc = cpu.fr_stream(42, n, mode=mode)
    expect = cpu.g1_to_affine(cpu.msm_g1(gs, sc, 8))[0]
    assert (srs.msm(sc) == expect).all()
    assert (cpu.g1_to_affine(srs.msm_jacobian(sc))[0] == expect).all()
    m = max(1, n // 3)  # multi_exp over a prefix: gs[..num_coeffs] (coeff_form.rs:60)
    assert (srs.msm(sc[:m]) == cpu.g1_to_affine(cpu.msm_g1(gs[:m], sc[:m], 8))[0]).all()


def test_msm_edge_cases(gpu, cpu):
    s = cpu.fr_stream(69, 1)[0]
    n = 64
    gs = cpu.setup_g1(s, n)
    srs = gpu.SRS.from_jacobian(gs)
    assert not srs.msm(np.zeros((n, 4), dtype=np.uint64)).any()          # all-zero scalars -> identity
    one = np.repeat(scalar.fr(1).reshape(1, 4), n, axis=0)
    assert (srs.msm(one) == cpu.g1_to_affine(cpu.msm_g1(gs, one, 1))[0]).all()
    big = np.repeat(scalar.fr(R - 1).reshape(1, 4), n, axis=0)            # r-1: every signed digit carries
    assert (srs.msm(big) == cpu.g1_to_affine(cpu.msm_g1(gs, big, 1))[0]).all()
    same = np.repeat(gs[:1], n, axis=0)                                  # identical points: doubling branch
    srs2 = gpu.SRS.from_jacobian(same)
    sc = cpu.fr_stream(3, n)
    assert (srs2.msm(sc) == cpu.g1_to_affine(cpu.msm_g1(same, sc, 1))[0]).all()
    assert (srs2.msm(one) == cpu.g1_to_affine(cpu.msm_g1(same, one, 1))[0]).all()
    # P and -P in one bucket cancel
    pair = np.vstack([gs[1:2], gs[1:2]])
    negone = np.vstack([scalar.fr(5).reshape(1, 4), scalar.fr(R - 5).reshape(1, 4)])
    assert not gpu.SRS.from_jacobian(pair).msm(negone).any()
    with pytest.raises(gpu.KZGBackendError):
        srs.msm(np.zeros((n + 1, 4), dtype=np.uint64))                   # more scalars than points
    assert not srs.msm(np.zeros((0, 4), dtype=np.uint64)).any()          # empty MSM -> identity
    one_pt = gpu.SRS.from_jacobian(gs[:1])                               # single-point SRS: gs[0] * k
    k = cpu.fr_stream(11, 1)
    assert (one_pt.msm(k) == cpu.g1_to_affine(cpu.g1_mul(gs[0], k[0]))[0]).all()


def test_msm_batch_pipelined_matches_single_calls(gpu, cpu):
    """kzgb_msm_g1_batch / _batch_dev (two work areas, tails overlapped with the next accumulation) give exactly the
    results of separate calls: mixed lengths incl. empty and full, more items than pipeline slots, repeated calls."""
    import torch
    s = cpu.fr_stream(69, 1)[0]
    n = 5000
    srs = gpu.srs_generate(0, s, n)
    lens = [n, 1, 0, 4097, n - 1, 33, n, 2]
    polys = [cpu.fr_stream(100 + i, max(m, 1))[:m] for i, m in enumerate(lens)]
    singles = np.stack([srs.msm(p) for p in polys])
    for _ in range(2):
        assert (srs.msm_batch(polys) == singles).all()
    assert srs.msm_batch([]).shape == (0, 12)
    assert (srs.msm_batch(polys[:1]) == singles[:1]).all()
    # device-resident variant, affine and raw XYZZ outputs, on torch's current stream
    dts = [torch.from_numpy(np.ascontiguousarray(p).view(np.int64)).cuda() if len(p) else torch.zeros((0, 4), dtype=torch.int64, device="cuda") for p in polys]
    out = torch.zeros((len(polys), 12), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    srs.msm_batch_dev([t.data_ptr() for t in dts], lens, out.data_ptr(), 0, st)
    srs.msm_batch_dev([t.data_ptr() for t in dts], lens, out.data_ptr(), 0, st)   # back to back: slots are reused safely
    torch.cuda.synchronize()
    assert (out.cpu().numpy().view(np.uint64) == singles).all()
    xyzz = torch.zeros((len(polys), 24), dtype=torch.int64, device="cuda")
    srs.msm_batch_dev([t.data_ptr() for t in dts], lens, xyzz.data_ptr(), 2, st)
    aff = torch.zeros(12, dtype=torch.int64, device="cuda")
    for i in range(len(polys)):
        gpu.g1_combine_dev(xyzz[i].data_ptr(), 1, aff.data_ptr(), 0, st)
        torch.cuda.synchronize()
        assert (aff.cpu().numpy().view(np.uint64) == singles[i]).all()
    # batched multi-GPU combine: pretend 3 ranks gathered [3][K][24] partials -> out[i] = sum over ranks
    K = len(polys)
    gathered = torch.stack([xyzz, xyzz.roll(1, 0), xyzz.roll(2, 0)]).contiguous()
    outs = torch.zeros((K, 12), dtype=torch.int64, device="cuda")
    gpu.g1_combine_batch_dev(gathered.data_ptr(), 3, K, outs.data_ptr(), 0, st)
    torch.cuda.synchronize()
    J = lambda a: np.concatenate([a, np.array(O.fq_to_mont_limbs(1), dtype=np.uint64)]) if a.any() else np.zeros(18, dtype=np.uint64)
    for i in range(K):
        e = cpu.g1_add(cpu.g1_add(J(singles[i]), J(singles[(i - 1) % K])), J(singles[(i - 2) % K]))
        assert (outs[i].cpu().numpy().view(np.uint64) == cpu.g1_to_affine(e)[0]).all()
    # host scalars in, device results out (the multi-GPU commit path from host buffers)
    xy2 = torch.zeros((K, 24), dtype=torch.int64, device="cuda")
    srs.msm_batch_h2d(polys, xy2.data_ptr(), 2, st)
    torch.cuda.synchronize()
    assert torch.equal(xy2, xyzz)
    from kzg_b200.distributed import sharded_msm_many as smm
    o3 = smm(srs, [p for i, p in enumerate(polys) if i != 2], device=torch.device("cuda"))
    torch.cuda.synchronize()
    assert (o3.cpu().numpy().view(np.uint64) == np.delete(singles, 2, axis=0)).all()
    with pytest.raises(gpu.KZGBackendError):
        srs.msm_batch([np.zeros((n + 1, 4), dtype=np.uint64)])
    # host mirror
    from kzg_b200 import KZGParams, KZGProver, Polynomial
    from kzg_b200.distributed import sharded_msm_many
    outs = sharded_msm_many(srs, dts[:2] + dts[3:])
    torch.cuda.synchronize()
    assert (outs.cpu().numpy().view(np.uint64) == np.delete(singles, 2, axis=0)).all()


def test_msm_published_known_answers(gpu, cpu, published_kats):
    """The CUDA MSM reproduces k*G answers published by independent implementations (blst / py_ecc; see
    tests/golden/published_kats.json): one-point commits [k]*G and a three-term sum, compared as compressed bytes."""
    G = cpu.g1_generator().reshape(1, 18)
    one_pt = gpu.SRS.from_jacobian(G)
    for k, h in published_kats["small_multiples"].items():
        assert scalar.g1_affine_compress(one_pt.msm(scalar.fr(int(k)).reshape(1, 4))).hex() == h
    for sk, h in published_kats["secret_to_public"]:
        assert scalar.g1_affine_compress(one_pt.msm(scalar.fr(int(sk, 16)).reshape(1, 4))).hex() == h
    # k1*G + k2*G + k3*G = (k1+k2+k3)*G: three copies of G as the SRS, the published keys as scalars
    sks = [int(sk, 16) for sk, _ in published_kats["secret_to_public"]]
    srs3 = gpu.SRS.from_jacobian(np.repeat(G, 3, axis=0))
    got = srs3.msm(scalar.fr_vec(sks))
    assert scalar.g1_affine_compress(got) == O.g1_compress(O.g1_mul(O.G1_GENERATOR, sum(sks) % R))
    # compressed import/export of the published encodings round-trips through the device codec
    enc = bytes.fromhex("".join(h for _, h in published_kats["secret_to_public"]))
    aff = gpu.g1_decompress(enc)
    assert gpu.g1_compress(aff) == enc
    for i, sk in enumerate(sks):
        assert (aff[i] == cpu.g1_to_affine(cpu.g1_mul(G[0], scalar.fr(sk)))[0]).all()


# ------------------------------------------------------------------ K7 NTT
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 9, 10, 11, 12, 13, 15, 16, 17, 19, 21])
def test_ntt_matches_oracle(gpu, cpu, log_n):
    a = cpu.fr_stream(5, 1 << log_n)
    st, f = gpu.ntt(a, log_n, False)
    assert st == 0 and (f == cpu.fft(a, log_n, False, 8)).all()
    st, fi = gpu.ntt(a, log_n, True)
    assert (fi == cpu.fft(a, log_n, True, 8)).all()
    assert (gpu.ntt(f, log_n, True)[1] == a).all()


def test_ntt_golden_and_limits(gpu, cpu, golden):
    for case in golden["ntt"]:
        a = cpu.fr_stream(5, 1 << case["log_n"])
        assert hexes(gpu.ntt(a, case["log_n"], False)[1]) == case["fft"]
        assert hexes(gpu.ntt(a, case["log_n"], True)[1]) == case["ifft"]
    st, _ = gpu.ntt(np.zeros((1, 4), dtype=np.uint64), 32, False)
    assert st == gpu.KZGB_DEGREE_TOO_LARGE  # ft.rs:67-69


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_full_size_properties(gpu, cpu, log_n):
    """BASELINE sizes: round trip, linearity and two spot values against the definition."""
    n = 1 << log_n
    a, b = cpu.fr_stream(5, n), cpu.fr_stream(6, n)
    fa = gpu.ntt(a, log_n, False)[1]
    assert (gpu.ntt(fa, log_n, True)[1] == a).all()
    if log_n == 20:
        assert (fa == cpu.fft(a, log_n, False, 16)).all()
    fb = gpu.ntt(b, log_n, False)[1]
    fab = gpu.ntt(gpu.test_fr_op(a, b, 1), log_n, False)[1]
    assert (fab == gpu.test_fr_op(fa, fb, 1)).all()
    omega = O.compute_omega(n)[2]
    assert (fa[0] == cpu.poly_eval(a, scalar.fr(1))).all()              # out[0] = sum a_j
    k = 12345
    assert (fa[k] == cpu.poly_eval(a, scalar.fr(pow(omega, k, R)))).all()  # out[k] = a(omega^k)


# ------------------------------------------------------------------ K9 linear quotient
@pytest.mark.parametrize("n", [1, 2, 3, 16, 17, 2047, 2048, 2049, 4095, 4096, 4097, 6145, 65541])
def test_divide_linear_matches_long_division(gpu, cpu, n):
    p = cpu.fr_stream(42, n)
    x = cpu.fr_stream(7, 1)[0]
    y = cpu.poly_eval(p, x)
    q, rem = gpu.divide_linear(p, x, y)
    assert not rem.any()
    assert (gpu.divide_linear(p, x, None)[1] == y).all()
    if n > 1:
        pp = p.copy()
        pp[0] = scalar.fr_sub(p[0], y)
        div = np.vstack([scalar.fr_neg(x), scalar.fr(1)])
        st, qq, _ = cpu.long_division(pp, n - 1, div, 1)          # polynomial.rs:193-227 restated
        assert st == 0 and (qq == q).all()


def test_divide_linear_zero_runs(gpu, cpu):
    """Zero runs trigger the reference's shrink_degree skip / early exit (polynomial.rs:209,218)."""
    n = 5000
    p = cpu.fr_stream(42, n)
    p[100:3000] = 0
    p[4000:] = 0      # stale degree: leading zeros
    x = cpu.fr_stream(7, 1)[0]
    q, rem = gpu.divide_linear(p, x, None)
    div = np.vstack([scalar.fr_neg(x), scalar.fr(1)])
    st, qq, rr = cpu.long_division(p, n - 1, div, 1)
    assert (qq == q).all() and (rr[0] == rem).all()
    # exactly divisible: p = (X - x) * g  -> early exit path in the reference
    g = cpu.fr_stream(3, 64)
    prod = np.zeros((65, 4), dtype=np.uint64)
    gi, xi = scalar.fr_vec_int(g), scalar.fr_int(x)
    prod_i = [0] * 65
    for k, c in enumerate(gi):
        prod_i[k + 1] = (prod_i[k + 1] + c) % R
        prod_i[k] = (prod_i[k] - xi * c) % R
    q, rem = gpu.divide_linear(scalar.fr_vec(prod_i), x, None)
    assert not rem.any() and (q == g).all()


# ------------------------------------------------------------------ K12 eval-form division
@pytest.mark.parametrize("log_d,i", [(1, 0), (1, 1), (4, 3), (5, 0), (10, 777), (13, 8191)])
def test_div_by_omega_i(gpu, cpu, golden, log_d, i):
    d = 1 << log_d
    ev = cpu.fr_stream(42, d)
    got = gpu.div_by_omega_i(ev, log_d, i)
    if d <= 64:
        num = gpu.test_fr_op(ev, np.repeat(ev[i:i + 1], d, axis=0), 2)   # evals - evals[i]
        assert (got == cpu.div_by_omega_i(num, log_d, i)).all()          # eval_form.rs:58-84 literally
    evi = scalar.fr_vec_int(ev)
    omega = O.compute_omega(d)[2]
    num_i = [(c - evi[i]) % R for c in evi]
    assert scalar.fr_vec_int(got) == O.div_by_omega_i_fast(num_i, omega, i)
    for case in golden["eval"]:
        if case["log_d"] == log_d and case["i"] == i:
            assert hexes(got) == case["div_by_omega_i"]


# ------------------------------------------------------------------ K11 interpolation
@pytest.mark.parametrize("m", [1, 2, 3, 7, 8, 64, 256])
def test_interpolation_matches_oracle(gpu, cpu, m):
    xs_i, ys_i = O.fr_stream(7, m), O.fr_stream(8, m)
    r = gpu.interpolate(scalar.fr_vec(xs_i), scalar.fr_vec(ys_i))
    tree = O.SubProductTree.new_from_points(xs_i)
    ref = O.Polynomial.lagrange_interpolation_with_tree(xs_i, ys_i, tree)
    assert len(r) == ref.degree + 1 and scalar.fr_vec_int(r) == ref.coeffs[: ref.degree + 1]
    if m == 1:
        assert scalar.fr_vec_int(r) == [(ys_i[0] - xs_i[0]) % R, 1]     # X + (y - x): polynomial.rs:244-247


def test_interpolation_rejects_repeated_x(gpu):
    xs = scalar.fr_vec([5, 6, 5])
    with pytest.raises(gpu.KZGBackendError):
        gpu.interpolate(xs, scalar.fr_vec([1, 2, 3]))


# ------------------------------------------------------------------ full size: BASELINE configs C2-C4
@pytest.fixture(scope="module")
def full(gpu, cpu):
    n = 1 << 20
    s = cpu.fr_stream(69, 1)[0]
    srs = gpu.srs_generate(0, s, n)
    p = cpu.fr_stream(42, n)
    return {"n": n, "s": s, "srs": srs, "p": p, "ps": cpu.poly_eval(p, s)}


def test_c2_commit_and_witness_full_size(gpu, cpu, golden, full):
    srs, p, s = full["srs"], full["p"], full["s"]
    c = srs.msm(p)
    assert (c == closed_form_point(cpu, full["ps"])).all()                          # commit = p(s) G
    assert scalar.g1_affine_compress(c).hex() == golden["coeff_full"]["commit"]
    x = cpu.fr_stream(7, 1)[0]
    y = cpu.poly_eval(p, x)
    st, w = gpu.witness_coeff(srs, p, x, y)
    k = (scalar.fr_int(full["ps"]) - scalar.fr_int(y)) * pow(scalar.fr_int(s) - scalar.fr_int(x), -1, R) % R
    assert st == 0 and (w == closed_form_point(cpu, scalar.fr(k))).all()          # ((p(s)-y)/(s-x)) G
    assert scalar.g1_affine_compress(w).hex() == golden["coeff_full"]["witness"]
    st, _ = gpu.witness_coeff(srs, p, x, scalar.fr(scalar.fr_int(y) + 1))
    assert st == gpu.KZGB_POINT_NOT_ON_POLYNOMIAL
    # reference-style u64 scalars (benches/commit_coeff_form.rs:16-20)
    p64 = cpu.fr_stream(42, full["n"], mode=1)
    assert (srs.msm(p64) == closed_form_point(cpu, cpu.poly_eval(p64, s))).all()


def test_c3_eval_form_full_size(gpu, cpu, full):
    n, s = full["n"], full["s"]
    lag = gpu.srs_generate(1, s, n)
    evals = cpu.fr_stream(42, n)
    c_lag = lag.msm(evals)
    coeffs = gpu.ntt(evals, 20, True)[1]
    c_mono = full["srs"].msm(coeffs)                   # eval_form.rs:162-171: iNTT + monomial MSM
    assert (c_lag == c_mono).all()
    assert (c_lag == closed_form_point(cpu, cpu.poly_eval(coeffs, s))).all()
    i = 424242
    w = gpu.witness_eval(lag, evals, 20, i)
    omega_i = pow(O.compute_omega(n)[2], i, R)
    ps = scalar.fr_int(cpu.poly_eval(coeffs, s))
    k = (ps - scalar.fr_int(evals[i])) * pow(scalar.fr_int(s) - omega_i, -1, R) % R
    assert (w == closed_form_point(cpu, scalar.fr(k))).all()


def test_c4_batched_witness_full_size(gpu, cpu, full):
    srs, p, s = full["srs"], full["p"], full["s"]
    m = 256
    xs = cpu.fr_stream(7, m)
    # y_i = p(x_i): remainders of linear scans on the device, spot-checked against the oracle's Horner
    ys = np.vstack([gpu.divide_linear(p, xs[i], None)[1] for i in range(m)])
    assert (ys[5] == cpu.poly_eval(p, xs[5])).all()
    st, r, w = gpu.witness_batched(srs, p, xs, ys)
    assert st == 0 and len(r) == m
    tree = O.SubProductTree.new_from_points(scalar.fr_vec_int(xs))
    ref = O.Polynomial.lagrange_interpolation_with_tree(scalar.fr_vec_int(xs), scalar.fr_vec_int(ys), tree)
    assert scalar.fr_vec_int(r) == ref.coeffs[:m] and ref.degree == m - 1
    si = scalar.fr_int(s)
    z = 1
    for x in scalar.fr_vec_int(xs):
        z = z * (si - x) % R
    k = (scalar.fr_int(full["ps"]) - ref.eval(si)) * pow(z, -1, R) % R
    assert (w == closed_form_point(cpu, scalar.fr(k))).all()
    ys_bad = ys.copy()
    ys_bad[17] = scalar.fr(scalar.fr_int(ys[17]) + 1)
    st, _, _ = gpu.witness_batched(srs, p, xs, ys_bad)
    assert st == gpu.KZGB_POINT_NOT_ON_POLYNOMIAL


# ------------------------------------------------------------------ K8 four-step (multi-GPU) building blocks
@pytest.mark.parametrize("log_n,world,inverse", [(8, 2, False), (12, 4, False), (12, 4, True), (16, 8, False), (20, 2, True)])
def test_four_step_blocks_emulated_ranks(gpu, cpu, log_n, world, inverse):
    """Runs every rank's local CUDA steps on this one GPU and performs the all-to-all by hand; the assembled result
    must equal the single-GPU transform (and the oracle)."""
    import torch
    from kzg_b200 import distributed as kd
    vec = cpu.fr_stream(5, 1 << log_n)
    l1, l2, lc = kd.four_step_shape(log_n, world)
    n1, c = 1 << l1, 1 << lc
    ops = kd.CudaNttOps()
    slabs = [torch.from_numpy(s.view(np.int64)).cuda() for s in kd.four_step_scatter(vec, log_n, world)]
    for g in range(world):
        ops.columns(slabs[g], log_n, l1, lc, g * c, inverse)
    torch.cuda.synchronize()
    blocks = []
    for h in range(world):  # rank h receives row block h of every rank's slab
        recv = torch.stack([slabs[g].view(world, n1 // world, c, 4)[h] for g in range(world)]).contiguous()
        rows = ops.unpack(recv, world, n1 // world, c)
        scale = scalar.fr(pow(1 << log_n, -1, R)) if inverse else None
        ops.rows(rows, l2, inverse, scale)
        torch.cuda.synchronize()
        blocks.append(rows.cpu().numpy().view(np.uint64))
    full = kd.four_step_gather(blocks, log_n, world)
    assert (full == gpu.ntt(vec, log_n, inverse)[1]).all()
    if log_n <= 16:
        assert (full == cpu.fft(vec, log_n, inverse, 8)).all()


@pytest.mark.parametrize("log_n,world,inverse", [(8, 2, False), (12, 4, True), (16, 8, False), (20, 2, True), (20, 16, False)])
def test_four_step_fused_scatter_emulated_ranks(gpu, cpu, log_n, world, inverse):
    """The fused column pass (epilogue stores straight into the owners' row matrices, here all on this GPU: the
    peer pointers are local) + the row pass == the single-GPU transform; the slab is left untouched.  The row
    matrices go through kzgb_peer_alloc and an IPC handle is produced for each (opened only across processes)."""
    import torch
    from kzg_b200 import distributed as kd
    vec = cpu.fr_stream(5, 1 << log_n)
    l1, l2, lc = kd.four_step_shape(log_n, world)
    n1, c = 1 << l1, 1 << lc
    nrows = n1 // world
    pops = kd.CudaPeerOps()
    nbytes = nrows * (1 << l2) * 32
    allocs = [pops.peer_alloc(nbytes) for _ in range(world)]
    assert all(len(h) == 64 for _, h in allocs)
    peers = [p for p, _ in allocs]
    host_slabs = kd.four_step_scatter(vec, log_n, world)
    slabs = [torch.from_numpy(s.view(np.int64)).cuda() for s in host_slabs]
    for g in range(world):
        pops.columns_scatter(slabs[g], peers, g, log_n, l1, lc, g * c, inverse)
    torch.cuda.synchronize()
    blocks = []
    scale = scalar.fr(pow(1 << log_n, -1, R)) if inverse else None
    for h in range(world):
        rows = pops.wrap(peers[h], nbytes).view(nrows, 1 << l2, 4)
        pops.rows(rows, l2, inverse, scale)
        torch.cuda.synchronize()
        blocks.append(rows.cpu().numpy().view(np.uint64))
    for g in range(world):
        assert (slabs[g].cpu().numpy().view(np.uint64) == host_slabs[g]).all()
    full = kd.four_step_gather(blocks, log_n, world)
    assert (full == gpu.ntt(vec, log_n, inverse)[1]).all()
    del rows
    for p in peers:
        pops.peer_free(p)


def test_batched_witness_more_points_than_coefficients(gpu, cpu):
    """coeff_form.rs:377-398 generalised: m >= n openings of a degree-(n-1) polynomial -> psi = 0, identity witness;
    a wrong value is still rejected (numerator of degree < m must be the zero polynomial)."""
    s = cpu.fr_stream(69, 1)[0]
    srs = gpu.srs_generate(0, s, 64)
    for n, m in ((40, 40), (17, 50), (1, 3)):
        p = cpu.fr_stream(42, n)
        xs = cpu.fr_stream(7, m)
        ys = np.vstack([cpu.poly_eval(p, xs[i]) for i in range(m)])
        st, r, w = gpu.witness_batched(srs, p, xs, ys)
        assert st == 0 and not w.any() and len(r) == m
        ys[m // 2] = scalar.fr(scalar.fr_int(ys[m // 2]) + 1)
        assert gpu.witness_batched(srs, p, xs, ys)[0] == gpu.KZGB_POINT_NOT_ON_POLYNOMIAL


def test_batched_witness_stale_degree_and_adversarial_coset_point(gpu, cpu):
    """x_i on the first coset 7*H forces the device to pick another shift; the result must not change."""
    s_i = O.fr_stream(69, 1)[0]
    n = 256
    srs = gpu.srs_generate(0, scalar.fr(s_i), n)
    p_i = O.fr_stream(42, n)
    omega = O.compute_omega(n)[2]
    xs_i = [7 * pow(omega, 5, R) % R, 49 * pow(omega, 9, R) % R] + O.fr_stream(7, 6)   # two points on cosets 7H, 49H
    poly = O.Polynomial.new(p_i)
    ys_i = [poly.eval(x) for x in xs_i]
    st, r, w = gpu.witness_batched(srs, scalar.fr_vec(p_i), scalar.fr_vec(xs_i), scalar.fr_vec(ys_i))
    tree = O.SubProductTree.new_from_points(xs_i)
    ref = O.Polynomial.lagrange_interpolation_with_tree(xs_i, ys_i, tree)
    z = 1
    for x in xs_i:
        z = z * (s_i - x) % R
    k = (poly.eval(s_i) - ref.eval(s_i)) * pow(z, -1, R) % R
    assert st == 0 and scalar.fr_vec_int(r) == ref.coeffs[: ref.degree + 1]
    assert (w == closed_form_point(cpu, scalar.fr(k))).all()
