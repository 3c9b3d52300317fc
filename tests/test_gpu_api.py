"""GPU suite: the host mirror of the reference API, written to read like the reference's own tests
(coeff_form.rs:185-399, eval_form.rs:282-484, ft.rs:406-509).  The reference's pairing-based verifier
(out of scope) is replaced by the closed form under the known setup secret."""
import numpy as np
import pytest

import kzg_b200 as kzg
from kzg_b200 import scalar
from oracle import kzg_oracle as O

pytestmark = pytest.mark.gpu
R = O.R_MOD
SECRET = O.fr_stream(69, 1)[0]


def point(k):
    """k*G as blst_p1_affine limbs via the oracle."""
    return np.array(O.g1_affine_to_blst(O.g1_to_affine(O.g1_mul(O.G1_GENERATOR, k % R))), dtype=np.uint64)


def ev(p_ints, x):
    return O.Polynomial.new(p_ints).eval(x)


@pytest.fixture(scope="module")
def params(gpu):
    return kzg.setup(SECRET, 16)


def test_basic(params, golden):
    """coeff_form.rs:271-283 test_basic."""
    prover = kzg.KZGProver(params)
    coeffs = O.fr_stream_u64(42, 12)
    polynomial = kzg.Polynomial.from_ints(coeffs)
    commitment = prover.commit(polynomial)
    assert (commitment == point(ev(coeffs, SECRET))).all()
    other = kzg.Polynomial.from_ints(O.fr_stream_u64(43, 12))
    assert (prover.commit(other) != commitment).any()
    full16 = kzg.Polynomial.from_ints(O.fr_stream(42, 16))
    assert scalar.g1_affine_compress(prover.commit(full16)).hex() == golden["coeff"][0]["commit"]
    with pytest.raises(IndexError):                       # coeff_form.rs:60 slice panic
        prover.commit(kzg.Polynomial.from_ints(list(range(1, 18))))


def test_commit_many_equals_commit_loop(params):
    """commit_many (pipelined batch behind kzgb_msm_g1_batch) == [commit(p) for p in ...], incl. padded degrees."""
    prover = kzg.KZGProver(params)
    polys = [kzg.Polynomial.from_ints(O.fr_stream(50 + i, k)) for i, k in enumerate((16, 1, 7, 16, 3))]
    polys.append(kzg.Polynomial(scalar.fr_vec(O.fr_stream_u64(42, 6) + [0] * 10), 15))
    many = prover.commit_many(polys)
    for p, c in zip(polys, many):
        assert (prover.commit(p) == c).all()
    assert (many[0] == point(ev(O.fr_stream(50, 16), SECRET))).all()
    with pytest.raises(IndexError):
        prover.commit_many([polys[0], kzg.Polynomial.from_ints(list(range(1, 18)))])


def test_create_witness_many_equals_loop(params):
    """create_witness_many (kzgb_witness_coeff_many: one upload, pipelined scans + MSMs) == a loop of create_witness;
    a point off the polynomial raises with its index; constant polynomials follow polynomial.rs:194-199."""
    prover = kzg.KZGProver(params)
    coeffs = O.fr_stream(42, 16)
    poly = kzg.Polynomial.from_ints(coeffs)
    xs = O.fr_stream(7, 7)
    ys = [ev(coeffs, x) for x in xs]
    many = prover.create_witness_many(poly, scalar.fr_vec(xs), scalar.fr_vec(ys))
    for i, (x, y) in enumerate(zip(xs, ys)):
        assert (many[i] == prover.create_witness(poly, (scalar.fr(x), scalar.fr(y)))).all()
        assert (many[i] == point((ev(coeffs, SECRET) - y) * pow(SECRET - x, -1, R))).all()
    ys_bad = list(ys)
    ys_bad[4] = (ys_bad[4] + 1) % R
    with pytest.raises(kzg.PointNotOnPolynomial) as ei:
        prover.create_witness_many(poly, scalar.fr_vec(xs), scalar.fr_vec(ys_bad))
    assert ei.value.index == 4
    const = kzg.Polynomial.from_ints([5])
    w = prover.create_witness_many(const, scalar.fr_vec([3, 9]), scalar.fr_vec([5, 5]))   # psi = 0 -> identity
    assert not w.any()
    with pytest.raises(kzg.PointNotOnPolynomial):
        prover.create_witness_many(const, scalar.fr_vec([3]), scalar.fr_vec([6]))
    assert prover.create_witness_many(poly, scalar.fr_vec([]).reshape(0, 4), scalar.fr_vec([]).reshape(0, 4)).shape == (0, 12)


def test_modify_single_coeff(params):
    """coeff_form.rs:295-311."""
    prover = kzg.KZGProver(params)
    coeffs = O.fr_stream_u64(42, 8)
    c = prover.commit(kzg.Polynomial.from_ints(coeffs))
    coeffs2 = list(coeffs)
    coeffs2[2] = (coeffs2[2] + 1) % R
    c2 = prover.commit(kzg.Polynomial.from_ints(coeffs2))
    assert (c2 == point(ev(coeffs, SECRET) + SECRET ** 2)).all() and (c != c2).any()


def test_padded_polynomial_uses_declared_degree(params):
    """coeff_form.rs:208-219: trailing zero padding then shrink_degree; only coeffs[..degree+1] are read."""
    prover = kzg.KZGProver(params)
    coeffs = O.fr_stream_u64(42, 6)
    padded = kzg.Polynomial(scalar.fr_vec(coeffs + [0] * 10), 15)
    c_stale = prover.commit(padded)                       # stale degree: zeros are still fed to the MSM
    padded.shrink_degree()
    assert padded.degree == 5
    assert (prover.commit(padded) == c_stale).all() and (c_stale == point(ev(coeffs, SECRET))).all()


def test_eval_basic(params, golden):
    """coeff_form.rs:313-342 test_eval_basic."""
    prover = kzg.KZGProver(params)
    coeffs = O.fr_stream_u64(42, 13)
    polynomial = kzg.Polynomial.from_ints(coeffs)
    x = O.fr_stream(7, 1)[0]
    y = ev(coeffs, x)
    assert scalar.fr_int(polynomial.eval(scalar.fr(x))) == y
    w = prover.create_witness(polynomial, (scalar.fr(x), scalar.fr(y)))
    assert (w == point((ev(coeffs, SECRET) - y) * pow(SECRET - x, -1, R))).all()
    with pytest.raises(kzg.PointNotOnPolynomial):
        prover.create_witness(polynomial, (scalar.fr(x), scalar.fr(y + 1)))
    # degree-1 edge case p = X + 3 at (1, 4)  (:332-341): psi has one coefficient -> gs[0] * psi[0]
    w = prover.create_witness(kzg.Polynomial.from_ints([3, 1]), (scalar.fr(1), scalar.fr(4)))
    assert (w == point(1)).all()
    # constant polynomial: zero dividend -> identity; wrong value -> error (Appendix A.2, polynomial.rs:194-199)
    assert not prover.create_witness(kzg.Polynomial.from_ints([9]), (scalar.fr(5), scalar.fr(9))).any()
    with pytest.raises(kzg.PointNotOnPolynomial):
        prover.create_witness(kzg.Polynomial.from_ints([9]), (scalar.fr(5), scalar.fr(8)))
    g = golden["coeff"][0]
    p16 = kzg.Polynomial.from_ints(O.fr_stream(42, 16))
    w = prover.create_witness(p16, (scalar.fr(int.from_bytes(bytes.fromhex(g["x"]), "little")),
                                    scalar.fr(int.from_bytes(bytes.fromhex(g["y"]), "little"))))
    assert scalar.g1_affine_compress(w).hex() == g["witness"]


@pytest.mark.parametrize("n,m", [(15, 8), (16, 4), (9, 1), (8, 8), (3, 5)])
def test_eval_batched(params, golden, n, m):
    """coeff_form.rs:344-398 test_eval_batched / test_eval_batched_all_points (#points == #coeffs, and more)."""
    prover = kzg.KZGProver(params)
    coeffs = O.fr_stream(42, n)
    polynomial = kzg.Polynomial.from_ints(coeffs)
    xs = O.fr_stream(7, m)
    ys = [ev(coeffs, x) for x in xs]
    bw = prover.create_witness_batched(polynomial, scalar.fr_vec(xs), scalar.fr_vec(ys))
    tree = O.SubProductTree.new_from_points(xs)
    ref_r = O.Polynomial.lagrange_interpolation_with_tree(xs, ys, tree)
    assert bw.polynomial().degree == ref_r.degree
    assert scalar.fr_vec_int(bw.polynomial().slice_coeffs()) == ref_r.coeffs[: ref_r.degree + 1]
    z = 1
    for x in xs:
        z = z * (SECRET - x) % R
    assert (bw.elem() == point((ev(coeffs, SECRET) - ref_r.eval(SECRET)) * pow(z, -1, R))).all()
    ys_bad = list(ys)
    ys_bad[-1] = (ys_bad[-1] + 1) % R
    with pytest.raises(kzg.PointNotOnPolynomial):
        prover.create_witness_batched(polynomial, scalar.fr_vec(xs), scalar.fr_vec(ys_bad))
    if (n, m) == (16, 4):
        g = golden["coeff"][0]["batched"]
        assert scalar.g1_affine_compress(bw.elem()).hex() == g["w"]
        assert [scalar.fr_to_bytes_le(c).hex() for c in bw.polynomial().slice_coeffs()] == g["r"]


def test_fft_composition():
    """ft.rs:447-479: ifft(fft(x)) == x for sizes 2^0..2^9, plus fft vs the naive DFT."""
    for lg in range(0, 10):
        v = O.fr_stream(5, 1 << lg)
        dom = kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(v))
        dom.fft()
        if lg <= 6:
            assert scalar.fr_vec_int(dom.coeffs) == O.naive_dft(v, dom.omega)
        dom.ifft()
        assert scalar.fr_vec_int(dom.coeffs) == v


def test_eval_form_prover(golden):
    """eval_form.rs:386-483: commit / create_witness over the Lagrange basis; :318-339 div_by_omega_i."""
    d = 16
    params = kzg.setup(SECRET, d)
    lagrange = kzg.compute_lagrange_basis_known_secret(SECRET, d)
    prover = kzg.KZGProverEvalForm(params, lagrange)
    assert prover.degree() == 16 and prover.omega() == O.compute_omega(16)[2]
    values = O.fr_stream(42, d)
    evals = kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(values))
    commitment = prover.commit(evals)
    coeffs = kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(values))
    coeffs.ifft()
    mono = kzg.KZGProver(params).commit(coeffs.into_polynomial())     # verify_poly path, eval_form.rs:162-171
    assert (commitment == mono).all()
    others = [kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(O.fr_stream(60 + k, d))) for k in range(3)]
    many = prover.commit_many([evals] + others)                      # pipelined batch == commit in a loop
    assert (many[0] == commitment).all() and all((many[k + 1] == prover.commit(e)).all() for k, e in enumerate(others))
    idx = [0, 5, 15, 5, 9]
    wm = prover.create_witness_many(evals, idx)                        # pipelined multi-opening == loop
    assert all((wm[j] == prover.create_witness(evals, i)).all() for j, i in enumerate(idx))
    with pytest.raises(IndexError):
        prover.create_witness_many(evals, [3, 16])
    g = golden["eval"][0]
    assert scalar.g1_affine_compress(commitment).hex() == g["commit"]
    w = prover.create_witness(evals, 3)
    assert scalar.g1_affine_compress(w).hex() == g["witness"]
    # test_div_by_omega_i: ifft(div_by_omega_i(evals - y, 3)) == long_division(p - y, X - w^3)
    q = kzg.div_by_omega_i(evals, 3)
    q.ifft()
    p = coeffs.into_polynomial().sub_constant(evals.coeffs[3])
    quot, rem = p.long_division_linear(scalar.fr(pow(evals.omega, 3, R)))
    assert rem is None and q.into_polynomial() == quot
    assert not prover.create_witness_all().any()
    with pytest.raises(AssertionError):                               # eval_form.rs:115
        prover.commit(kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(values[:8])))


@pytest.mark.parametrize("d", [1, 2, 16, 64, 1024])
def test_compute_lagrange_basis_without_the_secret(gpu, d):
    """eval_form.rs:254-280: the Lagrange SRS derived from params.gs alone (group inverse NTT) equals the closed form
    L_i(s)*G under the known secret, point for point; and an eval-form commitment over it equals the monomial one."""
    params = kzg.setup(SECRET, d)
    lagrange = kzg.compute_lagrange_basis(params)
    if d > 1:
        expect = kzg.compute_lagrange_basis_known_secret(SECRET, d)
        assert (lagrange.read_affine() == expect.read_affine()).all()
    else:
        assert (lagrange.read_affine() == params.gs_affine()).all()     # L_0 = 1
    values = O.fr_stream(42, d)
    evals = kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(values))
    prover = kzg.KZGProverEvalForm(params, lagrange)
    coeffs = kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(values))
    coeffs.ifft()
    assert (prover.commit(evals) == kzg.KZGProver(params).commit(coeffs.into_polynomial())).all()


def test_coset_and_pointwise_helpers(gpu):
    """ft.rs:142-271 on the device vs the oracle restatement, incl. the quotient-on-a-coset pipeline."""
    n = 1 << 10
    a_i, b_i = O.fr_stream(2, n // 2), O.fr_stream(3, n // 2)
    prod = O.Polynomial.new(a_i).fft_mul(O.Polynomial.new(b_i)).coeffs + [0] * (2 * n)
    c_i = [(prod[i] + prod[i + n]) % R for i in range(n)]
    pad = lambda v: scalar.fr_vec(v + [0] * (n - len(v)))
    A, B, C = (kzg.EvaluationDomain.from_coeffs(pad(v)) for v in (a_i, b_i, c_i))
    oa = O.EvaluationDomain.from_coeffs(a_i + [0] * (n - len(a_i)))
    oa.coset_fft()
    for dom in (A, B, C):
        dom.coset_fft()
    assert scalar.fr_vec_int(A.coeffs) == oa.coeffs
    A.mul_assign(B)
    A.sub_assign(C)
    A.divide_by_z_on_coset()
    A.icoset_fft()
    assert scalar.fr_vec_int(A.coeffs) == [prod[i + n] for i in range(n)]
    d = kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(a_i))
    d.distribute_powers(5)
    assert scalar.fr_vec_int(d.coeffs) == [v * pow(5, i, R) % R for i, v in enumerate(a_i)]
    assert d.z(7) == (pow(7, n // 2, R) - 1) % R


@pytest.mark.parametrize("n,k", [(1, 1), (5, 10), (50, 50), (129, 130), (1000, 3000)])
def test_polynomial_mul(gpu, n, k):
    """ft.rs:411-434 polynomial_arith (fft_mul == naive mul) and the best_mul switch (polynomial.rs:185-191)."""
    a_i, b_i = O.fr_stream(1, n), O.fr_stream(2, k)
    a, b = kzg.Polynomial.from_ints(a_i), kzg.Polynomial.from_ints(b_i)
    oa, ob = O.Polynomial.new(a_i), O.Polynomial.new(b_i)
    ref = oa.best_mul(ob)
    got = a.best_mul(b)
    assert got.degree == ref.degree and len(got.coeffs) == len(ref.coeffs)
    assert scalar.fr_vec_int(got.coeffs) == ref.coeffs
    if n <= 129:
        naive = oa.mul_naive(ob)
        assert scalar.fr_vec_int(a.mul(b).coeffs) == naive.coeffs and a.mul(b).degree == naive.degree


def test_verify_poly_and_wire_formats(gpu, params):
    """coeff_form.rs:119-124 / eval_form.rs:162-171 verify_poly; compressed-G1 and LE-Fr encodings vs the oracle."""
    prover, verifier = kzg.KZGProver(params), kzg.KZGVerifier(params)
    coeffs = O.fr_stream(42, 12)
    poly = kzg.Polynomial.from_ints(coeffs)
    c = prover.commit(poly)
    assert verifier.verify_poly(c, poly)
    assert not verifier.verify_poly(c, kzg.Polynomial.from_ints(O.fr_stream(43, 12)))
    evals = kzg.EvaluationDomain.from_coeffs(scalar.fr_vec(O.fr_stream(42, 16)))
    lag = kzg.compute_lagrange_basis(params)
    ce = kzg.KZGProverEvalForm(params, lag).commit(evals)
    assert kzg.KZGVerifierEvalForm(params).verify_poly(ce, evals)
    # SRS export / import through 48-byte compressed points
    blob = params.to_compressed()
    expect = b"".join(O.g1_compress(O.g1_mul(O.G1_GENERATOR, pow(SECRET, i, R))) for i in range(len(params)))
    assert blob == expect
    again = kzg.KZGParams.from_compressed(blob)
    assert (again.gs_affine() == params.gs_affine()).all()
    assert (kzg.KZGProver(again).commit(poly) == c).all()
    # identity, sign bit, and malformed encodings
    pts = np.vstack([params.gs_affine(0, 3), np.zeros((1, 12), dtype=np.uint64)])
    comp = gpu.g1_compress(pts)
    assert comp[-48:] == bytes([0xC0]) + bytes(47)
    assert (gpu.g1_decompress(comp) == pts).all()
    neg = bytearray(comp[:48]); neg[0] ^= 0x20
    xy, nxy = scalar.g1_affine_xy(pts[0]), scalar.g1_affine_xy(gpu.g1_decompress(bytes(neg))[0])
    assert nxy == (xy[0], O.P_MOD - xy[1])
    with pytest.raises(gpu.KZGBackendError):
        gpu.g1_decompress(bytes([0x80]) + bytes(46) + bytes([1]))       # x = 1: 1 + 4 = 5 is not a square mod p?  or on-curve
    with pytest.raises(gpu.KZGBackendError):
        gpu.g1_decompress(bytes([0x9f]) + bytes([0xff]) * 47)             # x >= p
    with pytest.raises(gpu.KZGBackendError):
        gpu.g1_decompress(bytes(48))                                      # uncompressed flag clear
    # Fr canonical little-endian bytes
    vals = [0, 1, R - 1, SECRET]
    b = gpu.fr_to_bytes_le(scalar.fr_vec(vals))
    assert b == b"".join(O.fr_to_bytes_le(v) for v in vals)
    assert (gpu.fr_from_bytes_le(b) == scalar.fr_vec(vals)).all()
    with pytest.raises(gpu.KZGBackendError):
        gpu.fr_from_bytes_le(R.to_bytes(32, "little"))                    # r itself is not canonical


def test_cpp_host_mirror_demo_runs():
    """The header-only C++ mirror (include/kzg_b200.hpp) end to end in its own process: commit, create_witness,
    commit_many, create_witness_many, rejection of a point off the polynomial, kzgb_shutdown."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "host_mirror_demo")
    if not os.path.exists(exe):
        pytest.skip("build/host_mirror_demo not built (make)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "commit_many ok=1 create_witness_many ok=1 bad rejected=1" in r.stdout
