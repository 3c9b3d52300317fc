"""CPU oracle for the KZG prover hot path (TEST INFRASTRUCTURE -- NOT PRODUCT CODE).

Pure-Python big-integer restatement of the algorithms on the hot path of
proxima-one/kzg (reference files cited per function as `file:line`, relative
to the reference's root).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module.
The product path (`kzg_b200/`) never does.

PARITY STATUS: the reference itself cannot be run here, so parity is **unpinned against
the reference's own outputs**; the blst-level arithmetic IS pinned by published known
answers of the blst family (below).  The reference's field
and curve arithmetic (including `G1Projective::multi_exp`) lives in the
un-vendored git dependency `blstrs` (proxima-one fork, rev b98fc83,
Cargo.toml:27) which wraps `blst`; neither is under /root/reference and there
is no Rust toolchain here, so the reference cannot be executed.  The reference
holds no golden byte vectors for commitments / witnesses / NTT outputs.  What
pins this oracle instead (see `tests/test_oracle.py`):
  * the reference's own literal known-answers (polynomial.rs:494-597 division
    and evaluation KATs, polynomial.rs:666-690 interpolation property),
  * published BLS12-381 constants (generator compression 97f1d3a7...c6bb,
    2*G = a572cbea...0f4e, root of unity of order 2^32, Montgomery constants),
  * published known answers produced by independent implementations (blst / py_ecc,
    Ethereum consensus-spec BLS vectors; tests/golden/published_kats.json): compressed
    k*G for k = 1..5 and three full-width secret-key -> public-key pairs, reproduced by
    this file, by oracle/kzg_cpu.c and by the CUDA MSM,
  * closed forms under a known SRS secret: commit(p) = p(s)*G,
    witness = ((p(s)-y)/(s-x))*G, batched w = ((p(s)-r(s))/Z(s))*G,
  * differential structure mirrored from the reference's tests (NTT vs naive
    DFT, ifft(fft(x)) = x, div_by_omega_i vs coefficient-form long division,
    Lagrange-basis MSM vs iNTT + monomial MSM).
Every output on the path is a mathematically unique group element / field
vector, so "bit-exact with the reference" is compared as compressed G1
(48 bytes, ZCash encoding) and canonical little-endian Fr (32 bytes).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

# ---------------------------------------------------------------------------
# BLS12-381 constants (blstrs `Scalar` / `G1Affine`; ft.rs:67,73,89 use S,
# root_of_unity, multiplicative_generator)
# ---------------------------------------------------------------------------
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
FR_S = 32
FR_GENERATOR = 7
ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD)
G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G1_B = 4

FR_MONT_R = (1 << 256) % R_MOD
FQ_MONT_R = (1 << 384) % P_MOD
MASK64 = (1 << 64) - 1


# ---------------------------------------------------------------------------
# Encodings across the C ABI (blst_fr / blst_fp: little-endian u64 limbs,
# Montgomery form) and comparison encodings
# ---------------------------------------------------------------------------
def fr_to_mont_limbs(a: int) -> List[int]:
    v = (a * FR_MONT_R) % R_MOD
    return [(v >> (64 * i)) & MASK64 for i in range(4)]


def fr_from_mont_limbs(l: Sequence[int]) -> int:
    v = sum(int(x) << (64 * i) for i, x in enumerate(l))
    return (v * pow(FR_MONT_R, -1, R_MOD)) % R_MOD


def fq_to_mont_limbs(a: int) -> List[int]:
    v = (a * FQ_MONT_R) % P_MOD
    return [(v >> (64 * i)) & MASK64 for i in range(6)]


def fq_from_mont_limbs(l: Sequence[int]) -> int:
    v = sum(int(x) << (64 * i) for i, x in enumerate(l))
    return (v * pow(FQ_MONT_R, -1, P_MOD)) % P_MOD


def fr_to_bytes_le(a: int) -> bytes:
    return int(a % R_MOD).to_bytes(32, "little")


# ---------------------------------------------------------------------------
# Deterministic synthetic inputs: counter-based splitmix64
# ---------------------------------------------------------------------------
def splitmix64(idx: int) -> int:
    """Output #idx of the splitmix64 stream with state0 = 0 (counter form)."""
    z = ((idx + 1) * 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def fr_stream(seed: int, n: int, start: int = 0) -> List[int]:
    """Element i = (4 consecutive outputs as a 256-bit LE integer) mod r.
    Stream `seed` occupies counters [seed << 40, ...)."""
    out = []
    base = (seed << 40) & MASK64
    for i in range(start, start + n):
        c = base + 4 * i
        v = (splitmix64(c) | (splitmix64(c + 1) << 64)
             | (splitmix64(c + 2) << 128) | (splitmix64(c + 3) << 192))
        out.append(v % R_MOD)
    return out


def fr_stream_u64(seed: int, n: int, start: int = 0) -> List[int]:
    """Reference-style scalars: one u64 per scalar
    (benches/commit_coeff_form.rs:16-20, coeff_form.rs:213)."""
    base = (seed << 40) & MASK64
    return [splitmix64(base + 4 * i) for i in range(start, start + n)]


# ---------------------------------------------------------------------------
# G1 arithmetic (stands in for blstrs G1Projective / blst_p1; Jacobian)
# Points are (X, Y, Z) integer tuples; identity has Z == 0.
# ---------------------------------------------------------------------------
G1_IDENTITY = (0, 1, 0)
G1_GENERATOR = (G1_X, G1_Y, 1)


def g1_is_identity(P) -> bool:
    return P[2] % P_MOD == 0


def g1_on_curve_affine(x: int, y: int) -> bool:
    return (y * y - x * x * x - G1_B) % P_MOD == 0


def g1_double(P):
    X1, Y1, Z1 = P
    if Z1 == 0 or Y1 == 0:
        return G1_IDENTITY
    p = P_MOD
    A = X1 * X1 % p
    B = Y1 * Y1 % p
    C = B * B % p
    D = 2 * ((X1 + B) * (X1 + B) - A - C) % p
    E = 3 * A % p
    F = E * E % p
    X3 = (F - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y1 * Z1 % p
    return (X3, Y3, Z3)


def g1_add(P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    p = P_MOD
    Z1Z1 = Z1 * Z1 % p
    Z2Z2 = Z2 * Z2 % p
    U1 = X1 * Z2Z2 % p
    U2 = X2 * Z1Z1 % p
    S1 = Y1 * Z2 * Z2Z2 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if U1 == U2:
        if S1 == S2:
            return g1_double(P)
        return G1_IDENTITY
    H = (U2 - U1) % p
    Rr = (S2 - S1) % p
    HH = H * H % p
    HHH = H * HH % p
    V = U1 * HH % p
    X3 = (Rr * Rr - HHH - 2 * V) % p
    Y3 = (Rr * (V - X3) - S1 * HHH) % p
    Z3 = Z1 * Z2 * H % p
    return (X3, Y3, Z3)


def g1_neg(P):
    return (P[0], (-P[1]) % P_MOD, P[2])


def g1_mul(P, k: int):
    """Scalar multiplication (`G1Projective * Scalar`; lib.rs:44,
    coeff_form.rs:75,99; eval_form.rs:133)."""
    k %= R_MOD
    acc = G1_IDENTITY
    for bit in bin(k)[2:] if k else "":
        acc = g1_double(acc)
        if bit == "1":
            acc = g1_add(acc, P)
    return acc


def g1_to_affine(P) -> Optional[Tuple[int, int]]:
    """`.to_affine()` (coeff_form.rs:63,78,107; eval_form.rs:120,139).
    Returns None for the identity."""
    X, Y, Z = P
    if Z % P_MOD == 0:
        return None
    zi = pow(Z, -1, P_MOD)
    zi2 = zi * zi % P_MOD
    return (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)


def g1_from_affine(a: Optional[Tuple[int, int]]):
    return G1_IDENTITY if a is None else (a[0], a[1], 1)


def g1_eq(P, Q) -> bool:
    return g1_to_affine(P) == g1_to_affine(Q)


def g1_compress(P) -> bytes:
    """ZCash/IETF 48-byte compressed encoding (what blstrs `to_compressed`
    emits): big-endian x, bit7 = compressed, bit6 = infinity, bit5 = y is the
    lexicographically larger root."""
    a = g1_to_affine(P)
    if a is None:
        return bytes([0xC0]) + bytes(47)
    x, y = a
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P_MOD - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def g1_affine_compress(a: Optional[Tuple[int, int]]) -> bytes:
    return g1_compress(g1_from_affine(a))


# blst_p1 / blst_p1_affine images (Montgomery limbs) for the C ABI
def g1_to_blst_p1(P) -> List[int]:
    """144-byte Jacobian image as 18 u64.  Identity: z = 0."""
    if g1_is_identity(P):
        return [0] * 18
    return fq_to_mont_limbs(P[0]) + fq_to_mont_limbs(P[1]) + fq_to_mont_limbs(P[2])


def g1_from_blst_p1(l: Sequence[int]):
    x, y, z = fq_from_mont_limbs(l[0:6]), fq_from_mont_limbs(l[6:12]), fq_from_mont_limbs(l[12:18])
    return G1_IDENTITY if z == 0 else (x, y, z)


def g1_affine_to_blst(a: Optional[Tuple[int, int]]) -> List[int]:
    """96-byte affine image as 12 u64.  Identity: all-zero."""
    if a is None:
        return [0] * 12
    return fq_to_mont_limbs(a[0]) + fq_to_mont_limbs(a[1])


def g1_affine_from_blst(l: Sequence[int]) -> Optional[Tuple[int, int]]:
    if all(int(v) == 0 for v in l):
        return None
    return (fq_from_mont_limbs(l[0:6]), fq_from_mont_limbs(l[6:12]))


def g1_msm_naive(points: Sequence, scalars: Sequence[int]):
    """Definition of `G1Projective::multi_exp` (call sites coeff_form.rs:61,78,
    102; eval_form.rs:118,136): sum_i scalars[i] * points[i] over min(len)."""
    acc = G1_IDENTITY
    for P, k in zip(points, scalars):
        acc = g1_add(acc, g1_mul(P, k))
    return acc


def g1_msm_pippenger(points: Sequence, scalars: Sequence[int], c: int = 8):
    """Same value as g1_msm_naive via bucket method (faster for n >~ 64)."""
    n = min(len(points), len(scalars))
    nwin = (255 + c - 1) // c
    total = G1_IDENTITY
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = g1_double(total)
        buckets = [G1_IDENTITY] * (1 << c)
        for i in range(n):
            d = ((scalars[i] % R_MOD) >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = g1_add(buckets[d], points[i])
        run = G1_IDENTITY
        acc = G1_IDENTITY
        for d in range((1 << c) - 1, 0, -1):
            run = g1_add(run, buckets[d])
            acc = g1_add(acc, run)
        total = g1_add(total, acc)
    return total


# ---------------------------------------------------------------------------
# lib.rs
# ---------------------------------------------------------------------------
class KZGError(Exception):
    """lib.rs:26-36."""


class PointNotOnPolynomial(KZGError):
    pass


class PolynomialDegreeTooLarge(KZGError):
    pass


class KZGParams:
    """lib.rs:12-19 (G1 half only; `hs` is verifier-side, out of scope)."""

    def __init__(self, gs):
        self.gs = list(gs)


def setup(s: int, num_coeffs: int) -> KZGParams:
    """lib.rs:38-46: gs[0] = G, gs[i] = gs[i-1] * s."""
    gs = [G1_GENERATOR]
    curr = G1_GENERATOR
    for _ in range(1, num_coeffs):
        curr = g1_mul(curr, s)
        gs.append(curr)
    return KZGParams(gs)


def setup_fast(s: int, num_coeffs: int) -> KZGParams:
    """Same points as `setup` (gs[i] = s^i * G), each computed independently
    from the generator (identical group elements, different Jacobian
    representatives -- compare after to_affine)."""
    gs, e = [], 1
    for _ in range(num_coeffs):
        gs.append(g1_mul(G1_GENERATOR, e))
        e = e * s % R_MOD
    return KZGParams(gs)


# ---------------------------------------------------------------------------
# ft.rs
# ---------------------------------------------------------------------------
def compute_omega(d: int) -> Tuple[int, int, int]:
    """ft.rs:55-76: smallest power of two m >= d, exp = log2 m,
    omega = root_of_unity^(2^(S-exp)); error if exp >= S."""
    m, exp = 1, 0
    while m < d:
        m *= 2
        exp += 1
        if exp >= FR_S:
            raise PolynomialDegreeTooLarge()
    omega = pow(ROOT_OF_UNITY, 1 << (FR_S - exp), R_MOD)
    return m, exp, omega


def _bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def serial_fft(a: List[int], omega: int, log_n: int) -> None:
    """ft.rs:291-333: in place, bit-reversal then DIT butterflies;
    natural order in, natural order out."""
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = _bitreverse(k, log_n)
        if k < rk:
            a[rk], a[k] = a[k], a[rk]
    m = 1
    r = R_MOD
    for _ in range(log_n):
        w_m = pow(omega, n // (2 * m), r)
        k = 0
        while k < n:
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % r
                tmp = (a[k + j] - t) % r
                a[k + j + m] = tmp
                a[k + j] = (a[k + j] + t) % r
                w = w * w_m % r
            k += 2 * m
        m *= 2


def parallel_fft(a: List[int], omega: int, log_n: int, log_cpus: int) -> None:
    """ft.rs:336-387: one radix-2^log_cpus decimation step, independent
    sub-FFTs, interleave.  Same result as serial_fft."""
    assert log_n >= log_cpus
    num_cpus = 1 << log_cpus
    log_new_n = log_n - log_cpus
    r = R_MOD
    tmp = [[0] * (1 << log_new_n) for _ in range(num_cpus)]
    new_omega = pow(omega, num_cpus, r)
    for j in range(num_cpus):
        omega_j = pow(omega, j, r)
        omega_step = pow(omega, j << log_new_n, r)
        elt = 1
        for i in range(1 << log_new_n):
            for s in range(num_cpus):
                idx = (i + (s << log_new_n)) % (1 << log_n)
                tmp[j][i] = (tmp[j][i] + a[idx] * elt) % r
                elt = elt * omega_step % r
            elt = elt * omega_j % r
        serial_fft(tmp[j], new_omega, log_new_n)
    mask = num_cpus - 1
    for idx in range(len(a)):
        a[idx] = tmp[idx & mask][idx >> log_cpus]


def naive_dft(a: Sequence[int], omega: int) -> List[int]:
    """out[k] = sum_j a[j] * omega^(j*k)  (definition)."""
    n = len(a)
    r = R_MOD
    pw = [pow(omega, i, r) for i in range(n)]
    return [sum(a[j] * pw[(j * k) % n] for j in range(n)) % r for k in range(n)]


class EvaluationDomain:
    """ft.rs:17-25 and methods fft / ifft (ft.rs:111-140)."""

    def __init__(self, coeffs: List[int], d: int, exp: int, omega: int):
        # ft.rs:82-92
        self.coeffs = list(coeffs)
        self.d = d
        self.exp = exp
        self.omega = omega
        self.omegainv = pow(omega, -1, R_MOD)
        self.geninv = pow(FR_GENERATOR, -1, R_MOD)
        self.minv = pow(d % R_MOD, -1, R_MOD)

    @classmethod
    def from_coeffs(cls, coeffs: Sequence[int]) -> "EvaluationDomain":
        # ft.rs:94-109
        m, exp, omega = compute_omega(len(coeffs))
        c = list(coeffs) + [0] * (m - len(coeffs))
        return cls(c, m, exp, omega)

    def __len__(self):
        return len(self.coeffs)

    def clone_with_different_coeffs(self, coeffs):
        return EvaluationDomain(coeffs, self.d, self.exp, self.omega)

    def fft(self):
        serial_fft(self.coeffs, self.omega, self.exp)

    def ifft(self):
        serial_fft(self.coeffs, self.omegainv, self.exp)
        self.coeffs = [v * self.minv % R_MOD for v in self.coeffs]

    def distribute_powers(self, g: int):
        # ft.rs:142-166: coeffs[i] *= g^i
        self.coeffs = [v * pow(g, i, R_MOD) % R_MOD for i, v in enumerate(self.coeffs)]

    def coset_fft(self):
        # ft.rs:168-171
        self.distribute_powers(FR_GENERATOR)
        self.fft()

    def icoset_fft(self):
        # ft.rs:173-178
        self.ifft()
        self.distribute_powers(self.geninv)

    def z(self, tau: int) -> int:
        # ft.rs:182-187
        return (pow(tau, len(self.coeffs), R_MOD) - 1) % R_MOD

    def divide_by_z_on_coset(self):
        # ft.rs:192-217
        i = pow(self.z(FR_GENERATOR), -1, R_MOD)
        self.coeffs = [v * i % R_MOD for v in self.coeffs]

    def sub_assign(self, other: "EvaluationDomain"):
        # ft.rs:247-271
        assert len(self.coeffs) == len(other.coeffs)
        self.coeffs = [(a - b) % R_MOD for a, b in zip(self.coeffs, other.coeffs)]

    def mul_assign(self, other: "EvaluationDomain"):
        # ft.rs:220-244
        assert len(self.coeffs) == len(other.coeffs)
        self.coeffs = [a * b % R_MOD for a, b in zip(self.coeffs, other.coeffs)]

    def to_polynomial(self) -> "Polynomial":
        # ft.rs:27-31
        return Polynomial.new(self.coeffs)


# ---------------------------------------------------------------------------
# polynomial.rs
# ---------------------------------------------------------------------------
FFT_MUL_THRESHOLD = 128  # polynomial.rs:13


class Polynomial:
    """polynomial.rs:24-27: coeffs low -> high with an explicit `degree` that
    is authoritative (readers use coeffs[..degree+1], polynomial.rs:148-154)."""

    def __init__(self, coeffs: List[int], degree: int):
        # new_from_coeffs (polynomial.rs:90-92) -- trusts the caller
        self.coeffs = [c % R_MOD for c in coeffs]
        self.degree = degree

    @staticmethod
    def compute_degree(coeffs: Sequence[int], upper_bound: int) -> int:
        # polynomial.rs:94-105
        i = upper_bound
        while True:
            if i == 0:
                return 0
            if coeffs[i] != 0:
                return i
            i -= 1

    @classmethod
    def new(cls, coeffs: Sequence[int]) -> "Polynomial":
        # polynomial.rs:83-87
        c = [x % R_MOD for x in coeffs]
        return cls(c, cls.compute_degree(c, len(c) - 1))

    @classmethod
    def new_zero(cls):
        return cls([0], 0)

    def clone(self):
        return Polynomial(list(self.coeffs), self.degree)

    def is_zero(self) -> bool:
        # polynomial.rs:45-47
        return self.degree == 0 and self.coeffs[0] == 0

    def shrink_degree(self):
        # polynomial.rs:117-120
        self.degree = self.compute_degree(self.coeffs, self.degree)

    def lead(self) -> int:
        return self.coeffs[self.degree]

    def num_coeffs(self) -> int:
        return self.degree + 1

    def slice_coeffs(self) -> List[int]:
        return self.coeffs[: self.degree + 1]

    def __eq__(self, other) -> bool:
        # polynomial.rs:29-40
        if self.degree != other.degree:
            return False
        return all(l == r for l, r in zip(self.coeffs, other.coeffs))

    def eval(self, x: int) -> int:
        # polynomial.rs:156-165 (Horner)
        res = self.coeffs[self.degree]
        for i in range(self.degree - 1, -1, -1):
            res = (res * x + self.coeffs[i]) % R_MOD
        return res

    def mul_naive(self, rhs: "Polynomial") -> "Polynomial":
        # Mul for Polynomial, polynomial.rs:473-487
        res = [0] * (self.degree + rhs.degree + 1)
        for i in range(self.num_coeffs()):
            ai = self.coeffs[i]
            if ai == 0:
                continue
            for j in range(rhs.num_coeffs()):
                res[i + j] = (res[i + j] + ai * rhs.coeffs[j]) % R_MOD
        return Polynomial(res, self.degree + rhs.degree)

    def fft_mul(self, other: "Polynomial") -> "Polynomial":
        # polynomial.rs:167-183
        n, k = self.num_coeffs(), other.num_coeffs()
        lhs = list(self.coeffs)[: n + k] + [0] * max(0, n + k - len(self.coeffs))
        rhs = list(other.coeffs)[: n + k] + [0] * max(0, n + k - len(other.coeffs))
        L = EvaluationDomain.from_coeffs(lhs)
        Rr = EvaluationDomain.from_coeffs(rhs)
        L.fft()
        Rr.fft()
        L.mul_assign(Rr)
        L.ifft()
        return L.to_polynomial()

    def best_mul(self, other: "Polynomial") -> "Polynomial":
        # polynomial.rs:185-191
        if self.degree < FFT_MUL_THRESHOLD or other.degree < FFT_MUL_THRESHOLD:
            return self.mul_naive(other)
        return self.fft_mul(other)

    def long_division(self, divisor: "Polynomial") -> Tuple["Polynomial", Optional["Polynomial"]]:
        """polynomial.rs:193-227 (schoolbook, with the shrink_degree skip at
        :218 and the early exit on a zero remainder at :209)."""
        if self.is_zero():
            return Polynomial.new_zero(), None
        if divisor.is_zero():
            raise ZeroDivisionError("divisor must not be zero!")
        if self.degree < divisor.degree:
            return Polynomial.new_zero(), self.clone()
        r = R_MOD
        remainder = self.clone()
        quotient = Polynomial([0] * (self.degree - divisor.degree + 1), self.degree - divisor.degree)
        lead_inverse = pow(divisor.lead(), -1, r)
        dcoeffs = divisor.coeffs[: divisor.degree + 1]
        while (not remainder.is_zero()) and remainder.degree >= divisor.degree:
            factor = remainder.lead() * lead_inverse % r
            i = remainder.degree - divisor.degree
            quotient.coeffs[i] = factor
            rc = remainder.coeffs
            for j, coeff in enumerate(dcoeffs):
                rc[i + j] = (rc[i + j] - coeff * factor) % r
            remainder.shrink_degree()
        if remainder.is_zero():
            return quotient, None
        return quotient, remainder

    def add_owned(self, rhs: "Polynomial") -> "Polynomial":
        # Add for Polynomial, polynomial.rs:412-428
        res, shorter = (rhs.clone(), self) if rhs.degree > self.degree else (self.clone(), rhs)
        for i in range(shorter.num_coeffs()):
            res.coeffs[i] = (res.coeffs[i] + shorter.coeffs[i]) % R_MOD
        return res

    def sub_ref(self, rhs: "Polynomial") -> "Polynomial":
        # Sub for &Polynomial, polynomial.rs:443-460
        res = self.clone()
        if rhs.num_coeffs() > self.num_coeffs():
            res.coeffs = res.coeffs + [0] * (rhs.num_coeffs() - len(res.coeffs))
            res.degree = rhs.degree
        for i in range(rhs.num_coeffs()):
            res.coeffs[i] = (res.coeffs[i] - rhs.coeffs[i]) % R_MOD
        res.shrink_degree()
        return res

    def multi_eval(self, xs: Sequence[int]) -> List[int]:
        # polynomial.rs:229-233
        assert len(xs) > self.degree
        tree = SubProductTree.new_from_points(xs)
        return tree.eval(xs, self)

    @staticmethod
    def lagrange_interpolation_with_tree(xs, ys, tree: "SubProductTree") -> "Polynomial":
        """polynomial.rs:237-264.  NB the one-point case returns X + (y - x)
        (:244-247), reproduced here."""
        assert len(xs) == len(ys)
        if len(xs) == 1:
            return Polynomial([(ys[0] - xs[0]) % R_MOD, 1], 1)
        m_prime = tree.product.clone()
        for i in range(1, m_prime.num_coeffs()):
            m_prime.coeffs[i] = m_prime.coeffs[i] * i % R_MOD
        m_prime.coeffs.pop(0)
        m_prime.degree -= 1
        ev = m_prime.multi_eval(xs)
        cs = [ys[i] * pow(c, -1, R_MOD) % R_MOD for i, c in enumerate(ev)]
        return tree.linear_mod_combination(cs)


class SubProductTree:
    """polynomial.rs:303-365."""

    def __init__(self, product, left=None, right=None):
        self.product = product
        self.left = left
        self.right = right

    @classmethod
    def new_from_points(cls, xs: Sequence[int]) -> "SubProductTree":
        # polynomial.rs:310-327
        n = len(xs)
        if n == 1:
            return cls(Polynomial([(-xs[0]) % R_MOD, 1], 1))
        left = cls.new_from_points(xs[: n // 2])
        right = cls.new_from_points(xs[n // 2:])
        return cls(left.product.best_mul(right.product), left, right)

    def eval(self, xs: Sequence[int], f: Polynomial) -> List[int]:
        # polynomial.rs:329-348 (unwraps remainders: raises if one is None)
        n = len(xs)
        if n == 1:
            return [f.eval(xs[0])]
        _, r0 = f.long_division(self.left.product)
        _, r1 = f.long_division(self.right.product)
        if r0 is None or r1 is None:
            raise RuntimeError("called `Option::unwrap()` on a `None` value")
        return self.left.eval(xs[: n // 2], r0) + self.right.eval(xs[n // 2:], r1)

    def linear_mod_combination(self, cs: Sequence[int]) -> Polynomial:
        # polynomial.rs:350-364
        n = len(cs)
        if n == 1:
            return Polynomial([cs[0]], 0)
        l = self.left.linear_mod_combination(cs[: n // 2])
        r = self.right.linear_mod_combination(cs[n // 2:])
        return self.right.product.best_mul(l).add_owned(self.left.product.best_mul(r))


# ---------------------------------------------------------------------------
# coeff_form.rs (prover half)
# ---------------------------------------------------------------------------
class KZGBatchWitness:
    # coeff_form.rs:14-35
    def __init__(self, r: Polynomial, w):
        self.r = r
        self.w = w  # affine (x, y) or None


class KZGProver:
    """coeff_form.rs:38-112.  `msm` lets tests choose naive / Pippenger."""

    def __init__(self, parameters: KZGParams, msm=g1_msm_pippenger):
        self.parameters = parameters
        self._msm = msm

    def commit(self, polynomial: Polynomial):
        # coeff_form.rs:59-64
        n = polynomial.num_coeffs()
        if n > len(self.parameters.gs):
            raise IndexError("range end index out of range for slice")  # panic at :60
        gs = self.parameters.gs[:n]
        return g1_to_affine(self._msm(gs, polynomial.slice_coeffs()))

    def create_witness(self, polynomial: Polynomial, point: Tuple[int, int]):
        # coeff_form.rs:66-81
        x, y = point
        dividend = polynomial.clone()
        dividend.coeffs[0] = (dividend.coeffs[0] - y) % R_MOD
        divisor = Polynomial([(-x) % R_MOD, 1], 1)
        psi, rem = dividend.long_division(divisor)
        if rem is not None:
            raise PointNotOnPolynomial()
        if psi.num_coeffs() == 1:
            return g1_to_affine(g1_mul(self.parameters.gs[0], psi.coeffs[0]))
        gs = self.parameters.gs[: psi.num_coeffs()]
        return g1_to_affine(self._msm(gs, psi.slice_coeffs()))

    def create_witness_batched(self, polynomial: Polynomial, xs, ys) -> KZGBatchWitness:
        # coeff_form.rs:83-111
        tree = SubProductTree.new_from_points(xs)
        interpolation = Polynomial.lagrange_interpolation_with_tree(xs, ys, tree)
        numerator = polynomial.sub_ref(interpolation)
        psi, rem = numerator.long_division(tree.product)
        if rem is not None:
            raise PointNotOnPolynomial()
        if psi.num_coeffs() == 1:
            w = g1_mul(self.parameters.gs[0], psi.coeffs[0])
        else:
            w = self._msm(self.parameters.gs[: psi.num_coeffs()], psi.slice_coeffs())
        return KZGBatchWitness(interpolation, g1_to_affine(w))


# ---------------------------------------------------------------------------
# eval_form.rs (prover half)
# ---------------------------------------------------------------------------
def div_by_omega_i(evals: EvaluationDomain, m: int) -> EvaluationDomain:
    """eval_form.rs:58-84, literally (O(d) inversions)."""
    r = R_MOD
    coeffs = []
    omega_m = pow(evals.omega, m, r)
    for j, f in enumerate(evals.coeffs):
        if j == m:
            qm = 0
            am = evals.d * pow(omega_m, -1, r) % r
            for i in range(evals.d):
                if i == m:
                    continue
                omega_i = pow(evals.omega, i, r)
                ai = evals.d * pow(omega_i, -1, r) % r
                term = evals.coeffs[i]
                term = term * (am * pow(ai, -1, r) % r) % r
                term = term * pow((omega_m - omega_i) % r, -1, r) % r
                qm = (qm + term) % r
            coeffs.append(qm)
        else:
            omega_j = pow(evals.omega, j, r)
            coeffs.append(f * pow((omega_j - omega_m) % r, -1, r) % r)
    return evals.clone_with_different_coeffs(coeffs)


def div_by_omega_i_fast(coeffs: Sequence[int], omega: int, m: int) -> List[int]:
    """Closed form of div_by_omega_i (a_m/a_i = omega^(i-m)); same values, used
    at sizes where the literal version is too slow in Python."""
    r = R_MOD
    d = len(coeffs)
    pw = [1] * d
    for i in range(1, d):
        pw[i] = pw[i - 1] * omega % r
    om = pw[m]
    out = [0] * d
    qm = 0
    for j in range(d):
        if j == m:
            continue
        inv = pow((pw[j] - om) % r, -1, r)
        out[j] = coeffs[j] * inv % r
        # term_i = f_i * omega^(i-m) / (omega^m - omega^i) = -f_i * omega^(i-m) * inv
        qm = (qm - coeffs[j] * pw[(j - m) % d] % r * inv) % r
    out[m] = qm
    return out


class KZGProverEvalForm:
    """eval_form.rs:40-46, 86-147."""

    def __init__(self, parameters: KZGParams, lagrange_basis_g, msm=g1_msm_pippenger):
        # eval_form.rs:88-100
        d, exp, omega = compute_omega(len(parameters.gs))
        self.parameters = parameters
        self.lagrange_basis_g = list(lagrange_basis_g)
        self.d, self.exp, self.omega = d, exp, omega
        self._msm = msm

    def commit(self, evals: EvaluationDomain):
        # eval_form.rs:114-122
        assert self.d == evals.d
        gs = self.lagrange_basis_g[: len(evals)]
        return g1_to_affine(self._msm(gs, evals.coeffs))

    def create_witness(self, evals: EvaluationDomain, i: int):
        # eval_form.rs:124-140
        y = evals.coeffs[i]
        numerator = evals.clone_with_different_coeffs([(c - y) % R_MOD for c in evals.coeffs])
        q = div_by_omega_i(numerator, i)
        if len(q.coeffs) == 1:
            w = g1_mul(self.lagrange_basis_g[0], q.coeffs[0])
        else:
            w = self._msm(self.lagrange_basis_g[: len(q)], q.coeffs)
        return g1_to_affine(w)

    def create_witness_all(self):
        # eval_form.rs:142-146
        return None


def lagrange_basis_known_secret(s: int, d: int) -> List:
    """Lagrange-basis SRS L_i(s)*G in natural omega^i order (same points as
    compute_lagrange_basis, eval_form.rs:254-280, which is O(d^3)); closed form
    L_i(s) = (s^d - 1) * omega^i / (d * (s - omega^i)).  d a power of two."""
    _, _, omega = compute_omega(d)
    r = R_MOD
    zs = (pow(s, d, r) - 1) % r
    dinv = pow(d, -1, r)
    out = []
    w = 1
    for _ in range(d):
        li = zs * w % r * dinv % r * pow((s - w) % r, -1, r) % r
        out.append(g1_mul(G1_GENERATOR, li))
        w = w * omega % r
    return out


def lagrange_scalars_known_secret(s: int, d: int) -> List[int]:
    _, _, omega = compute_omega(d)
    r = R_MOD
    zs = (pow(s, d, r) - 1) % r
    dinv = pow(d, -1, r)
    out, w = [], 1
    for _ in range(d):
        out.append(zs * w % r * dinv % r * pow((s - w) % r, -1, r) % r)
        w = w * omega % r
    return out
