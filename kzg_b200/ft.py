"""ft.rs: EvaluationDomain -- host mirror; fft/ifft run on the device."""
from __future__ import annotations

import numpy as np

from . import _ffi
from .params import PolynomialDegreeTooLarge
from .polynomial import Polynomial
from .scalar import FR_S, R_MOD, ROOT_OF_UNITY, fr


def compute_omega(d: int):
    """ft.rs:55-76: (m, exp, omega) with m the next power of two >= d."""
    m, exp = 1, 0
    while m < d:
        m *= 2
        exp += 1
        if exp >= FR_S:
            raise PolynomialDegreeTooLarge()
    omega = pow(ROOT_OF_UNITY, 1 << (FR_S - exp), R_MOD)
    return m, exp, omega


class EvaluationDomain:
    """ft.rs:17-25."""

    def __init__(self, coeffs, d: int, exp: int, omega: int):
        # ft.rs:82-92
        self.coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        self.d = d
        self.exp = exp
        self.omega = omega
        self.omegainv = pow(omega, -1, R_MOD)
        self.geninv = pow(7, -1, R_MOD)
        self.minv = pow(d % R_MOD, -1, R_MOD)

    @property
    def coeffs(self) -> np.ndarray:
        return self._coeffs

    @coeffs.setter
    def coeffs(self, value) -> None:
        self._coeffs = value
        self.release_device()     # a resident copy of the old values would be stale

    @classmethod
    def from_coeffs(cls, coeffs) -> "EvaluationDomain":
        # ft.rs:94-109: pad with zeros to the next power of two
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        m, exp, omega = compute_omega(len(c))
        if m > len(c):
            c = np.vstack([c, np.zeros((m - len(c), 4), dtype=np.uint64)])
        return cls(c, m, exp, omega)

    # ---- device residency ----
    def to_device(self) -> "EvaluationDomain":
        """Keeps a copy of the evaluations in HBM (kzgb_poly_upload): KZGProverEvalForm.commit / create_witness then
        skip the per-call upload.  Host-side mutations (fft, ...) drop the resident copy."""
        self.release_device()
        self._dev = _ffi.DevicePoly(self.coeffs)
        return self

    def release_device(self) -> None:
        d = getattr(self, "_dev", None)
        if d is not None:
            d.close()
        self._dev = None

    @property
    def device(self):
        return getattr(self, "_dev", None)

    def __len__(self) -> int:
        return len(self.coeffs)

    def as_ref(self) -> np.ndarray:
        return self.coeffs

    def clone_with_different_coeffs(self, coeffs) -> "EvaluationDomain":
        return EvaluationDomain(coeffs, self.d, self.exp, self.omega)

    def _check(self):
        if len(self.coeffs) != (1 << self.exp):
            raise AssertionError("assertion failed: n == 1 << log_n")  # ft.rs:302
        # The device transform always uses the canonical 2^exp-th root root_of_unity^(2^(32-exp)) and the scale
        # 2^-exp (what compute_omega / from_coeffs produce, ft.rs:55-109).  The reference's best_fft would honour ANY
        # omega / d passed to `new`; a domain built with other parameters must fail loudly here, not transform
        # silently with different constants.
        if self.d != (1 << self.exp) or self.omega != pow(ROOT_OF_UNITY, 1 << (FR_S - self.exp), R_MOD):
            raise ValueError("EvaluationDomain: only the canonical domain (d = 2^exp, omega = compute_omega(d)) is "
                             "supported by the device transform")

    def fft(self) -> None:
        # ft.rs:111-113
        self._check()
        st, out = _ffi.ntt(self.coeffs, self.exp, False)
        if st == _ffi.KZGB_DEGREE_TOO_LARGE:
            raise PolynomialDegreeTooLarge()
        self.coeffs = out

    def ifft(self) -> None:
        # ft.rs:115-140
        self._check()
        st, out = _ffi.ntt(self.coeffs, self.exp, True)
        if st == _ffi.KZGB_DEGREE_TOO_LARGE:
            raise PolynomialDegreeTooLarge()
        self.coeffs = out

    # ---- helpers off the prover path (ft.rs:142-271) ----
    def distribute_powers(self, g: int) -> None:
        # ft.rs:142-166: coeffs[i] *= g^i
        self.coeffs = _ffi.fr_vec_op(2, self.coeffs, None, fr(g))

    def coset_fft(self) -> None:
        # ft.rs:168-171
        self._check()
        self.coeffs = _ffi.coset_ntt(self.coeffs, self.exp, False)[1]

    def icoset_fft(self) -> None:
        # ft.rs:173-178
        self._check()
        self.coeffs = _ffi.coset_ntt(self.coeffs, self.exp, True)[1]

    def z(self, tau: int) -> int:
        # ft.rs:182-187: tau^m - 1
        return (pow(tau, len(self.coeffs), R_MOD) - 1) % R_MOD

    def divide_by_z_on_coset(self) -> None:
        # ft.rs:192-217: multiply by z(7)^-1
        i = pow(self.z(7), -1, R_MOD)
        self.coeffs = _ffi.fr_vec_op(3, self.coeffs, None, fr(i))

    def mul_assign(self, other: "EvaluationDomain") -> None:
        # ft.rs:220-244
        assert len(self.coeffs) == len(other.coeffs)
        self.coeffs = _ffi.fr_vec_op(0, self.coeffs, other.coeffs)

    def sub_assign(self, other: "EvaluationDomain") -> None:
        # ft.rs:247-271
        assert len(self.coeffs) == len(other.coeffs)
        self.coeffs = _ffi.fr_vec_op(1, self.coeffs, other.coeffs)

    def into_polynomial(self) -> Polynomial:
        # From<EvaluationDomain> for Polynomial, ft.rs:27-31
        return Polynomial.new(self.coeffs)

    def __eq__(self, other) -> bool:
        return (self.d == other.d and self.exp == other.exp and self.omega == other.omega
                and self.coeffs.shape == other.coeffs.shape and bool((self.coeffs == other.coeffs).all()))
