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

    @classmethod
    def from_coeffs(cls, coeffs) -> "EvaluationDomain":
        # ft.rs:94-109: pad with zeros to the next power of two
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        m, exp, omega = compute_omega(len(c))
        if m > len(c):
            c = np.vstack([c, np.zeros((m - len(c), 4), dtype=np.uint64)])
        return cls(c, m, exp, omega)

    def __len__(self) -> int:
        return len(self.coeffs)

    def as_ref(self) -> np.ndarray:
        return self.coeffs

    def clone_with_different_coeffs(self, coeffs) -> "EvaluationDomain":
        return EvaluationDomain(coeffs, self.d, self.exp, self.omega)

    def _check(self):
        if len(self.coeffs) != (1 << self.exp):
            raise AssertionError("assertion failed: n == 1 << log_n")  # ft.rs:302

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

    def into_polynomial(self) -> Polynomial:
        # From<EvaluationDomain> for Polynomial, ft.rs:27-31
        return Polynomial.new(self.coeffs)

    def __eq__(self, other) -> bool:
        return (self.d == other.d and self.exp == other.exp and self.omega == other.omega
                and self.coeffs.shape == other.coeffs.shape and bool((self.coeffs == other.coeffs).all()))
