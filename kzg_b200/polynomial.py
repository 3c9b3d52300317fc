"""polynomial.rs: `Polynomial` host-side bookkeeping (degree rules) over device arithmetic."""
from __future__ import annotations

import numpy as np

from . import _ffi
from .scalar import FR_ZERO, fr, fr_int, fr_sub


def _is_zero(row) -> bool:
    return not np.asarray(row).any()


class Polynomial:
    """polynomial.rs:24-27.  coeffs: uint64[len, 4] Montgomery scalars, low -> high; `degree` is authoritative:
    every hot-path reader uses coeffs[..degree+1] (polynomial.rs:148-154)."""

    def __init__(self, coeffs, degree: int):
        # new_from_coeffs (polynomial.rs:90-92): trusts the caller's degree
        self.coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        self.degree = int(degree)

    new_from_coeffs = classmethod(lambda cls, coeffs, degree: cls(coeffs, degree))

    @staticmethod
    def compute_degree(coeffs: np.ndarray, upper_bound: int) -> int:
        # polynomial.rs:94-105
        nz = np.nonzero(coeffs[: upper_bound + 1].any(axis=1))[0]
        return int(nz[-1]) if len(nz) else 0

    @classmethod
    def new(cls, coeffs) -> "Polynomial":
        # polynomial.rs:83-87
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        return cls(c, cls.compute_degree(c, len(c) - 1))

    @classmethod
    def new_zero(cls) -> "Polynomial":
        return cls(np.zeros((1, 4), dtype=np.uint64), 0)

    @classmethod
    def from_ints(cls, values) -> "Polynomial":
        from .scalar import fr_vec
        return cls.new(fr_vec(values))

    def clone(self) -> "Polynomial":
        return Polynomial(self.coeffs.copy(), self.degree)

    # ---- device residency ----
    def to_device(self) -> "Polynomial":
        """Uploads the coefficients once (kzgb_poly_upload); commit / create_witness / create_witness_batched of a
        prover then run on the resident copy instead of re-uploading per call.  The host array stays authoritative for
        the bookkeeping (degree rules, equality); mutate it and the resident copy is stale -- call `to_device` again."""
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
        """The resident copy (an _ffi.DevicePoly) or None."""
        return getattr(self, "_dev", None)

    def is_zero(self) -> bool:
        # polynomial.rs:45-47
        return self.degree == 0 and _is_zero(self.coeffs[0])

    def shrink_degree(self) -> None:
        # polynomial.rs:117-120
        self.degree = self.compute_degree(self.coeffs, self.degree)

    def num_coeffs(self) -> int:
        return self.degree + 1

    def slice_coeffs(self) -> np.ndarray:
        return self.coeffs[: self.degree + 1]

    def lead(self) -> np.ndarray:
        return self.coeffs[self.degree]

    def __eq__(self, other) -> bool:
        # polynomial.rs:29-40: degree, then zip(coeffs) (stops at the shorter vector)
        if self.degree != other.degree:
            return False
        k = min(len(self.coeffs), len(other.coeffs))
        return bool((self.coeffs[:k] == other.coeffs[:k]).all())

    def eval(self, x) -> np.ndarray:
        """polynomial.rs:156-165 (Horner) -- computed on the device as the remainder of the linear scan."""
        if isinstance(x, int):
            x = fr(x)
        _, rem = _ffi.divide_linear(self.slice_coeffs(), x, None)
        return rem

    def long_division_linear(self, x):
        """long_division by X - x (polynomial.rs:193-227 for a degree-1 divisor).
        Returns (quotient, remainder_or_None) with the reference's conventions."""
        if self.is_zero():
            return Polynomial.new_zero(), None  # :194-195
        if self.degree < 1:
            return Polynomial.new_zero(), self.clone()  # :198-199
        q, rem = _ffi.divide_linear(self.slice_coeffs(), x, None)
        quotient = Polynomial(q, self.degree - 1)  # not degree-shrunk (:202-205)
        if _is_zero(rem):
            return quotient, None
        return quotient, Polynomial(rem.reshape(1, 4), 0)

    def sub_constant(self, y) -> "Polynomial":
        """dividend.coeffs[0] -= y (coeff_form.rs:67-68); keeps the declared degree."""
        r = self.clone()
        r.coeffs[0] = fr_sub(r.coeffs[0], y)
        return r

    # ---- multiplication (off the prover path; polynomial.rs:167-191, 473-487) ----
    FFT_MUL_THRESHOLD = 128  # polynomial.rs:13

    def fft_mul(self, other: "Polynomial") -> "Polynomial":
        """polynomial.rs:167-183: operands padded to 2^k >= n + k coefficients, fft, pointwise product, ifft; the
        result keeps all 2^k coefficients and its degree is re-measured (Polynomial::new)."""
        return Polynomial.new(_ffi.poly_mul(self.slice_coeffs(), other.slice_coeffs()))

    def mul(self, other: "Polynomial") -> "Polynomial":
        """Mul for Polynomial (polynomial.rs:473-487): degree+degree+1 coefficients, declared degree = the sum
        (not shrunk).  Same coefficients as the schoolbook loop, computed through the transform pipeline."""
        full = _ffi.poly_mul(self.slice_coeffs(), other.slice_coeffs())
        deg = self.degree + other.degree
        return Polynomial(full[: deg + 1].copy(), deg)

    def best_mul(self, other: "Polynomial") -> "Polynomial":
        # polynomial.rs:185-191
        if self.degree < self.FFT_MUL_THRESHOLD or other.degree < self.FFT_MUL_THRESHOLD:
            return self.mul(other)
        return self.fft_mul(other)

    @staticmethod
    def from_interpolant(r, m: int) -> "Polynomial":
        """Wraps the m interpolant coefficients the device returns in the SHAPE the reference's tree recursion gives
        them (polynomial.rs:303-365): the top-level `right.product.best_mul(&l) + left.product.best_mul(&r)` takes the
        fft_mul path once both operands have degree >= 128 (polynomial.rs:185-191), i.e. for m >= 258, and fft_mul
        returns the full padded transform length with a re-measured degree (polynomial.rs:167-183, Polynomial::new);
        `Add` keeps the longer operand's vector.  Below the threshold: exactly m coefficients, declared degree m - 1."""
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
        if m >= 258:
            size = 1
            while size < m + 1:
                size *= 2
            padded = np.zeros((size, 4), dtype=np.uint64)
            padded[: len(r)] = r
            # re-measured (Polynomial::new): generically m - 1; lower only if the top coefficients really vanish
            return Polynomial(padded, Polynomial.compute_degree(padded, len(r) - 1))
        return Polynomial(r, len(r) - 1)

    @staticmethod
    def lagrange_interpolation(xs, ys) -> "Polynomial":
        """polynomial.rs:237-293.  One point gives X + (y - x) (degree 1) exactly as the reference does
        (:244-247); m >= 2 points give m coefficients with declared degree m - 1 (m >= 258: see from_interpolant)."""
        r = _ffi.interpolate(xs, ys)
        return Polynomial.from_interpolant(r, len(_ffi.u64(xs, (-1, 4))))
