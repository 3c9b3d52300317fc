"""coeff_form.rs (prover half): KZGProver, KZGBatchWitness -- host mirror over the C ABI."""
from __future__ import annotations

import numpy as np

from . import _ffi
from .params import KZGParams, PointNotOnPolynomial
from .polynomial import Polynomial


class KZGBatchWitness:
    """coeff_form.rs:14-35."""

    def __init__(self, r: Polynomial, w: np.ndarray):
        self.r = r
        self.w = w

    def elem(self) -> np.ndarray:
        return self.w

    def polynomial(self) -> Polynomial:
        return self.r

    def __eq__(self, other) -> bool:
        return self.r == other.r and bool((self.w == other.w).all())


class KZGProver:
    """coeff_form.rs:38-112."""

    def __init__(self, parameters: KZGParams):
        self._parameters = parameters

    def parameters(self) -> KZGParams:
        return self._parameters

    def commit(self, polynomial: Polynomial) -> np.ndarray:
        """coeff_form.rs:59-64 -> G1Affine (uint64[12])."""
        n = polynomial.num_coeffs()
        if n > len(self._parameters):
            # `&self.parameters.gs[..polynomial.num_coeffs()]` panics (coeff_form.rs:60)
            raise IndexError(f"range end index {n} out of range for slice of length {len(self._parameters)}")
        if polynomial.device is not None:      # resident copy: no per-call upload (Polynomial.to_device)
            return polynomial.device.commit(self._parameters.srs, n)
        return self._parameters.srs.msm(polynomial.slice_coeffs())

    def commit_many(self, polynomials) -> np.ndarray:
        """`commit` (coeff_form.rs:59-64) of every polynomial in one pipelined call -> [count, 12] G1Affine.
        Same results as calling `commit` in a loop; an extension of the reference API (a prover usually commits to
        several polynomials against one SRS) that lets one MSM's tail overlap the next one's copy and accumulation."""
        polynomials = list(polynomials)
        for p in polynomials:
            if p.num_coeffs() > len(self._parameters):
                raise IndexError(f"range end index {p.num_coeffs()} out of range for slice of length {len(self._parameters)}")
        return self._parameters.srs.msm_batch([p.slice_coeffs() for p in polynomials])

    def create_witness(self, polynomial: Polynomial, point) -> np.ndarray:
        """coeff_form.rs:66-81.  Raises PointNotOnPolynomial when (X - x) does not divide p - y."""
        x, y = point
        coeffs = polynomial.slice_coeffs()
        # degree-0 dividends never reach the device: zero -> ([0], None) -> gs[0]*0; non-zero constant ->
        # ([0], Some(self)) -> Err (polynomial.rs:194-199).  The C ABI implements exactly the same rule.
        if polynomial.device is not None:
            st, w = polynomial.device.witness(self._parameters.srs, polynomial.num_coeffs(), x, y)
        else:
            st, w = _ffi.witness_coeff(self._parameters.srs, coeffs, x, y)
        if st == _ffi.KZGB_POINT_NOT_ON_POLYNOMIAL:
            raise PointNotOnPolynomial()
        return w

    def create_witness_many(self, polynomial: Polynomial, xs, ys) -> np.ndarray:
        """`create_witness` (coeff_form.rs:66-81) at every (xs[i], ys[i]) in one pipelined call -> [m, 12] G1Affine:
        the polynomial is uploaded once and witness i's tail overlaps witness i+1's scan and accumulation.  Like
        collecting the reference's Results: raises PointNotOnPolynomial (with `.index` = first offending point) if
        any point is not on the polynomial.  An addition to the reference API for multi-opening workloads."""
        st, w = _ffi.witness_coeff_many(self._parameters.srs, polynomial.slice_coeffs(), xs, ys)
        bad = np.nonzero(st == _ffi.KZGB_POINT_NOT_ON_POLYNOMIAL)[0]
        if len(bad):
            err = PointNotOnPolynomial()
            err.index = int(bad[0])
            raise err
        return w

    def create_witness_batched(self, polynomial: Polynomial, xs, ys) -> KZGBatchWitness:
        """coeff_form.rs:83-111."""
        xs = _ffi.u64(xs, (-1, 4))
        ys = _ffi.u64(ys, (-1, 4))
        assert len(xs) == len(ys)  # polynomial.rs:242
        if polynomial.device is not None:
            st, r, w = polynomial.device.witness_batched(self._parameters.srs, polynomial.num_coeffs(), xs, ys)
        else:
            st, r, w = _ffi.witness_batched(self._parameters.srs, polynomial.slice_coeffs(), xs, ys)
        if st == _ffi.KZGB_POINT_NOT_ON_POLYNOMIAL:
            raise PointNotOnPolynomial()
        return KZGBatchWitness(Polynomial.from_interpolant(r, len(xs)), w)


class KZGVerifier:
    """coeff_form.rs:43-45, 114-183 -- only the MSM-backed check is provided.  `verify_eval` and
    `verify_eval_batched` need pairings and G2 arithmetic, which stay on the CPU side of an integration."""

    def __init__(self, parameters: KZGParams):
        self._parameters = parameters

    def verify_poly(self, commitment, polynomial: Polynomial) -> bool:
        # coeff_form.rs:119-124: recompute the commitment and compare
        n = polynomial.num_coeffs()
        if n > len(self._parameters):
            raise IndexError(f"range end index {n} out of range for slice of length {len(self._parameters)}")
        check = self._parameters.srs.msm(polynomial.slice_coeffs())
        return bool((check == np.asarray(commitment, dtype=np.uint64).reshape(12)).all())

    def verify_eval(self, *a, **k):
        raise NotImplementedError("pairing-based verification (coeff_form.rs:126-142) is outside this library")

    verify_eval_batched = verify_eval
