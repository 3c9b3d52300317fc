"""eval_form.rs (prover half): KZGProverEvalForm over a Lagrange-basis SRS -- host mirror."""
from __future__ import annotations

import numpy as np

from . import _ffi
from .ft import EvaluationDomain, compute_omega
from .params import KZGParams
from .scalar import G1_IDENTITY_AFFINE, fr


def compute_lagrange_basis_known_secret(s, d: int, window_bits: int = 0) -> _ffi.SRS:
    """The points `compute_lagrange_basis` returns (eval_form.rs:254-280: L_i(s)*G in natural omega^i order),
    built from the known setup secret by the closed form L_i(s) = (s^d - 1) w^i / (d (s - w^i)) on the device
    (the reference's O(d^3) construction is unusable beyond toy sizes)."""
    if isinstance(s, int):
        s = fr(s)
    return _ffi.srs_generate(1, s, d, window_bits=window_bits)


def compute_lagrange_basis(params: KZGParams, window_bits: int = 0) -> _ffi.SRS:
    """eval_form.rs:254-280 `compute_lagrange_basis(params)` (G1 half): the Lagrange-basis SRS derived from the
    monomial SRS alone -- no secret needed.  The reference builds it in O(d^3) from Lagrange polynomials and MSMs;
    here it is the inverse NTT of `gs` over the group, L_i(s)*G = (1/d) sum_j omega^(-ij) gs[j].  `len(params)` must be
    a power of two (the reference asserts d == gs.len(), eval_form.rs:256)."""
    d = len(params)
    m, exp, _ = compute_omega(d)
    assert m == d, "assertion failed: d == params.gs.len()"
    return _ffi.srs_lagrange_from_monomial(params.srs, exp, window_bits)


def div_by_omega_i(evals: EvaluationDomain, m: int) -> EvaluationDomain:
    """eval_form.rs:58-84 applied to `evals - evals[m]` (as create_witness does, :125-131)."""
    if len(evals.coeffs) == 1:
        return evals.clone_with_different_coeffs(np.zeros((1, 4), dtype=np.uint64))
    return evals.clone_with_different_coeffs(_ffi.div_by_omega_i(evals.coeffs, evals.exp, m))


class KZGProverEvalForm:
    """eval_form.rs:40-46, 86-147."""

    def __init__(self, parameters: KZGParams, lagrange_basis_g: _ffi.SRS):
        # eval_form.rs:88-100
        d, exp, omega = compute_omega(len(parameters))
        self._parameters = parameters
        self.lagrange_basis_g = lagrange_basis_g
        self.d, self.exp, self._omega = d, exp, omega

    def parameters(self) -> KZGParams:
        return self._parameters

    def degree(self) -> int:
        return self.d

    def omega(self) -> int:
        return self._omega

    def commit(self, evals: EvaluationDomain) -> np.ndarray:
        # eval_form.rs:114-122
        assert self.d == evals.d
        if len(evals) > len(self.lagrange_basis_g):
            raise IndexError("range end index out of range for slice")  # :117
        if evals.device is not None:
            return evals.device.commit(self.lagrange_basis_g, len(evals))
        return self.lagrange_basis_g.msm(evals.coeffs)

    def commit_many(self, evals_list) -> np.ndarray:
        """`commit` (eval_form.rs:114-122) of several evaluation vectors in one pipelined call -> [count, 12]."""
        evals_list = list(evals_list)
        for evals in evals_list:
            assert self.d == evals.d
            if len(evals) > len(self.lagrange_basis_g):
                raise IndexError("range end index out of range for slice")
        return self.lagrange_basis_g.msm_batch([e.coeffs for e in evals_list])

    def create_witness(self, evals: EvaluationDomain, i: int) -> np.ndarray:
        # eval_form.rs:124-140 (infallible)
        if i >= len(evals.coeffs):
            raise IndexError("index out of bounds")  # evals.coeffs[i], :125
        if evals.device is not None:
            return evals.device.witness_eval(self.lagrange_basis_g, evals.exp, i)
        return _ffi.witness_eval(self.lagrange_basis_g, evals.coeffs, evals.exp, i)

    def create_witness_many(self, evals: EvaluationDomain, indices) -> np.ndarray:
        """`create_witness` (eval_form.rs:124-140) at every index of `indices` in one pipelined call -> [m, 12]:
        the evaluation vector is uploaded once and witness j's tail overlaps witness j+1's division and accumulation."""
        indices = [int(i) for i in indices]
        for i in indices:
            if i >= len(evals.coeffs):
                raise IndexError("index out of bounds")
        return _ffi.witness_eval_many(self.lagrange_basis_g, evals.coeffs, evals.exp, indices)

    def create_witness_all(self) -> np.ndarray:
        # eval_form.rs:142-146: identity * 0
        return G1_IDENTITY_AFFINE.copy()


class KZGVerifierEvalForm:
    """eval_form.rs:149-219 -- verify_poly only (iFFT + monomial MSM, :162-171); the pairing checks are out of scope."""

    def __init__(self, parameters: KZGParams):
        self._parameters = parameters

    def verify_poly(self, commitment, evals: EvaluationDomain) -> bool:
        coeffs = evals.clone_with_different_coeffs(evals.coeffs.copy())
        coeffs.ifft()
        p = coeffs.into_polynomial()
        check = self._parameters.srs.msm(p.slice_coeffs())
        return bool((check == np.asarray(commitment, dtype=np.uint64).reshape(12)).all())
