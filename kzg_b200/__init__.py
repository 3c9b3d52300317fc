"""kzg_b200 -- B200-native KZG prover hot path (G1 MSM, Fr NTT, quotient division) behind the
proxima-one/kzg prover API (KZGParams / KZGProver / KZGProverEvalForm / EvaluationDomain / Polynomial).

The compute path is libkzg_b200.so (hand-written sm_100a CUDA, C ABI in include/kzg_b200.h); this package is
the host-side mirror of the reference's Rust API used by the tests and the benchmark.  There is no CPU
fallback: every operation raises if the CUDA library or a B200 is missing."""
from .coeff_form import KZGBatchWitness, KZGProver, KZGVerifier
from .eval_form import (KZGProverEvalForm, KZGVerifierEvalForm, compute_lagrange_basis, compute_lagrange_basis_known_secret,
                        div_by_omega_i)
from .ft import EvaluationDomain, compute_omega
from .params import (BatchOpeningZeroRemainder, KZGError, KZGParams, NoPolynomial, PointNotOnPolynomial,
                     PolynomialDegreeTooLarge, setup)
from .polynomial import Polynomial
from .scalar import fr, fr_int, fr_vec, fr_vec_int, g1_affine_compress

__all__ = [
    "KZGBatchWitness", "KZGProver", "KZGVerifier", "KZGVerifierEvalForm", "KZGProverEvalForm", "compute_lagrange_basis", "compute_lagrange_basis_known_secret", "div_by_omega_i",
    "EvaluationDomain", "compute_omega", "KZGError", "KZGParams", "NoPolynomial", "PointNotOnPolynomial",
    "BatchOpeningZeroRemainder", "PolynomialDegreeTooLarge", "setup", "Polynomial", "fr", "fr_int", "fr_vec",
    "fr_vec_int", "g1_affine_compress",
]
