"""CPU suite, part 2: the C-ABI library loads and exports what include/kzg_b200.h declares; the host mirror's
bookkeeping (degree rules, encodings, domain parameters) behaves like the reference; no compute call works
without a GPU (there is no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import kzg_b200
from kzg_b200 import _ffi, scalar
from kzg_b200.ft import EvaluationDomain, compute_omega
from kzg_b200.polynomial import Polynomial
from oracle import kzg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kzg_b200.h")).read()
    return sorted(set(re.findall(r"KZGB_API\s+[\w\s\*]+?\b(kzgb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.load()
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/kzg_b200.h but not exported"


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "kzg_b200.h")).read()
    assert "torch" not in text.lower() and "at::" not in text


def test_layout_sizes_match_blst():
    class Fr(ctypes.Structure):
        _fields_ = [("l", ctypes.c_uint64 * 4)]

    class Fp(ctypes.Structure):
        _fields_ = [("l", ctypes.c_uint64 * 6)]

    class P1(ctypes.Structure):
        _fields_ = [("x", Fp), ("y", Fp), ("z", Fp)]

    class P1Affine(ctypes.Structure):
        _fields_ = [("x", Fp), ("y", Fp)]

    assert (ctypes.sizeof(Fr), ctypes.sizeof(Fp), ctypes.sizeof(P1), ctypes.sizeof(P1Affine)) == (32, 48, 144, 96)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="this check is for the GPU-less container")
def test_compute_fails_loudly_without_gpu():
    lib = _ffi.load()
    out = np.zeros(4, dtype=np.uint64)
    # not initialised -> negative status, never a silent CPU result
    st = lib.kzgb_ntt_fr(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(0), ctypes.c_int(0))
    assert st < 0
    assert lib.kzgb_init(ctypes.c_int(0)) < 0
    assert len(lib.kzgb_last_error_message()) > 0
    with pytest.raises(_ffi.KZGBackendError):
        _ffi.check(st)
    # every compute entry point of the hot path refuses to run without an initialised B200 (no CPU fallback):
    # all of them must return a negative status before touching their (here: null) arguments
    null = ctypes.c_void_p(0)
    z = ctypes.c_size_t(0)
    calls = {
        "kzgb_msm_g1": (null, null, z, null),
        "kzgb_msm_g1_batch": (null, null, null, z, null),
        "kzgb_msm_g1_batch_dev": (null, null, null, z, null, ctypes.c_int(0), null),
        "kzgb_msm_g1_batch_h2d": (null, null, null, z, null, ctypes.c_int(0), null),
        "kzgb_witness_coeff": (null, null, z, null, null, null),
        "kzgb_witness_coeff_many": (null, null, z, null, null, z, null, null),
        "kzgb_witness_eval_many": (null, null, ctypes.c_uint32(0), null, z, null),
        "kzgb_witness_batched": (null, null, z, null, null, z, null, null, null),
        "kzgb_divide_linear": (null, z, null, null, null, null),
        "kzgb_ntt4_columns_scatter_dev": (null, null, ctypes.c_uint32(1), ctypes.c_uint32(0), ctypes.c_uint32(4), ctypes.c_uint32(2),
                                          ctypes.c_uint32(2), ctypes.c_uint32(0), ctypes.c_int(0), null),
        "kzgb_peer_alloc": (z, null, null),
        "kzgb_g1_combine_batch_dev": (null, ctypes.c_uint32(1), ctypes.c_uint32(1), null, ctypes.c_int(0), null),
    }
    for name, args in calls.items():
        assert getattr(lib, name)(*args) < 0, name


def test_scalar_encodings_roundtrip_and_match_oracle():
    for v in (0, 1, 7, O.R_MOD - 1, O.ROOT_OF_UNITY, 2 ** 200 + 12345):
        limbs = scalar.fr(v)
        assert list(map(int, limbs)) == O.fr_to_mont_limbs(v % O.R_MOD)
        assert scalar.fr_int(limbs) == v % O.R_MOD
        assert scalar.fr_to_bytes_le(limbs) == O.fr_to_bytes_le(v)
    assert scalar.ROOT_OF_UNITY == O.ROOT_OF_UNITY
    g = np.array(O.g1_affine_to_blst((O.G1_X, O.G1_Y)), dtype=np.uint64)
    assert scalar.g1_affine_compress(g) == O.g1_compress(O.G1_GENERATOR)
    neg = np.array(O.g1_affine_to_blst((O.G1_X, O.P_MOD - O.G1_Y)), dtype=np.uint64)
    assert scalar.g1_affine_compress(neg)[0] & 0x20
    assert scalar.g1_affine_compress(np.zeros(12, dtype=np.uint64)) == bytes([0xC0]) + bytes(47)


def test_polynomial_degree_rules():
    """polynomial.rs:44-154 and Appendix A.1."""
    p = Polynomial.from_ints([1, 2, 3, 0, 0])
    assert p.degree == 2 and p.num_coeffs() == 3 and len(p.slice_coeffs()) == 3 and len(p.coeffs) == 5
    stale = Polynomial(scalar.fr_vec([1, 2, 3, 0, 0]), 4)
    assert stale.num_coeffs() == 5
    stale.shrink_degree()
    assert stale.degree == 2
    assert Polynomial.new_zero().is_zero() and not Polynomial.from_ints([0, 1]).is_zero()
    assert Polynomial.from_ints([0, 0, 0]).degree == 0
    assert Polynomial.from_ints([1, 2]) == Polynomial(scalar.fr_vec([1, 2, 0]), 1)     # zip stops at the shorter
    assert Polynomial.from_ints([1, 2]) != Polynomial(scalar.fr_vec([1, 2, 0]), 2)     # degree differs
    q = Polynomial.from_ints([5, 1]).sub_constant(scalar.fr(5))
    assert scalar.fr_int(q.coeffs[0]) == 0 and q.degree == 1


def test_compute_omega_and_domain_padding():
    """ft.rs:55-109."""
    assert compute_omega(0) == (1, 0, 1) and compute_omega(1) == (1, 0, 1)
    for d in (2, 5, 1024, 1 << 20):
        assert compute_omega(d) == O.compute_omega(d)
    with pytest.raises(kzg_b200.PolynomialDegreeTooLarge):
        compute_omega((1 << 31) + 1)
    dom = EvaluationDomain.from_coeffs(scalar.fr_vec([1, 2, 3, 4, 5]))
    assert dom.d == 8 and dom.exp == 3 and len(dom) == 8 and not dom.coeffs[5:].any()
    assert dom.minv == pow(8, -1, O.R_MOD) and dom.omegainv == pow(dom.omega, -1, O.R_MOD)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", str(tmp_path / "libkzg_b200.so"))
    with pytest.raises(_ffi.KZGBackendError):
        _ffi.load()
