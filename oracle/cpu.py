"""ctypes loader for oracle/libkzg_oracle.so (C restatement; TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this.  Arrays are numpy uint64 with the blst layouts
(Fr: [n,4], Fp: [n,6], p1: [n,18], p1_affine: [n,12])."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkzg_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kzg_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def fr_stream(seed: int, n: int, start: int = 0, mode: int = 0) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint64)
    lib().kzgo_fr_stream(_p(out), C.c_uint64(seed), C.c_size_t(start), C.c_size_t(n), C.c_int(mode))
    return out


def fr_from_canonical(a) -> np.ndarray:
    a = _u64(a, (-1, 4))
    out = np.empty_like(a)
    lib().kzgo_fr_from_canonical(_p(out), _p(a), C.c_size_t(len(a)))
    return out


def fr_to_canonical(a) -> np.ndarray:
    a = _u64(a, (-1, 4))
    out = np.empty_like(a)
    lib().kzgo_fr_to_canonical(_p(out), _p(a), C.c_size_t(len(a)))
    return out


def fr_mul_vec(a, b) -> np.ndarray:
    a, b = _u64(a, (-1, 4)), _u64(b, (-1, 4))
    out = np.empty_like(a)
    lib().kzgo_fr_mul_vec(_p(out), _p(a), _p(b), C.c_size_t(len(a)))
    return out


def fp_mul_vec(a, b) -> np.ndarray:
    a, b = _u64(a, (-1, 6)), _u64(b, (-1, 6))
    out = np.empty_like(a)
    lib().kzgo_fp_mul_vec(_p(out), _p(a), _p(b), C.c_size_t(len(a)))
    return out


def fr_inv(a) -> np.ndarray:
    a = _u64(a, (4,))
    out = np.empty_like(a)
    lib().kzgo_fr_inv(_p(out), _p(a))
    return out


def poly_eval(coeffs, x) -> np.ndarray:
    coeffs, x = _u64(coeffs, (-1, 4)), _u64(x, (4,))
    out = np.empty(4, dtype=np.uint64)
    lib().kzgo_poly_eval(_p(out), _p(coeffs), C.c_size_t(len(coeffs)), _p(x))
    return out


def domain_omega(exp: int) -> np.ndarray:
    out = np.empty(4, dtype=np.uint64)
    lib().kzgo_domain_omega(_p(out), C.c_uint32(exp))
    return out


def fft(a, log_n: int, inverse: bool = False, threads: int = 1) -> np.ndarray:
    a = _u64(a, (-1, 4)).copy()
    assert len(a) == 1 << log_n
    st = lib().kzgo_fft(_p(a), C.c_uint32(log_n), C.c_int(int(inverse)), C.c_int(threads))
    assert st == 0
    return a


def long_division(p, deg_p: int, d, deg_d: int):
    """Returns (status, quotient, remainder); status per kzg_cpu.c."""
    p, d = _u64(p, (-1, 4)), _u64(d, (-1, 4))
    q = np.zeros((max(deg_p - deg_d + 1, 1), 4), dtype=np.uint64)
    rem = np.zeros((deg_p + 1, 4), dtype=np.uint64)
    st = lib().kzgo_long_division(_p(q), _p(rem), _p(p), C.c_size_t(deg_p), _p(d), C.c_size_t(deg_d))
    return st, q, rem


def div_by_omega_i(evals, log_d: int, m: int, limit_j=None, do_qm: bool = True) -> np.ndarray:
    evals = _u64(evals, (-1, 4))
    out = np.zeros_like(evals)
    lim = len(evals) if limit_j is None else limit_j
    lib().kzgo_div_by_omega_i(_p(out), _p(evals), C.c_uint32(log_d), C.c_size_t(m), C.c_size_t(lim), C.c_int(int(do_qm)))
    return out


def g1_generator() -> np.ndarray:
    out = np.empty(18, dtype=np.uint64)
    lib().kzgo_g1_generator(_p(out))
    return out


def g1_mul(p, k) -> np.ndarray:
    p, k = _u64(p, (18,)), _u64(k, (4,))
    out = np.empty(18, dtype=np.uint64)
    lib().kzgo_g1_mul(_p(out), _p(p), _p(k))
    return out


def g1_add(a, b) -> np.ndarray:
    a, b = _u64(a, (18,)), _u64(b, (18,))
    out = np.empty(18, dtype=np.uint64)
    lib().kzgo_g1_add(_p(out), _p(a), _p(b))
    return out


def g1_to_affine(p) -> np.ndarray:
    p = _u64(p, (-1, 18))
    out = np.empty((len(p), 12), dtype=np.uint64)
    lib().kzgo_g1_to_affine_batch(_p(out), _p(p), C.c_size_t(len(p)))
    return out


def g1_affine_compress(a) -> bytes:
    a = _u64(a, (12,))
    out = (C.c_uint8 * 48)()
    lib().kzgo_g1_affine_compress(out, _p(a))
    return bytes(out)


def g1_affine_on_curve(a) -> bool:
    return bool(lib().kzgo_g1_affine_on_curve(_p(_u64(a, (12,)))))


def setup_g1(s, n: int) -> np.ndarray:
    out = np.empty((n, 18), dtype=np.uint64)
    lib().kzgo_setup_g1(_p(out), _p(_u64(s, (4,))), C.c_size_t(n))
    return out


def setup_g1_fast(s, n: int, first: int = 0, threads: int = 0) -> np.ndarray:
    """The points of `setup_g1` (gs[i] = s^i*G, i in [first, first+n)) by a threaded fixed-base multiplication:
    benchmark-size SRS for the CPU baseline arm.  Jacobian, NOT normalised (z != 1)."""
    import os
    out = np.zeros((n, 18), dtype=np.uint64)
    lib().kzgo_setup_g1_fast(_p(out), _p(_u64(s, (4,))), C.c_size_t(first), C.c_size_t(n), C.c_int(threads or (os.cpu_count() or 1)))
    return out


def msm_g1(points, scalars, threads: int = 1) -> np.ndarray:
    points, scalars = _u64(points, (-1, 18)), _u64(scalars, (-1, 4))
    n = min(len(points), len(scalars))
    out = np.empty(18, dtype=np.uint64)
    lib().kzgo_msm_g1(_p(out), _p(points), _p(scalars), C.c_size_t(n), C.c_int(threads))
    return out


def commit(gs, coeffs, threads: int = 1) -> np.ndarray:
    gs, coeffs = _u64(gs, (-1, 18)), _u64(coeffs, (-1, 4))
    out = np.empty(12, dtype=np.uint64)
    lib().kzgo_commit(_p(out), _p(gs), _p(coeffs), C.c_size_t(len(coeffs)), C.c_int(threads))
    return out


def create_witness(gs, coeffs, x, y, threads: int = 1):
    gs, coeffs = _u64(gs, (-1, 18)), _u64(coeffs, (-1, 4))
    out = np.zeros(12, dtype=np.uint64)
    st = lib().kzgo_create_witness(_p(out), _p(gs), _p(coeffs), C.c_size_t(len(coeffs) - 1), _p(_u64(x, (4,))),
                                   _p(_u64(y, (4,))), C.c_int(threads))
    return st, out
