"""ctypes binding of libkzg_b200.so (the C ABI declared in include/kzg_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing or cannot
initialise a B200, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KZGB_LIB") or os.path.join(_HERE, "libkzg_b200.so")   # KZGB_LIB: A/B builds of the library

KZGB_OK = 0
KZGB_POINT_NOT_ON_POLYNOMIAL = 1
KZGB_DEGREE_TOO_LARGE = 2


class KZGBackendError(RuntimeError):
    """Negative status from the C ABI (invalid argument / CUDA failure)."""


_lib = None
_initialised_device = None


def load():
    """dlopen the library (no GPU needed for this)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KZGBackendError(
                f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                "kzg_b200 has no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.kzgb_last_error_message.restype = C.c_char_p
        _lib.kzgb_launch_count.restype = C.c_uint64
    return _lib


def check(status: int) -> int:
    if status < 0:
        msg = load().kzgb_last_error_message().decode(errors="replace")
        raise KZGBackendError(f"kzg_b200 status {status}: {msg}")
    return status


def init(device: int | None = None):
    """kzgb_init on `device` (default: LOCAL_RANK or 0).  One process drives one GPU."""
    global _initialised_device
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised_device is None:
        check(lib.kzgb_init(C.c_int(device)))
        _initialised_device = device
    elif _initialised_device != device:
        raise KZGBackendError(f"already initialised on device {_initialised_device}")
    return lib


def lib():
    if _initialised_device is None:
        return init()
    return _lib


def launch_count() -> int:
    return int(load().kzgb_launch_count())


def ptr(a) -> C.c_void_p:
    """Host numpy array -> void*; int -> device pointer as void*; None -> NULL."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def u64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def microbench(mode: int, blocks_per_sm: int = 4, threads: int = 256, iters: int = 64):
    out = (C.c_double * 3)()
    check(lib().kzgb_microbench(C.c_int(mode), C.c_int(blocks_per_sm), C.c_int(threads), C.c_int(iters), out))
    return {"ops_per_s": out[0], "ops_per_clk_sm": out[1], "ms": out[2]}


def test_fr_op(a, b, op: int) -> np.ndarray:
    a = u64(a, (-1, 4))
    b = None if b is None else u64(b, (-1, 4))
    out = np.empty_like(a)
    check(lib().kzgb_test_fr_op(ptr(a), ptr(b), ptr(out), C.c_size_t(len(a)), C.c_int(op)))
    return out


def test_fp_op(a, b, op: int) -> np.ndarray:
    a = u64(a, (-1, 6))
    b = None if b is None else u64(b, (-1, 6))
    out = np.empty_like(a)
    check(lib().kzgb_test_fp_op(ptr(a), ptr(b), ptr(out), C.c_size_t(len(a)), C.c_int(op)))
    return out


# ---------------------------------------------------------------------------
# SRS handles and MSM
# ---------------------------------------------------------------------------
class SRS:
    """Owns a kzgb_srs* (device-resident affine SRS + window-multiple table)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_jacobian(cls, points, window_bits: int = 0) -> "SRS":
        pts = u64(points, (-1, 18))
        h = C.c_void_p()
        check(lib().kzgb_srs_create_g1(ptr(pts), C.c_size_t(len(pts)), C.c_int(window_bits), C.byref(h)))
        return cls(h)

    @classmethod
    def from_affine(cls, points, window_bits: int = 0) -> "SRS":
        pts = u64(points, (-1, 12))
        h = C.c_void_p()
        check(lib().kzgb_srs_create_g1_affine(ptr(pts), C.c_size_t(len(pts)), C.c_int(window_bits), C.c_int(0),
                                              C.byref(h)))
        return cls(h)

    @classmethod
    def from_device_affine(cls, dev_ptr: int, n: int, window_bits: int = 0) -> "SRS":
        h = C.c_void_p()
        check(lib().kzgb_srs_create_g1_affine(C.c_void_p(dev_ptr), C.c_size_t(n), C.c_int(window_bits), C.c_int(1),
                                              C.byref(h)))
        return cls(h)

    def __len__(self) -> int:
        f = lib().kzgb_srs_size
        f.restype = C.c_size_t
        return int(f(self._h))

    @property
    def window_bits(self) -> int:
        return int(lib().kzgb_srs_window_bits(self._h))

    @property
    def device_bytes(self) -> int:
        f = lib().kzgb_srs_device_bytes
        f.restype = C.c_size_t
        return int(f(self._h))

    @property
    def device_points(self) -> int:
        f = lib().kzgb_srs_device_points
        f.restype = C.c_void_p
        return int(f(self._h) or 0)

    def read_affine(self, start: int = 0, count: int | None = None) -> np.ndarray:
        count = len(self) - start if count is None else count
        out = np.empty((count, 12), dtype=np.uint64)
        check(lib().kzgb_srs_read_affine(self._h, C.c_size_t(start), C.c_size_t(count), ptr(out)))
        return out

    def msm(self, scalars) -> np.ndarray:
        """Host scalars (Montgomery Fr, [m,4] u64) -> affine point (12 u64)."""
        sc = u64(scalars, (-1, 4))
        out = np.empty(12, dtype=np.uint64)
        check(lib().kzgb_msm_g1(self._h, ptr(sc), C.c_size_t(len(sc)), ptr(out)))
        return out

    def msm_jacobian(self, scalars) -> np.ndarray:
        sc = u64(scalars, (-1, 4))
        out = np.empty(18, dtype=np.uint64)
        check(lib().kzgb_msm_g1_jacobian(self._h, ptr(sc), C.c_size_t(len(sc)), ptr(out)))
        return out

    def msm_dev(self, d_scalars: int, m: int, d_out: int, out_mode: int = 0, stream: int = 0) -> None:
        check(lib().kzgb_msm_g1_dev(self._h, C.c_void_p(d_scalars), C.c_size_t(m), C.c_void_p(d_out),
                                    C.c_int(out_mode), C.c_void_p(stream)))

    def msm_batch(self, scalars_list) -> np.ndarray:
        """Several commits in one pipelined call: list of host scalar arrays -> [count, 12] affine points."""
        arrs = [u64(a, (-1, 4)) for a in scalars_list]
        count = len(arrs)
        ptrs = (C.c_void_p * max(count, 1))(*[a.ctypes.data if len(a) else None for a in arrs])
        ms = (C.c_size_t * max(count, 1))(*[len(a) for a in arrs])
        out = np.zeros((count, 12), dtype=np.uint64)
        check(lib().kzgb_msm_g1_batch(self._h, ptrs, ms, C.c_size_t(count), ptr(out)))
        return out

    def msm_batch_dev(self, d_scalars, ms, d_out: int, out_mode: int = 0, stream: int = 0) -> None:
        """Device-resident batch: d_scalars = device pointers, ms = lengths, d_out = device array of results."""
        count = len(d_scalars)
        ptrs = (C.c_void_p * max(count, 1))(*[C.c_void_p(int(p)) for p in d_scalars])
        lens = (C.c_size_t * max(count, 1))(*[int(m) for m in ms])
        check(lib().kzgb_msm_g1_batch_dev(self._h, ptrs, lens, C.c_size_t(count), C.c_void_p(d_out), C.c_int(out_mode),
                                          C.c_void_p(stream)))

    def msm_batch_h2d(self, scalars_list, d_out: int, out_mode: int = 0, stream: int = 0) -> None:
        """Host scalar arrays in, device results out (pipelined copies); returns when the batch is complete."""
        arrs = [u64(a, (-1, 4)) for a in scalars_list]
        count = len(arrs)
        ptrs = (C.c_void_p * max(count, 1))(*[a.ctypes.data if len(a) else None for a in arrs])
        lens = (C.c_size_t * max(count, 1))(*[len(a) for a in arrs])
        check(lib().kzgb_msm_g1_batch_h2d(self._h, ptrs, lens, C.c_size_t(count), C.c_void_p(d_out), C.c_int(out_mode),
                                          C.c_void_p(stream)))

    def close(self):
        if self._h is not None and _lib is not None and _initialised_device is not None:
            _lib.kzgb_srs_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def g1_combine_dev(d_partials: int, count: int, d_out: int, out_mode: int = 0, stream: int = 0) -> None:
    check(lib().kzgb_g1_combine_dev(C.c_void_p(d_partials), C.c_uint32(count), C.c_void_p(d_out), C.c_int(out_mode),
                                    C.c_void_p(stream)))


def g1_combine_batch_dev(d_partials: int, count: int, items: int, d_out: int, out_mode: int = 0, stream: int = 0) -> None:
    check(lib().kzgb_g1_combine_batch_dev(C.c_void_p(d_partials), C.c_uint32(count), C.c_uint32(items), C.c_void_p(d_out),
                                          C.c_int(out_mode), C.c_void_p(stream)))


def srs_generate(kind: int, secret, total: int, first: int = 0, count: int | None = None, window_bits: int = 0) -> SRS:
    """Synthetic SRS from a known secret, built on the device.  kind 0: s^i*G (lib.rs:38-46);
    kind 1: Lagrange L_i(s)*G (eval_form.rs:254-280)."""
    count = total - first if count is None else count
    sec = u64(secret, (4,))
    h = C.c_void_p()
    check(lib().kzgb_srs_generate(C.c_int(kind), ptr(sec), C.c_size_t(total), C.c_size_t(first), C.c_size_t(count),
                                  C.c_int(window_bits), C.byref(h)))
    return SRS(h)


def fr_stream_dev(d_out: int, seed: int, n: int, start: int = 0, mode: int = 0, stream: int = 0) -> None:
    check(lib().kzgb_fr_stream_dev(C.c_void_p(d_out), C.c_uint64(seed), C.c_uint64(start), C.c_size_t(n), C.c_int(mode),
                                   C.c_void_p(stream)))


def profile_enable(on: bool = True) -> None:
    check(lib().kzgb_profile_enable(C.c_int(int(on))))


def profile_read() -> dict:
    """{stage_name: (sum_ms, count)} since the last read (synchronises the device)."""
    l = lib()
    n = int(l.kzgb_profile_stage_count())
    sums = (C.c_double * n)()
    cnts = (C.c_uint64 * n)()
    check(l.kzgb_profile_read(sums, cnts))
    l.kzgb_profile_stage_name.restype = C.c_char_p
    return {l.kzgb_profile_stage_name(i).decode(): (sums[i], int(cnts[i])) for i in range(n) if cnts[i]}


def trace_enable(on: bool = True) -> None:
    check(lib().kzgb_trace_enable(C.c_int(int(on))))


def trace_read(cap: int = 4096):
    """[(source line of the launch, ms since the previous launch mark)] since trace_enable (synchronises)."""
    lines = (C.c_int * cap)()
    ms = (C.c_float * cap)()
    n = check(lib().kzgb_trace_read(lines, ms, C.c_int(cap)))
    return [(int(lines[i]), float(ms[i])) for i in range(n)]


def ntt(a, log_n: int, inverse: bool = False) -> np.ndarray:
    """Host vector ([2^log_n,4] u64 Montgomery Fr) -> transformed copy.  Status 2 raises in the host mirror."""
    a = u64(a, (-1, 4)).copy()
    st = check(lib().kzgb_ntt_fr(ptr(a), C.c_uint32(log_n), C.c_int(int(inverse))))
    return st, a


def ntt_dev(d_inout: int, log_n: int, inverse: bool = False, stream: int = 0) -> int:
    return check(lib().kzgb_ntt_fr_dev(C.c_void_p(d_inout), C.c_uint32(log_n), C.c_int(int(inverse)), C.c_void_p(stream)))


def divide_linear(p, x, y=None):
    """(q, rem) = ((p - y) div (X - x), p(x) - y) on host vectors."""
    p = u64(p, (-1, 4))
    n = len(p)
    q = np.zeros((max(n - 1, 0), 4), dtype=np.uint64)
    rem = np.zeros(4, dtype=np.uint64)
    yy = None if y is None else u64(y, (4,))
    check(lib().kzgb_divide_linear(ptr(p), C.c_size_t(n), ptr(u64(x, (4,))), ptr(yy), ptr(q) if n > 1 else None, ptr(rem)))
    return q, rem


def witness_coeff(srs: SRS, p, x, y):
    """(status, affine) of create_witness; status 1 = PointNotOnPolynomial."""
    p = u64(p, (-1, 4))
    out = np.zeros(12, dtype=np.uint64)
    st = check(lib().kzgb_witness_coeff(srs._h, ptr(p), C.c_size_t(len(p)), ptr(u64(x, (4,))), ptr(u64(y, (4,))), ptr(out)))
    return st, out


def witness_coeff_many(srs: SRS, p, xs, ys):
    """create_witness of one polynomial at m points, pipelined: (status[m] int32, affine[m, 12])."""
    p = u64(p, (-1, 4))
    xs, ys = u64(xs, (-1, 4)), u64(ys, (-1, 4))
    assert len(xs) == len(ys)
    m = len(xs)
    out = np.zeros((m, 12), dtype=np.uint64)
    status = np.zeros(max(m, 1), dtype=np.int32)
    check(lib().kzgb_witness_coeff_many(srs._h, ptr(p), C.c_size_t(len(p)), ptr(xs), ptr(ys), C.c_size_t(m), ptr(out),
                                        status.ctypes.data_as(C.c_void_p)))
    return status[:m], out


def witness_coeff_dev(srs: SRS, d_p: int, n: int, d_xy: int, d_work: int, d_out: int, stream: int = 0) -> int:
    return check(lib().kzgb_witness_coeff_dev(srs._h, C.c_void_p(d_p), C.c_size_t(n), C.c_void_p(d_xy), C.c_void_p(d_work),
                                              C.c_void_p(d_out), C.c_void_p(stream)))


def div_by_omega_i(evals, log_d: int, i: int) -> np.ndarray:
    evals = u64(evals, (-1, 4))
    out = np.zeros_like(evals)
    check(lib().kzgb_div_by_omega_i(ptr(evals), C.c_uint32(log_d), C.c_size_t(i), ptr(out)))
    return out


def witness_eval(srs: SRS, evals, log_d: int, i: int) -> np.ndarray:
    evals = u64(evals, (-1, 4))
    out = np.zeros(12, dtype=np.uint64)
    check(lib().kzgb_witness_eval(srs._h, ptr(evals), C.c_uint32(log_d), C.c_size_t(i), ptr(out)))
    return out


def witness_eval_dev(srs: SRS, d_evals: int, log_d: int, i: int, d_work: int, d_out: int, stream: int = 0) -> int:
    return check(lib().kzgb_witness_eval_dev(srs._h, C.c_void_p(d_evals), C.c_uint32(log_d), C.c_size_t(i),
                                             C.c_void_p(d_work), C.c_void_p(d_out), C.c_void_p(stream)))


def interpolate(xs, ys) -> np.ndarray:
    xs, ys = u64(xs, (-1, 4)), u64(ys, (-1, 4))
    m = len(xs)
    out = np.zeros((max(m, 2), 4), dtype=np.uint64)
    rl = C.c_size_t(0)
    check(lib().kzgb_interpolate(ptr(xs), ptr(ys), C.c_size_t(m), ptr(out), C.byref(rl)))
    return out[: rl.value]


def witness_batched(srs: SRS, p, xs, ys):
    """(status, r_coeffs, w_affine) of create_witness_batched; status 1 = PointNotOnPolynomial."""
    p, xs, ys = u64(p, (-1, 4)), u64(xs, (-1, 4)), u64(ys, (-1, 4))
    m = len(xs)
    r = np.zeros((max(m, 2), 4), dtype=np.uint64)
    w = np.zeros(12, dtype=np.uint64)
    rl = C.c_size_t(0)
    st = check(lib().kzgb_witness_batched(srs._h, ptr(p), C.c_size_t(len(p)), ptr(xs), ptr(ys), C.c_size_t(m), ptr(r),
                                          C.byref(rl), ptr(w)))
    return st, r[: rl.value], w


def witness_batched_dev(srs: SRS, d_p: int, n: int, d_xs: int, d_ys: int, m: int, d_r: int, d_out: int,
                        stream: int = 0) -> int:
    return check(lib().kzgb_witness_batched_dev(srs._h, C.c_void_p(d_p), C.c_size_t(n), C.c_void_p(d_xs),
                                                C.c_void_p(d_ys), C.c_size_t(m), C.c_void_p(d_r), C.c_void_p(d_out),
                                                C.c_void_p(stream)))


def ntt4_columns_dev(d_slab: int, log_n: int, log_n1: int, log_c: int, col_offset: int, inverse: bool, stream: int = 0):
    return check(lib().kzgb_ntt4_columns_dev(C.c_void_p(d_slab), C.c_uint32(log_n), C.c_uint32(log_n1), C.c_uint32(log_c),
                                             C.c_uint32(col_offset), C.c_int(int(inverse)), C.c_void_p(stream)))


def ntt4_unpack_dev(d_recv: int, d_rows: int, world: int, nrows: int, cols: int, stream: int = 0):
    return check(lib().kzgb_ntt4_unpack_dev(C.c_void_p(d_recv), C.c_void_p(d_rows), C.c_uint32(world), C.c_uint32(nrows),
                                            C.c_uint32(cols), C.c_void_p(stream)))


def ntt4_columns_scatter_dev(d_slab: int, peer_rows, rank: int, log_n: int, log_n1: int, log_c: int, col_offset: int,
                             inverse: bool, stream: int = 0):
    """Fused column pass + exchange: results go straight into the peers' row matrices (device pointers, one per rank)."""
    world = len(peer_rows)
    arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_rows])
    return check(lib().kzgb_ntt4_columns_scatter_dev(C.c_void_p(d_slab), arr, C.c_uint32(world), C.c_uint32(rank),
                                                     C.c_uint32(log_n), C.c_uint32(log_n1), C.c_uint32(log_c),
                                                     C.c_uint32(col_offset), C.c_int(int(inverse)), C.c_void_p(stream)))


def peer_alloc(nbytes: int):
    """(device pointer, 64-byte CUDA IPC handle) of a fresh cudaMalloc allocation other ranks can map."""
    p = C.c_void_p()
    h = (C.c_ubyte * 64)()
    check(lib().kzgb_peer_alloc(C.c_size_t(nbytes), C.byref(p), h))
    return int(p.value), bytes(h)


def peer_open(handle: bytes) -> int:
    p = C.c_void_p()
    h = (C.c_ubyte * 64).from_buffer_copy(handle)
    check(lib().kzgb_peer_open(h, C.byref(p)))
    return int(p.value)


def peer_close(d_ptr: int) -> None:
    check(lib().kzgb_peer_close(C.c_void_p(d_ptr)))


def peer_free(d_ptr: int) -> None:
    check(lib().kzgb_peer_free(C.c_void_p(d_ptr)))


def ntt_batch_dev(d_inout: int, d_tmp: int, log_n: int, batch: int, inverse: bool, scale=None, stream: int = 0):
    sc = None if scale is None else u64(scale, (4,))
    return check(lib().kzgb_ntt_batch_dev(C.c_void_p(d_inout), C.c_void_p(d_tmp), C.c_uint32(log_n), C.c_uint32(batch),
                                          C.c_int(int(inverse)), ptr(sc), C.c_void_p(stream)))


def witness_eval_many(srs: SRS, evals, log_d: int, indices) -> np.ndarray:
    """Eval-form create_witness at several indices of one evaluation vector, pipelined -> [m, 12]."""
    evals = u64(evals, (-1, 4))
    idx = [int(i) for i in indices]
    m = len(idx)
    arr = (C.c_size_t * max(m, 1))(*idx)
    out = np.zeros((m, 12), dtype=np.uint64)
    check(lib().kzgb_witness_eval_many(srs._h, ptr(evals), C.c_uint32(log_d), arr, C.c_size_t(m), ptr(out)))
    return out


def srs_lagrange_from_monomial(mono: SRS, log_d: int, window_bits: int = 0) -> SRS:
    """Secret-free Lagrange-basis SRS (eval_form.rs:254-280) as a group inverse NTT of the monomial SRS."""
    h = C.c_void_p()
    check(lib().kzgb_srs_lagrange_from_monomial(mono._h, C.c_uint32(log_d), C.c_int(window_bits), C.byref(h)))
    return SRS(h)


def fr_vec_op(op: int, a, b=None, k=None) -> np.ndarray:
    """op 0: a*b, 1: a-b, 2: a[i]*k^i, 3: a*k on host vectors (returns a new array)."""
    a = u64(a, (-1, 4)).copy()
    bb = None if b is None else u64(b, (-1, 4))
    kk = None if k is None else u64(k, (4,))
    check(lib().kzgb_fr_vec_op(C.c_int(op), ptr(a), ptr(bb), C.c_size_t(len(a)), ptr(kk)))
    return a


def coset_ntt(a, log_n: int, inverse: bool = False):
    a = u64(a, (-1, 4)).copy()
    st = check(lib().kzgb_coset_ntt_fr(ptr(a), C.c_uint32(log_n), C.c_int(int(inverse))))
    return st, a


def poly_mul(a, b) -> np.ndarray:
    """fft_mul pipeline: returns all m = 2^k >= len(a)+len(b) coefficients of a*b."""
    a, b = u64(a, (-1, 4)), u64(b, (-1, 4))
    m = 1
    while m < len(a) + len(b):
        m *= 2
    out = np.zeros((m, 4), dtype=np.uint64)
    mo = C.c_size_t(0)
    check(lib().kzgb_poly_mul(ptr(a), C.c_size_t(len(a)), ptr(b), C.c_size_t(len(b)), ptr(out), C.c_size_t(m), C.byref(mo)))
    return out[: mo.value]


# ---------------------------------------------------------------------------
# wire formats
# ---------------------------------------------------------------------------
def g1_compress(points) -> bytes:
    pts = u64(points, (-1, 12))
    out = np.zeros(48 * len(pts), dtype=np.uint8)
    check(lib().kzgb_g1_compress(ptr(pts), C.c_size_t(len(pts)), ptr(out)))
    return out.tobytes()


def g1_decompress(data: bytes) -> np.ndarray:
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    assert len(buf) % 48 == 0
    n = len(buf) // 48
    out = np.zeros((n, 12), dtype=np.uint64)
    check(lib().kzgb_g1_decompress(ptr(buf), C.c_size_t(n), ptr(out)))
    return out


def srs_from_compressed(data: bytes, window_bits: int = 0, subgroup_check: bool = True) -> SRS:
    """SRS from 48-byte compressed points; checked like blstrs `from_compressed` (subgroup included) unless
    subgroup_check=False (`from_compressed_unchecked`)."""
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    assert len(buf) % 48 == 0
    h = C.c_void_p()
    fn = lib().kzgb_srs_create_g1_compressed if subgroup_check else lib().kzgb_srs_create_g1_compressed_unchecked
    check(fn(ptr(buf), C.c_size_t(len(buf) // 48), C.c_int(window_bits), C.byref(h)))
    return SRS(h)


def g1_subgroup_check(points) -> bool:
    pts = u64(points, (-1, 12))
    ok = C.c_int(0)
    check(lib().kzgb_g1_subgroup_check(ptr(pts), C.c_size_t(len(pts)), C.byref(ok)))
    return bool(ok.value)


def srs_export_compressed(srs: SRS, start: int = 0, count=None) -> bytes:
    count = len(srs) - start if count is None else count
    out = np.zeros(48 * count, dtype=np.uint8)
    check(lib().kzgb_srs_export_compressed(srs._h, C.c_size_t(start), C.c_size_t(count), ptr(out)))
    return out.tobytes()


def fr_to_bytes_le(a) -> bytes:
    a = u64(a, (-1, 4))
    out = np.zeros_like(a)
    check(lib().kzgb_fr_bytes_le(ptr(a), ptr(out), C.c_size_t(len(a)), C.c_int(1)))
    return out.tobytes()


def fr_from_bytes_le(data: bytes) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint64).reshape(-1, 4).copy()
    out = np.zeros_like(a)
    check(lib().kzgb_fr_bytes_le(ptr(a), ptr(out), C.c_size_t(len(a)), C.c_int(0)))
    return out


# ---------------------------------------------------------------------------
# multi-GPU entry points (NCCL inside the library; one process per GPU)
# ---------------------------------------------------------------------------
COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """Rank 0: the 128-byte NCCL id to ship to the other ranks (any transport)."""
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    check(load().kzgb_comm_unique_id(buf))
    return bytes(buf)


def comm_init(rank: int, world: int, uid: bytes | None) -> None:
    buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(uid) if uid is not None else None
    check(lib().kzgb_comm_init(C.c_int(rank), C.c_int(world), buf))


def comm_destroy() -> None:
    check(lib().kzgb_comm_destroy())


def comm_barrier(stream: int = 0) -> None:
    check(lib().kzgb_comm_barrier(C.c_void_p(stream)))


def msm_sharded(srs: SRS, scalars) -> np.ndarray:
    """commit over a point-sharded SRS from host scalars (this rank's slice) -> affine point on every rank."""
    sc = u64(scalars, (-1, 4))
    out = np.empty(12, dtype=np.uint64)
    check(lib().kzgb_msm_g1_sharded(srs._h, ptr(sc), C.c_size_t(len(sc)), ptr(out)))
    return out


def msm_sharded_dev(srs: SRS, d_scalars: int, m: int, d_out: int, out_mode: int = 0, stream: int = 0) -> None:
    check(lib().kzgb_msm_g1_sharded_dev(srs._h, C.c_void_p(d_scalars), C.c_size_t(m), C.c_void_p(d_out), C.c_int(out_mode),
                                        C.c_void_p(stream)))


def msm_sharded_batch(srs: SRS, scalars_list) -> np.ndarray:
    arrs = [u64(a, (-1, 4)) for a in scalars_list]
    count = len(arrs)
    ptrs = (C.c_void_p * max(count, 1))(*[a.ctypes.data if len(a) else None for a in arrs])
    ms = (C.c_size_t * max(count, 1))(*[len(a) for a in arrs])
    out = np.zeros((count, 12), dtype=np.uint64)
    check(lib().kzgb_msm_g1_sharded_batch(srs._h, ptrs, ms, C.c_size_t(count), ptr(out)))
    return out


def msm_sharded_batch_dev(srs: SRS, d_scalars, ms, d_out: int, out_mode: int = 0, stream: int = 0) -> None:
    count = len(d_scalars)
    ptrs = (C.c_void_p * max(count, 1))(*[C.c_void_p(int(p)) for p in d_scalars])
    lens = (C.c_size_t * max(count, 1))(*[int(m) for m in ms])
    check(lib().kzgb_msm_g1_sharded_batch_dev(srs._h, ptrs, lens, C.c_size_t(count), C.c_void_p(d_out), C.c_int(out_mode),
                                              C.c_void_p(stream)))


class Ntt4Plan:
    """kzgb_ntt4_plan: the fused four-step NTT of one 2^log_n vector over the communicator's ranks (collective
    create/close).  Layouts: see include/kzg_b200.h."""

    def __init__(self, log_n: int, log_n1: int = 0):
        h = C.c_void_p()
        check(lib().kzgb_ntt4_plan_create(C.c_uint32(log_n), C.c_uint32(log_n1), C.byref(h)))
        self._h = h
        self.log_n = log_n
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().kzgb_ntt4_plan_shape(self._h, C.byref(a), C.byref(b), C.byref(c)))
        self.l1, self.l2, self.lc = int(a.value), int(b.value), int(c.value)

    def run_dev(self, d_slab: int, inverse: bool = False, stream: int = 0) -> int:
        """Device slab in (only read) -> device pointer of this rank's [N1/world][N2] result block."""
        rows = C.c_void_p()
        check(lib().kzgb_ntt_fr_sharded_dev(self._h, C.c_void_p(d_slab), C.c_int(int(inverse)), C.byref(rows), C.c_void_p(stream)))
        return int(rows.value)

    def run_host(self, vec, inverse: bool = False, gather: bool = True) -> np.ndarray:
        a = u64(vec, (-1, 4))
        assert len(a) == 1 << self.log_n
        out = np.zeros_like(a)
        check(lib().kzgb_ntt_fr_sharded(self._h, ptr(a), ptr(out), C.c_int(int(inverse)), C.c_int(int(gather))))
        return out

    def close(self):
        if self._h is not None:
            check(lib().kzgb_ntt4_plan_destroy(self._h))
            self._h = None


# ---------------------------------------------------------------------------
# device-resident polynomial handles (kzgb_poly)
# ---------------------------------------------------------------------------
class DevicePoly:
    """Owns a kzgb_poly*: a coefficient (or evaluation) vector kept in HBM across calls."""

    def __init__(self, coeffs):
        c = u64(coeffs, (-1, 4))
        h = C.c_void_p()
        check(lib().kzgb_poly_upload(ptr(c), C.c_size_t(len(c)), C.byref(h)))
        self._h = h
        self.n = len(c)

    def download(self) -> np.ndarray:
        out = np.empty((self.n, 4), dtype=np.uint64)
        check(lib().kzgb_poly_download(self._h, ptr(out)))
        return out

    @property
    def device_ptr(self) -> int:
        f = lib().kzgb_poly_device_ptr
        f.restype = C.c_void_p
        return int(f(self._h) or 0)

    def commit(self, srs: SRS, m: int | None = None) -> np.ndarray:
        out = np.empty(12, dtype=np.uint64)
        check(lib().kzgb_poly_commit(srs._h, self._h, C.c_size_t(self.n if m is None else m), ptr(out)))
        return out

    def witness(self, srs: SRS, n: int, x, y):
        out = np.zeros(12, dtype=np.uint64)
        st = check(lib().kzgb_poly_witness(srs._h, self._h, C.c_size_t(n), ptr(u64(x, (4,))), ptr(u64(y, (4,))), ptr(out)))
        return st, out

    def witness_batched(self, srs: SRS, n: int, xs, ys):
        xs, ys = u64(xs, (-1, 4)), u64(ys, (-1, 4))
        m = len(xs)
        r = np.zeros((max(m, 2), 4), dtype=np.uint64)
        w = np.zeros(12, dtype=np.uint64)
        rl = C.c_size_t(0)
        st = check(lib().kzgb_poly_witness_batched(srs._h, self._h, C.c_size_t(n), ptr(xs), ptr(ys), C.c_size_t(m), ptr(r),
                                                   C.byref(rl), ptr(w)))
        return st, r[: rl.value], w

    def witness_eval(self, srs: SRS, log_d: int, i: int) -> np.ndarray:
        out = np.zeros(12, dtype=np.uint64)
        check(lib().kzgb_poly_witness_eval(srs._h, self._h, C.c_uint32(log_d), C.c_size_t(i), ptr(out)))
        return out

    def ntt(self, log_n: int, inverse: bool = False) -> int:
        return check(lib().kzgb_poly_ntt(self._h, C.c_uint32(log_n), C.c_int(int(inverse))))

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.kzgb_poly_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
