"""Host-side encodings for blstrs `Scalar` / `G1Affine` values crossing the C ABI.

Scalars are numpy uint64[4] arrays holding blst_fr limbs (little-endian, Montgomery form R = 2^256 mod r);
vectors are uint64[n, 4].  G1 affine points are uint64[12] (blst_p1_affine; identity = all zero).  These helpers
only convert representations (big-int <-> limbs, compressed bytes); all arithmetic on vectors runs on the GPU."""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
FR_S = 32  # Scalar::S (two-adicity), ft.rs:67
_FR_R = (1 << 256) % R_MOD
_FR_RINV = pow(_FR_R, -1, R_MOD)
_FQ_R = (1 << 384) % P_MOD
_FQ_RINV = pow(_FQ_R, -1, P_MOD)
_M64 = (1 << 64) - 1
# Scalar::root_of_unity(): 7^((r-1)/2^32)  (ft.rs:73)
ROOT_OF_UNITY = pow(7, (R_MOD - 1) >> FR_S, R_MOD)


def fr(v: int) -> np.ndarray:
    """int -> Montgomery limbs (Scalar::from)."""
    m = (v % R_MOD) * _FR_R % R_MOD
    return np.array([(m >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)


def fr_int(limbs) -> int:
    """Montgomery limbs -> canonical int."""
    v = 0
    for i, x in enumerate(np.asarray(limbs, dtype=np.uint64).reshape(4)):
        v |= int(x) << (64 * i)
    return v * _FR_RINV % R_MOD


def fr_vec(values: Iterable[int]) -> np.ndarray:
    vals = list(values)
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for k, v in enumerate(vals):
        out[k] = fr(v)
    return out


def fr_vec_int(arr) -> list:
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [fr_int(row) for row in a]


def fr_neg(limbs) -> np.ndarray:
    return fr(-fr_int(limbs))


def fr_sub(a, b) -> np.ndarray:
    return fr(fr_int(a) - fr_int(b))


def fr_to_bytes_le(limbs) -> bytes:
    """Canonical 32-byte little-endian encoding (blstrs `to_bytes_le`)."""
    return fr_int(limbs).to_bytes(32, "little")


FR_ZERO = np.zeros(4, dtype=np.uint64)
FR_ONE = fr(1)


def _fq_int(limbs) -> int:
    v = 0
    for i, x in enumerate(limbs):
        v |= int(x) << (64 * i)
    return v * _FQ_RINV % P_MOD


def g1_affine_xy(aff) -> Optional[Tuple[int, int]]:
    a = np.asarray(aff, dtype=np.uint64).reshape(12)
    if not a.any():
        return None
    return _fq_int(a[:6]), _fq_int(a[6:])


def g1_affine_compress(aff) -> bytes:
    """48-byte ZCash/IETF compressed encoding (blstrs `G1Affine::to_compressed`)."""
    xy = g1_affine_xy(aff)
    if xy is None:
        return bytes([0xC0]) + bytes(47)
    x, y = xy
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P_MOD - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


G1_IDENTITY_AFFINE = np.zeros(12, dtype=np.uint64)
