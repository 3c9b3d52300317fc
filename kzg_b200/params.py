"""lib.rs: KZGParams, KZGError, setup -- host mirror over the C ABI."""
from __future__ import annotations

import numpy as np

from . import _ffi
from .scalar import fr


class KZGError(Exception):
    """lib.rs:26-36."""


class NoPolynomial(KZGError):
    def __str__(self):
        return "no polynomial!"


class PointNotOnPolynomial(KZGError):
    def __str__(self):
        return "point not on polynomial!"


class BatchOpeningZeroRemainder(KZGError):
    def __str__(self):
        return "batch opening remainder is zero!"


class PolynomialDegreeTooLarge(KZGError):
    def __str__(self):
        return "polynomial degree too large"


class KZGParams:
    """lib.rs:12-19.  `gs` is held as a device-resident SRS handle (affine points + window-multiple table);
    `hs` (G2) is verifier-side and out of scope of this library."""

    def __init__(self, srs: _ffi.SRS):
        self._srs = srs

    @classmethod
    def from_g1_projective(cls, gs, window_bits: int = 0) -> "KZGParams":
        """gs: uint64[n, 18] blst_p1 (the reference's Vec<G1Projective>)."""
        return cls(_ffi.SRS.from_jacobian(gs, window_bits))

    @classmethod
    def from_g1_affine(cls, gs, window_bits: int = 0) -> "KZGParams":
        return cls(_ffi.SRS.from_affine(gs, window_bits))

    @classmethod
    def from_compressed(cls, data: bytes, window_bits: int = 0, subgroup_check: bool = True) -> "KZGParams":
        """G1 powers from a ceremony-style file image: n x 48-byte compressed points (decompressed on the device).
        Checked like blstrs `from_compressed` -- encoding, on-curve and prime-order subgroup; pass
        subgroup_check=False (`from_compressed_unchecked`) only for input validated elsewhere."""
        return cls(_ffi.srs_from_compressed(data, window_bits, subgroup_check))

    def to_compressed(self) -> bytes:
        return _ffi.srs_export_compressed(self._srs)

    def __len__(self) -> int:
        return len(self._srs)

    @property
    def srs(self) -> _ffi.SRS:
        return self._srs

    def gs_affine(self, start: int = 0, count=None) -> np.ndarray:
        return self._srs.read_affine(start, count)


def setup(s, num_coeffs: int, window_bits: int = 0) -> KZGParams:
    """lib.rs:38-55 (G1 half): gs[i] = s^i * G.  The reference builds these by num_coeffs sequential
    scalar multiplications; here the same points come from an Fr power table and a fixed-base
    multiplication on the device."""
    if isinstance(s, int):
        s = fr(s)
    return KZGParams(_ffi.srs_generate(0, s, num_coeffs, window_bits=window_bits))
