"""Oracle for ``Scale`` (infur/src/processing.rs:142-281).  Test infrastructure only.

The resampling arithmetic itself lives in the un-vendored crate
``fast_image_resize`` (``version = "1"``, Cargo.toml:19; ``ResizeAlg::Nearest``,
processing.rs:189).  Its published nearest rule is restated here: the source
index of destination pixel ``x`` is ``trunc(0.5*s + s*x)`` with
``s = src/dst`` evaluated in f64 (centre-aligned nearest).  PARITY UNPINNED:
the reference tests only pin sizes and the two zero-size errors.
"""
from __future__ import annotations

import numpy as np


class ScaleError(Exception):
    """Mirrors ``ScaleProcError`` (processing.rs:201-211) and ``ValidScaleError`` (:145-157)."""

    def __init__(self, kind: str, msg: str):
        super().__init__(msg)
        self.kind = kind


def valid_scale(value) -> np.float32:
    """``ValidScale::try_from`` (processing.rs:159-168): reject iff ``value <= 0``; NaN passes."""
    v = np.float32(value)
    if v <= np.float32(0.0):
        raise ScaleError("NonPositive", "Cannot scale by negative number")
    return v


def _f32_to_u32_saturating(v: np.float32) -> int:
    """Rust ``f32 as u32``: truncate toward zero, saturate, NaN -> 0."""
    if np.isnan(v):
        return 0
    if v <= 0:
        return 0
    if v >= np.float32(4294967296.0):
        return 4294967295
    return int(v)


def scaled_size(w: int, h: int, factor) -> tuple[int, int]:
    """``nwidth = (w as f32 * factor) as u32`` (processing.rs:253-254)."""
    f = np.float32(factor)
    with np.errstate(over="ignore", invalid="ignore"):
        nw = _f32_to_u32_saturating(np.float32(np.float32(w) * f))
        nh = _f32_to_u32_saturating(np.float32(np.float32(h) * f))
    return nw, nh


def nearest_indices(src: int, dst: int) -> np.ndarray:
    """Centre-aligned nearest source index per destination index, f64 arithmetic.

    ``idx = trunc(0.5*s + s*x)``, ``s = src/dst``; two roundings (mul, add), no FMA.
    """
    s = np.float64(src) / np.float64(dst)
    x = np.arange(dst, dtype=np.float64)
    idx = (np.float64(0.5) * s + s * x).astype(np.int64)
    return np.minimum(idx, src - 1)


def scale_nearest(img: np.ndarray, factor) -> np.ndarray:
    """``Scale::advance`` (processing.rs:232-281) on an ``[H][W][3]`` u8 BGR image.

    factor == 1.0 -> deep copy (:238-241).  Zero-sized input -> ``ZeroSizeIn``,
    zero-sized output -> ``ZeroSizeOut`` (:246-256).
    """
    f = valid_scale(factor)
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
    h, w = img.shape[:2]
    if f == np.float32(1.0):
        return img.copy()
    if w == 0 or h == 0:
        raise ScaleError("ZeroSizeIn", "scaling from 0-sized input")
    nw, nh = scaled_size(w, h, f)
    if nw == 0 or nh == 0:
        raise ScaleError("ZeroSizeOut", "scaling to 0-sized output")
    ys = nearest_indices(h, nh)
    xs = nearest_indices(w, nw)
    return np.ascontiguousarray(img[ys][:, xs])


def scale_bilinear(img: np.ndarray, factor) -> np.ndarray:
    """Opt-in extension, NOT a reference mode (the reference's only mode is Nearest, processing.rs:189; "bilinear" is an
    open TODO in its README:74 and the mode BASELINE.json's north_star names).  Defined here: same size rule and error
    cases as ``scale_nearest``; half-pixel-centre bilinear taps (``oracle.upsample.bilinear_tables``: ``src = s*(dst+0.5)-0.5``
    clamped, f32), value ``ly0*(lx0*a + lx1*b) + ly1*(lx0*c + lx1*d)`` in un-fused f32 in exactly this order, then
    ``u8 = floor(v + 0.5)`` clamped to 0..255.  No antialiasing prefilter (like cv2 INTER_LINEAR)."""
    from .upsample import bilinear_tables

    f = valid_scale(factor)
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
    h, w = img.shape[:2]
    if f == np.float32(1.0):
        return img.copy()
    if w == 0 or h == 0:
        raise ScaleError("ZeroSizeIn", "scaling from 0-sized input")
    nw, nh = scaled_size(w, h, f)
    if nw == 0 or nh == 0:
        raise ScaleError("ZeroSizeOut", "scaling to 0-sized output")
    y0, y1, ly0, ly1 = bilinear_tables(h, nh)
    x0, x1, lx0, lx1 = bilinear_tables(w, nw)
    x = img.astype(np.float32)
    lx0 = lx0[None, :, None]; lx1 = lx1[None, :, None]
    top = (lx0 * x[y0][:, x0]).astype(np.float32) + (lx1 * x[y0][:, x1]).astype(np.float32)
    bot = (lx0 * x[y1][:, x0]).astype(np.float32) + (lx1 * x[y1][:, x1]).astype(np.float32)
    v = (ly0[:, None, None] * top).astype(np.float32) + (ly1[:, None, None] * bot).astype(np.float32)
    return np.clip(np.floor(v + np.float32(0.5)), 0, 255).astype(np.uint8)
