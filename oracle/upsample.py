"""Oracle for the network's final ``Resize(mode=linear, half_pixel)`` -- the
bilinear upsample FCN applies to its class logits (torchvision
``segmentation/_utils.py``: ``F.interpolate(..., mode="bilinear",
align_corners=False)``; exported as ONNX ``Resize`` and executed inside
``session.run``, infur/src/predict_onnx.rs:138).  Test infrastructure only.

f32 arithmetic, fixed operation order (no FMA):
    scale = in / out                     (f32 division)
    src   = scale * (dst + 0.5) - 0.5    clamped below at 0
    i0 = min(trunc(src), in-1), i1 = min(i0+1, in-1), l1 = src - i0, l0 = 1 - l1
    out = l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d)
"""
from __future__ import annotations

import numpy as np

_F = np.float32


def bilinear_tables(n_in: int, n_out: int):
    """Per destination index: (i0, i1, l0, l1) as (int32, int32, f32, f32)."""
    scale = _F(n_in) / _F(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    src = scale * (dst + _F(0.5)) - _F(0.5)
    src = np.where(src < 0, _F(0), src).astype(np.float32)
    i0 = np.minimum(src.astype(np.int32), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (_F(1.0) - l1).astype(np.float32)
    return i0, i1, l0, l1


def upsample_bilinear(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``[K][h][w]`` f32 -> ``[K][out_h][out_w]`` f32."""
    x = x.astype(np.float32, copy=False)
    k, h, w = x.shape
    y0, y1, ly0, ly1 = bilinear_tables(h, out_h)
    x0, x1, lx0, lx1 = bilinear_tables(w, out_w)
    top = x[:, y0]  # [K][out_h][w]
    bot = x[:, y1]
    t = lx0 * top[:, :, x0] + lx1 * top[:, :, x1]
    b = lx0 * bot[:, :, x0] + lx1 * bot[:, :, x1]
    return (ly0[None, :, None] * t + ly1[None, :, None] * b).astype(np.float32)
