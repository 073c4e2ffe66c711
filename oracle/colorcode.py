"""Oracle for ``ColorCode`` (infur/src/decode_predict.rs:9-83) and the display
buffer of ``ProcessingApp::advance`` (infur/src/app.rs:132-144).
Test infrastructure only.

``Color32::from_rgba_unmultiplied`` lives in the un-vendored crate ``epaint``
0.19 (infur/Cargo.toml:18); its published rule is restated here (gamma-space
u8 -> linear f32 -> multiply by alpha/255 -> gamma-space u8, with shortcuts for
alpha 0 and 255).  The reference evaluates the two ``powf`` calls in f32 through
the platform libm, so its low bit is platform dependent; this restatement
evaluates ``pow`` in f64 with the f32-rounded exponents and rounds the result to
f32 (the correctly-rounded ``powf``).  The stated tolerance against any real
reference build is therefore +-1 u8 on the three colour channels; alpha and
the class index are exact.
"""
from __future__ import annotations

import numpy as np

# decode_predict.rs:9-30 -- 20 high-contrast triplets, used as (r, g, b)
COLORS_PALETTE = np.array(
    [
        (75, 180, 60), (75, 25, 230), (25, 225, 255), (200, 130, 0), (48, 130, 245),
        (240, 240, 70), (230, 50, 240), (60, 245, 210), (180, 30, 145), (190, 190, 250),
        (128, 128, 0), (255, 190, 230), (40, 110, 170), (200, 250, 255), (0, 0, 128),
        (195, 255, 170), (0, 128, 128), (180, 215, 255), (128, 0, 0), (128, 128, 128),
    ],
    dtype=np.uint8,
)

_F = np.float32


def _powf(x: np.ndarray, e: np.float32) -> np.ndarray:
    return np.power(x.astype(np.float64), np.float64(e)).astype(np.float32)


def _linear_f32_from_gamma_u8(s: np.ndarray) -> np.ndarray:
    s32 = s.astype(np.float32)
    lo = s32 / _F(3294.6)
    hi = _powf((s32 + _F(14.025)) / _F(269.025), _F(2.4))
    return np.where(s <= 10, lo, hi).astype(np.float32)


def _gamma_u8_from_linear_f32(l: np.ndarray) -> np.ndarray:
    l = l.astype(np.float32)
    inv = _F(1.0) / _F(2.4)
    small = np.floor(_F(3294.6) * l + _F(0.5))
    with np.errstate(invalid="ignore"):
        big = np.floor((_F(269.025) * _powf(np.maximum(l, _F(0)), inv) - _F(14.025)) + _F(0.5))
    out = np.where(l <= 0, _F(0), np.where(l <= _F(0.0031308), small, np.where(l <= 1, big, _F(255))))
    return np.clip(out, 0, 255).astype(np.uint8)


def color32_from_rgba_unmultiplied(r, g, b, a) -> np.ndarray:
    """epaint 0.19 ``Color32::from_rgba_unmultiplied`` -> premultiplied ``[...,4]`` u8 (r,g,b,a)."""
    r, g, b, a = (np.asarray(v, dtype=np.uint8) for v in (r, g, b, a))
    r, g, b, a = np.broadcast_arrays(r, g, b, a)
    a_lin = a.astype(np.float32) / _F(255.0)
    out = np.empty(a.shape + (4,), dtype=np.uint8)
    for i, c in enumerate((r, g, b)):
        pm = _gamma_u8_from_linear_f32(_linear_f32_from_gamma_u8(c) * a_lin)
        out[..., i] = np.where(a == 255, c, np.where(a == 0, 0, pm))
    out[..., 3] = a
    return out


def alpha_u8(conf: np.ndarray) -> np.ndarray:
    """``(alpha * 255.0f32) as u8`` (decode_predict.rs:35): truncate, saturate, NaN -> 0."""
    with np.errstate(invalid="ignore", over="ignore"):
        v = conf.astype(np.float32) * _F(255.0)
    v = np.where(np.isnan(v), _F(0), v)
    return np.clip(np.trunc(v), 0, 255).astype(np.uint8)


def color_code(klass, alpha) -> np.ndarray:
    """``color_code`` (decode_predict.rs:32-36)."""
    klass = np.asarray(klass)
    rgb = COLORS_PALETTE[klass % len(COLORS_PALETTE)]
    return color32_from_rgba_unmultiplied(rgb[..., 0], rgb[..., 1], rgb[..., 2], alpha_u8(np.asarray(alpha, dtype=np.float32)))


def color_lut() -> np.ndarray:
    """``[20][256][4]`` u8: premultiplied colour for every (class % 20, alpha byte)."""
    k = np.arange(20)[:, None]
    a = np.arange(256, dtype=np.uint8)[None, :]
    rgb = COLORS_PALETTE[k]
    return color32_from_rgba_unmultiplied(rgb[..., 0] + 0 * a, rgb[..., 1] + 0 * a, rgb[..., 2] + 0 * a, a + 0 * k.astype(np.uint8))


def argmax_conf(hm: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Per-pixel strict-``>`` scan from ``(k_max, c_max) = (0, 0.0)`` (decode_predict.rs:67-77).

    First maximum wins ties; all-<=0 or NaN columns give class 0 / confidence 0.
    """
    assert hm.ndim == 3
    k, h, w = hm.shape
    k_max = np.zeros((h, w), dtype=np.int64)
    c_max = np.zeros((h, w), dtype=np.float32)
    hm = hm.astype(np.float32, copy=False)
    for i in range(k):
        with np.errstate(invalid="ignore"):
            upd = hm[i] > c_max
        k_max = np.where(upd, i, k_max)
        c_max = np.where(upd, hm[i], c_max)
    return k_max, c_max


def color_code_image(hm: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """``ColorCode::advance`` (decode_predict.rs:53-79): ``[K][H][W]`` f32 -> (class ``[H][W]``, RGBA ``[H][W][4]``)."""
    k_max, c_max = argmax_conf(hm)
    return k_max, color_code(k_max, c_max)


def softmax_confidence(hm: np.ndarray) -> np.ndarray:
    """Opt-in extension (README.md:76 TODO "softmax"; NOT a reference mode -- the reference feeds raw logits to ColorCode):
    per-pixel softmax over the K classes in f32, ``p_k = exp(v_k - max) / sum_j exp(v_j - max)``, the sum taken in class-index
    order.  Feeding the result to ``color_code_image`` gives the first maximum of the logits as the class (every p is > 0, so the
    strict-``>`` scan from (0, 0.0) always moves) and the winner's probability as confidence."""
    hm = hm.astype(np.float32, copy=False)
    m = hm.max(axis=0)
    e = np.exp((hm - m[None]).astype(np.float32)).astype(np.float32)
    s = np.zeros_like(m)
    for k in range(hm.shape[0]):
        s = (s + e[k]).astype(np.float32)
    return (e / s[None]).astype(np.float32)


def frame_rgba(bgr: np.ndarray) -> np.ndarray:
    """``Color32::from_rgb(c[2], c[1], c[0])`` per pixel (app.rs:132-139) -> ``[H][W][4]`` u8, alpha 255."""
    h, w = bgr.shape[:2]
    out = np.empty((h, w, 4), dtype=np.uint8)
    out[..., 0] = bgr[..., 2]
    out[..., 1] = bgr[..., 1]
    out[..., 2] = bgr[..., 0]
    out[..., 3] = 255
    return out


def blend_over(mask_rgba: np.ndarray, frame: np.ndarray) -> np.ndarray:
    """New feature (the reference stacks the two images in the GUI, gui.rs:324-329, "todo: blend
    somehow?"): premultiplied "over" of the mask onto the opaque frame in u8 gamma space,
    ``out = m + (f*(255-a) + 127) / 255`` (integer), alpha 255.  Defined here, not by the reference.
    """
    m = mask_rgba.astype(np.uint32)
    f = frame.astype(np.uint32)
    ia = 255 - m[..., 3:4]
    rgb = m[..., :3] + (f[..., :3] * ia + 127) // 255
    out = np.empty_like(mask_rgba)
    out[..., :3] = np.minimum(rgb, 255).astype(np.uint8)
    out[..., 3] = 255
    return out
