"""Oracle for the image pre-processing of ``ImageSession::forward``
(infur/src/predict_onnx.rs:97-142).  Test infrastructure only.

Float models (the FCN): channel axis reversed BGR->RGB (:103-107), NHWC->NCHW
(:108-111), ``x = (v as f32 * 1.0) / 255.0`` (:128), then per pixel
``x -= mean[c]`` and ``x *= (1.0f32 / std[c])`` (:131-136) -- three separately
rounded f32 operations, never fused.
"""
from __future__ import annotations

import numpy as np

# ColorNorm::new_torchvision_rgb (predict_onnx.rs:175-180), stored as f32
MEAN_RGB = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD_RGB = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def norm_lut() -> np.ndarray:
    """``[3][256]`` f32 table: channel c (R,G,B order) of byte value v -> normalised input."""
    v = np.arange(256, dtype=np.float32)
    x = (v * np.float32(1.0)) / np.float32(255.0)
    std1 = np.float32(1.0) / STD_RGB
    out = np.empty((3, 256), dtype=np.float32)
    for c in range(3):
        out[c] = (x - MEAN_RGB[c]) * std1[c]
    return out


def preprocess_f32(bgr: np.ndarray) -> np.ndarray:
    """``[H][W][3]`` u8 BGR -> ``[3][H][W]`` f32 RGB-normalised (the tensor handed to ``session.run``)."""
    assert bgr.dtype == np.uint8 and bgr.ndim == 3 and bgr.shape[2] == 3
    rgb = bgr[:, :, ::-1]
    x = (rgb.astype(np.float32) * np.float32(1.0)) / np.float32(255.0)
    x = x - MEAN_RGB  # lane -= mean
    x = x * (np.float32(1.0) / STD_RGB)  # lane *= 1/std
    return np.ascontiguousarray(np.transpose(x, (2, 0, 1)))


def preprocess_u8(bgr: np.ndarray) -> np.ndarray:
    """Uint8 models (predict_onnx.rs:117-122 with ColorSeq::BGR, :296-301): the bytes go to ``session.run`` as they are,
    B,G,R order, only permuted to the model's layout.  Returned as ``[3][H][W]`` f32 (the values the first Cast node
    of such a model produces), channel 0 = B."""
    assert bgr.dtype == np.uint8 and bgr.ndim == 3 and bgr.shape[2] == 3
    return np.ascontiguousarray(np.transpose(bgr.astype(np.float32), (2, 0, 1)))
