"""Generates tests/golden/stages.npz from the oracle (there is no runnable reference in this environment:
no Rust, onnxruntime or ffmpeg -- SURVEY.md §8c -- so these vectors pin the *restatement* against
regressions; the reference's own known-answer tests are ported separately in tests/test_oracle_kat.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from infur_b200 import synth  # noqa: E402

frame = synth.synth_frame(64, 48, 0)
rng = np.random.default_rng(7)
lowres = (rng.standard_normal((21, 6, 8)) * 3).astype(np.float32)
up = oracle.upsample_bilinear(lowres, 48, 64)
k, rgba = oracle.color_code_image(up)
half = oracle.scale_nearest(frame, 0.5)
np.savez_compressed(
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "stages.npz"),
    frame=frame, scaled_half=half, scaled_037=oracle.scale_nearest(frame, 0.37), pre_half=oracle.preprocess_f32(half),
    color_lut=oracle.color_lut(), lowres=lowres, upsampled=up, class_map=k.astype(np.uint8), decoded=rgba,
)
print("wrote stages.npz")

# ---- quantised operators (oracle/qlinear.py): seeded QuantizeLinear -> QLinearConv -> QLinearAdd -> DequantizeLinear
from oracle import qlinear  # noqa: E402

qr = np.random.default_rng(11)
qx = (qr.standard_normal((1, 64, 9, 11)) * 1.5).astype(np.float32)
q_in = qlinear.quantize_linear(qx, np.float32(0.023), np.uint8(121))
qw = qr.integers(-127, 128, size=(64, 64, 3, 3), dtype=np.int8)
qws = (qr.random(64).astype(np.float32) + np.float32(0.5)) * np.float32(0.0007)
qb = qr.integers(-3000, 3000, size=64, dtype=np.int32)
q_conv = qlinear.qlinear_conv(q_in, np.float32(0.023), np.uint8(121), qw, qws, np.zeros(64, np.int8), np.float32(0.031), np.uint8(131), qb, 1, 2, 2)
q_res = qr.integers(0, 256, size=q_conv.shape, dtype=np.uint8)
q_add = qlinear.qlinear_add(q_conv, np.float32(0.031), np.uint8(131), q_res, np.float32(0.019), np.uint8(0), np.float32(0.027), np.uint8(0))
np.savez_compressed(
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "qlinear.npz"),
    x=qx, q_in=q_in, w=qw, w_scale=qws, bias=qb, q_conv=q_conv, q_res=q_res, q_add=q_add,
    deq=qlinear.dequantize_linear(q_add, np.float32(0.027), np.uint8(0)),
)
print("wrote qlinear.npz")
