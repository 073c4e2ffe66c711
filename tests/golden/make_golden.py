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
