"""Golden vectors from INDEPENDENT implementations (not from oracle/): tests/golden/independent.npz.

The reference's arithmetic (ONNX Runtime) cannot run in this environment, and `stages.npz` / `qlinear.npz` are produced by the
oracle itself (a regression pin).  These vectors come from implementations that share no code with the oracle or the product:

  * `cv2.dnn.readNetFromONNX` (OpenCV 4.13: its own ONNX importer, convolution kernels, Resize) run on the SAME fixture files the
    product loads -- the fp32 FCN (`fcn_tiny_seed0.onnx`) and a QOperator int8 model with the operator set of
    fcn-resnet50-12-int8.onnx (QuantizeLinear / QLinearConv / QLinearAdd / DequantizeLinear);
  * the class map is then derived with plain numpy (strict '>' scan from (0, 0.0), decode_predict.rs:67-77), not oracle code.

Stored per case: the input frame index / size, the class map, the winner's logit and the top-2 margin as float16 (so that a
consumer can tell genuine disagreements from near-ties).  Regenerate with:  python tests/golden/make_golden_independent.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from infur_b200 import quantize, synth  # noqa: E402

W, H = 320, 240


def preprocess(bgr):   # predict_onnx.rs:103-137, plain numpy
    x = bgr[:, :, ::-1].astype(np.float32) * np.float32(1.0) / np.float32(255.0)
    mean = np.array([0.485, 0.456, 0.406], np.float32)
    inv = np.float32(1.0) / np.array([0.229, 0.224, 0.225], np.float32)
    return np.ascontiguousarray(((x - mean) * inv).transpose(2, 0, 1))[None]


def scan(logits):      # decode_predict.rs:67-77: strict '>' from (0, 0.0), first maximum wins
    k_max = np.zeros(logits.shape[1:], np.uint8)
    c_max = np.zeros(logits.shape[1:], np.float32)
    for k in range(logits.shape[0]):
        upd = logits[k] > c_max
        k_max[upd] = k
        c_max[upd] = logits[k][upd]
    srt = np.sort(logits, axis=0)
    margin = srt[-1] - np.maximum(srt[-2], 0.0)
    return k_max, c_max, margin


out = {}
path, _ = synth.ensure_fixture("fcn_tiny")
net = cv2.dnn.readNetFromONNX(path)
for idx in (0, 7):
    frame = synth.synth_frame(W, H, idx)
    net.setInput(preprocess(frame))
    k, c, m = scan(net.forward("out")[0])
    out[f"fp32_frame{idx}_class"] = k
    out[f"fp32_frame{idx}_conf"] = c.astype(np.float16)
    out[f"fp32_frame{idx}_margin"] = np.minimum(m, 60000).astype(np.float16)

# the int8 model: same quantisation as the product's fixture (quantize.ensure_fixture("fcn_tiny_int8")), written with a static
# input shape because cv2.dnn cannot import the dynamic Shape subgraph; weights, scales and zero points are identical
model = synth.build_fcn(seed=0, layers=synth._LAYERS["fcn_tiny"])
static = os.path.join(ROOT, "build", "fixtures", "fcn_tiny_int8_static_%dx%d.onnx" % (W, H))
os.makedirs(os.path.dirname(static), exist_ok=True)
with open(static, "wb") as f:
    f.write(quantize.quantize_fcn(model, static_hw=(H, W)))
qnet = cv2.dnn.readNetFromONNX(static)
for idx in (0, 7):
    frame = synth.synth_frame(W, H, idx)
    qnet.setInput(preprocess(frame))
    k, c, m = scan(qnet.forward("out")[0])
    out[f"int8_frame{idx}_class"] = k
    out[f"int8_frame{idx}_conf"] = c.astype(np.float16)
    out[f"int8_frame{idx}_margin"] = np.minimum(m, 60000).astype(np.float16)
out["meta"] = np.array([W, H, cv2.__version__.encode()], dtype=object)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "independent.npz"), **{k: v for k, v in out.items() if k != "meta"},
                    width=W, height=H, cv2_version=cv2.__version__)
print("wrote independent.npz", {k: v.shape for k, v in out.items() if k != "meta"})
