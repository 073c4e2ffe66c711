"""Oracle for the network inside ``session.run`` (infur/src/predict_onnx.rs:138):
FCN-ResNet fp32 forward on the CPU.  Test infrastructure only.

The reference delegates this arithmetic to ONNX Runtime (unpinned git master of
``onnxruntime-rs``, Cargo.toml:20-22) on a model file that is downloaded at
build time; neither is available, so the forward is restated with PyTorch-CPU
fp32 on the *same weights* the product loads from the fixture ``.onnx``
(PARITY UNPINNED: the reference pins only output shapes,
predict_onnx.rs:378-380).  Floating point: the product computes in fp16 with fp32
accumulation, so class maps are compared by exact-match rate with every
mismatch required to be a near-tie of this oracle (see tests/test_gpu_model.py);
``forward_lowres_fp16emu`` additionally emulates the product's rounding points
so that the comparison can be made almost exact.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .colorcode import color_code_image, frame_rgba
from .preprocess import preprocess_f32, preprocess_u8
from .scale import scale_bilinear, scale_nearest
from .upsample import upsample_bilinear


@torch.no_grad()
def forward_lowres(model, x_nchw: np.ndarray, head: str = "out") -> np.ndarray:
    """fp32 logits before the final Resize: ``[N][K][h/8][w/8]``."""
    model.eval()
    x = torch.from_numpy(np.ascontiguousarray(x_nchw))
    feats = model.backbone(x)
    if head == "out":
        y = model.classifier(feats["out"])
    else:
        y = model.aux_classifier(feats["aux"])
    return y.numpy()


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d | None):
    w = conv.weight.detach().clone()
    b = conv.bias.detach().clone() if conv.bias is not None else torch.zeros(w.shape[0])
    if bn is not None:
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w = w * s[:, None, None, None]
        b = (b - bn.running_mean) * s + bn.bias
    return w, b


def _h(x: torch.Tensor) -> torch.Tensor:
    return x.half().float()


@torch.no_grad()
def forward_lowres_fp16emu(model, x_nchw: np.ndarray, head: str = "out") -> np.ndarray:
    """Same network with the product's rounding points emulated: BN folded into the conv,
    weights and every stored activation rounded to fp16, accumulation and bias/residual/ReLU in
    fp32, projection shortcuts summed into conv3 unrounded, final logits kept in fp32."""
    model.eval()
    bb = model.backbone

    def cbr(x, conv, bn, relu=True, res=None, round_out=True):
        w, b = _fold(conv, bn)
        y = F.conv2d(x, _h(w), b, conv.stride, conv.padding, conv.dilation)
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        return _h(y) if round_out else y

    x = _h(torch.from_numpy(np.ascontiguousarray(x_nchw)))
    x = cbr(x, bb.conv1, bb.bn1)
    x = F.max_pool2d(x, 3, 2, 1)
    feats = {}
    for lname in ("layer1", "layer2", "layer3", "layer4"):
        for blk in getattr(bb, lname):
            idt = x
            if blk.downsample is not None:
                # the product fuses the projection shortcut into conv3's accumulator (one f32 sum), so the shortcut
                # is never rounded to fp16 on its own
                idt = cbr(x, blk.downsample[0], blk.downsample[1], relu=False, round_out=False)
            y = cbr(x, blk.conv1, blk.bn1)
            y = cbr(y, blk.conv2, blk.bn2)
            x = cbr(y, blk.conv3, blk.bn3, relu=True, res=idt)
        feats[lname] = x
    if head == "out":
        hd, f = model.classifier, feats["layer4"]
    else:
        hd, f = model.aux_classifier, feats["layer3"]
    y = cbr(f, hd[0], hd[1])
    y = cbr(y, hd[4], None, relu=False, round_out=False)
    return y.numpy()


def pipeline(model, bgr: np.ndarray, factor=1.0, emulate_fp16: bool = False, uint8_input: bool = False, bilinear: bool = False) -> dict:
    """Whole path as sequenced by ``ProcessingApp::advance`` (infur/src/app.rs:107-153):
    Scale -> Model (``out`` head only, :116) -> ColorCode, plus the display buffer (:132-144)."""
    scaled = (scale_bilinear if bilinear else scale_nearest)(bgr, factor)   # bilinear: opt-in extension, not a reference mode
    x = (preprocess_u8(scaled) if uint8_input else preprocess_f32(scaled))[None]   # Uint8 models: raw bytes, B,G,R (predict_onnx.rs:296-301)
    fwd = forward_lowres_fp16emu if emulate_fp16 else forward_lowres
    low = fwd(model, x)[0]
    h, w = scaled.shape[:2]
    logits = upsample_bilinear(low, h, w)
    klass, rgba = color_code_image(logits)
    return {
        "scaled_bgr": scaled, "input_f32": x[0], "lowres": low, "logits": logits,
        "class_map": klass.astype(np.uint8), "decoded_rgba": rgba, "frame_rgba": frame_rgba(scaled),
    }
