"""Synthetic fixtures: seeded FCN-ResNet weights written as a genuine opset-12 ``.onnx`` file,
and seeded BGR test frames.

The reference downloads ``fcn-resnet50-12-int8.onnx`` at build time
(infur-test-gen/build.rs:89-91) and renders its clips with ffmpeg's ``testsrc``
(build.rs:12-31); neither a network nor ffmpeg exists here, so this module makes
stand-ins of the same shape: the torchvision ``fcn_resnet50`` architecture (the
network the zoo file is an export of) with random-init weights whose BatchNorm
statistics are calibrated so that activations stay O(1) and class maps are
spatially varied, exported by torch's own exporter, and testsrc-like frames.

This file only *creates input files*; the product path reads them back through
its own C++ ONNX reader (csrc/onnx_reader.cpp) and never touches torch modules.
"""
from __future__ import annotations

import os
import warnings

import numpy as np


# --------------------------------------------------------------------------- frames
def synth_frame(w: int, h: int, index: int = 0, seed: int = 1234) -> np.ndarray:
    """One ``[h][w][3]`` u8 BGR frame: moving colour bars + diagonal gradient + 8x8 checker + noise.

    Stand-in for lavfi ``testsrc`` (infur-test-gen/build.rs:19-23); tight HWC, B,G,R order, the
    layout ``FFMpegDecoder::read_frame`` fills (ff-video/src/decoder.rs:156-165).
    """
    rng = np.random.default_rng(seed + 7919 * index)
    ys, xs = np.mgrid[0:h, 0:w]
    bars = np.array(
        [(255, 255, 255), (0, 255, 255), (255, 255, 0), (0, 255, 0), (255, 0, 255), (0, 0, 255), (255, 0, 0), (16, 16, 16)],
        dtype=np.int32,
    )
    bar_w = max(w // 8, 1)
    bar_idx = ((xs + 3 * index) // bar_w) % 8
    img = bars[bar_idx]
    grad = ((xs + ys + 5 * index) * 255 // max(w + h, 1)).astype(np.int32)
    lower = ys > (h * 2) // 3
    img = np.where(lower[..., None], np.stack([grad, 255 - grad, (grad * 2) % 256], -1), img)
    checker = (((xs // 8) + (ys // 8)) % 2).astype(np.int32) * 24 - 12
    mid = (ys > h // 3) & ~lower
    img = img + np.where(mid, checker, 0)[..., None]
    # a moving disc so consecutive frames differ structurally
    cx, cy = (w // 4 + 11 * index) % max(w, 1), h // 2
    disc = (xs - cx) ** 2 + (ys - cy) ** 2 < (min(w, h) // 6) ** 2
    img = np.where(disc[..., None], np.array([40, 90, 200], dtype=np.int32), img)
    img = img + rng.integers(-8, 9, size=(h, w, 3), dtype=np.int32)
    return np.clip(img, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------- weights
def build_fcn(seed: int = 0, layers=(3, 4, 6, 3), num_classes: int = 21, aux: bool = True, calib_hw=(96, 128), uint8_input: bool = False):
    """torchvision FCN-ResNet (Bottleneck, stride-8 dilated backbone) with seeded synthetic weights.

    ``layers=(3,4,6,3)`` is FCN-ResNet50 (torchvision segmentation/fcn.py:102-114,168);
    ``(1,1,1,1)`` is the small variant the fast tests use (same op set and widths).
    """
    import torch
    from torch import nn
    from torchvision.models.resnet import Bottleneck, ResNet
    from torchvision.models.segmentation.fcn import FCN, FCNHead
    from torchvision.models._utils import IntermediateLayerGetter

    g = torch.Generator().manual_seed(seed)
    backbone = ResNet(Bottleneck, list(layers), replace_stride_with_dilation=[False, True, True])
    ret = {"layer4": "out"}
    if aux:
        ret["layer3"] = "aux"
    body = IntermediateLayerGetter(backbone, return_layers=ret)
    model = FCN(body, FCNHead(2048, num_classes), FCNHead(1024, num_classes) if aux else None)

    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, nn.Conv2d):
                fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
                if name.endswith("bn3"):
                    m.weight.mul_(0.2)
        # the last 1x1 of each head decides the class map: widen it so many classes win somewhere
        for head in (model.classifier, model.aux_classifier):
            if head is not None:
                head[4].weight.mul_(4.0)

        # calibrate BN running statistics on one synthetic frame (one train-mode pass, momentum 1)
        h, w = calib_hw
        bgr = synth_frame(w, h, index=0, seed=4321)
        if uint8_input:   # a Uint8 model sees the raw bytes in B,G,R order (infur/src/predict_onnx.rs:117-122,296-301)
            x = torch.from_numpy(bgr.astype(np.float32)).permute(2, 0, 1)[None]
        else:
            x = torch.from_numpy(bgr[:, :, ::-1].astype(np.float32) / 255.0)
            mean = torch.tensor([0.485, 0.456, 0.406])
            std = torch.tensor([0.229, 0.224, 0.225])
            x = ((x - mean) / std).permute(2, 0, 1)[None]
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.momentum = 1.0
        model.train()
        model(torch.cat([x, x.flip(-1)], 0))
        model.eval()
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.momentum = 0.1
                m.running_var.clamp_(min=1e-3)
    return model


def export_onnx(model, path: str, hw=(64, 64), nhwc: bool = False, uint8_input: bool = False) -> str:
    """Write ``model`` as an opset-12 ONNX file with inputs/outputs named like the zoo file
    (``input`` -> ``out``, ``aux``; infur/src/gui.rs:229-233 prints exactly these names)."""
    import torch
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    # the only step of the legacy exporter that imports the (absent) `onnx` package
    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes
    has_aux = getattr(model, "aux_classifier", None) is not None
    names = ["out", "aux"] if has_aux else ["out"]
    dyn = {n: {0: "batch", 2: "height", 3: "width"} for n in ["input"] + names}
    example = torch.zeros(1, 3, hw[0], hw[1])
    if nhwc or uint8_input:
        # the other input conventions infer_img_pre_proc accepts (predict_onnx.rs:240-262): NHWC layout and / or Uint8 values;
        # ONNX Conv is NCHW float, so the exported graph starts with Transpose / Cast on the input
        class Adapter(torch.nn.Module):
            def __init__(self, net):
                super().__init__()
                self.net = net

            def forward(self, x):
                if nhwc:
                    x = x.permute(0, 3, 1, 2)
                if uint8_input:
                    x = x.float()
                r = self.net(x)
                return (r["out"], r["aux"]) if has_aux else r["out"]

        model = Adapter(model).eval()
        if nhwc:
            example = example.permute(0, 2, 3, 1).contiguous()
            dyn["input"] = {0: "batch", 1: "height", 2: "width"}
        if uint8_input:
            example = example.to(torch.uint8)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp%d" % os.getpid()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(
            model, (example,), tmp, opset_version=12, dynamo=False,
            input_names=["input"], output_names=names, dynamic_axes=dyn,
        )
    os.replace(tmp, path)
    return path


def fixture_path(kind: str = "fcn50", seed: int = 0) -> str:
    root = os.environ.get("INFUR_B200_FIXTURES", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "fixtures"))
    return os.path.join(root, f"{kind}_seed{seed}.onnx")


_LAYERS = {"fcn50": (3, 4, 6, 3), "fcn_tiny": (1, 1, 1, 1), "fcn_tiny_u8_nhwc": (1, 1, 1, 1), "fcn_tiny_f32_nhwc": (1, 1, 1, 1)}


def ensure_fixture(kind: str = "fcn50", seed: int = 0):
    """Return (path, torch_model); the ``.onnx`` file is (re)generated when missing.  ``*_u8_nhwc`` / ``*_f32_nhwc`` are the
    tiny network behind a Uint8 / Float NHWC input (the returned torch model is always the plain NCHW float network)."""
    u8, nhwc = "_u8" in kind, "_nhwc" in kind
    model = build_fcn(seed=seed, layers=_LAYERS[kind], uint8_input=u8)
    path = fixture_path(kind, seed)
    if not os.path.exists(path):
        export_onnx(model, path, nhwc=nhwc, uint8_input=u8)
    return path, model
