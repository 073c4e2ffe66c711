"""Static post-training quantisation of the synthetic FCN-ResNet into a QOperator-format ``.onnx`` file.

Fixture tooling only.  The model the reference's own tests run is ``fcn-resnet50-12-int8.onnx``
(infur-test-gen/build.rs:89-91; infur/src/predict_onnx.rs:350-381), an Intel-Neural-Compressor export of
FCN-ResNet50 in QOperator format: ``QuantizeLinear`` on the input, ``QLinearConv`` (u8 activations, per-channel s8
weights, int32 bias; ReLU folded into the u8 clamp), ``QLinearAdd`` (com.microsoft) for the residual sums, ``MaxPool``
on u8, ``DequantizeLinear`` before the final ``Resize``.  That file cannot be obtained here (no network), so this module
writes stand-ins with exactly that operator set from the seeded networks of ``synth.build_fcn``: activation ranges are
calibrated on synthetic frames (min / max, like ONNX Runtime's static quantiser), weights are quantised symmetrically per
output channel.

The product loads the file through csrc/onnx_reader.cpp; the oracle interprets it with ``oracle/qlinear.py``.
Fidelity of the stand-ins (integer oracle vs the fp32 network they were quantised from, class-map agreement on synthetic
frames): 94-95 % for the tiny network, 85-88 % for FCN-ResNet50 -- random-init weights leave many near-ties; the point of
the fixtures is the operator set and realistic value ranges, not accuracy.
"""
from __future__ import annotations

import os

import numpy as np

from . import onnx_write as W
from . import synth


def _fold(conv, bn):
    import torch
    w = conv.weight.detach().clone()
    b = conv.bias.detach().clone() if conv.bias is not None else torch.zeros(w.shape[0])
    if bn is not None:
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w = w * s[:, None, None, None]
        b = (b - bn.running_mean) * s + bn.bias
    return w, b


def _layers(model, aux: bool):
    """The network as a list of steps: ("conv", name, in, out, w, b, stride, pad, dil, relu) / ("pool", in, out) /
    ("add", a, b, out) / ("head", tensor, output name)."""
    bb = model.backbone
    steps = []

    def conv(name, src, dst, c, bn, relu):
        w, b = _fold(c, bn)
        steps.append(("conv", name, src, dst, w, b, c.stride[0], c.padding[0], c.dilation[0], relu))

    conv("stem", "x", "stem", bb.conv1, bb.bn1, True)
    steps.append(("pool", "stem", "pool"))
    cur = "pool"
    feats = {}
    for lname in ("layer1", "layer2", "layer3", "layer4"):
        for bi, blk in enumerate(getattr(bb, lname)):
            p = f"{lname}_{bi}"
            idt = cur
            if blk.downsample is not None:
                conv(p + "_down", cur, p + "_down", blk.downsample[0], blk.downsample[1], False)
                idt = p + "_down"
            conv(p + "_c1", cur, p + "_c1", blk.conv1, blk.bn1, True)
            conv(p + "_c2", p + "_c1", p + "_c2", blk.conv2, blk.bn2, True)
            conv(p + "_c3", p + "_c2", p + "_c3", blk.conv3, blk.bn3, False)
            steps.append(("add", p + "_c3", idt, p + "_out"))
            cur = p + "_out"
        feats[lname] = cur
    heads = [("out", model.classifier, feats["layer4"])]
    if aux and model.aux_classifier is not None:
        heads.append(("aux", model.aux_classifier, feats["layer3"]))
    for hname, hd, src in heads:
        conv(hname + "_h0", src, hname + "_h0", hd[0], hd[1], True)
        conv(hname + "_h4", hname + "_h0", hname + "_low", hd[4], None, False)
        steps.append(("head", hname + "_low", hname))
    return steps


def _float_forward(steps, x):
    """fp32 forward over the step list; returns every tensor (post-ReLU where the ReLU is fused)."""
    import torch
    import torch.nn.functional as F
    t = {"x": x}
    for s in steps:
        if s[0] == "conv":
            _, _, src, dst, w, b, stride, pad, dil, relu = s
            y = F.conv2d(t[src], w, b, stride, pad, dil)
            t[dst] = F.relu(y) if relu else y
        elif s[0] == "pool":
            t[s[2]] = F.max_pool2d(t[s[1]], 3, 2, 1)
        elif s[0] == "add":
            t[s[3]] = F.relu(t[s[1]] + t[s[2]])
    return t


def _u8_params(lo: float, hi: float):
    """Asymmetric u8 (scale, zero point) covering [min(lo, 0), max(hi, 0)] -- the rule of ONNX Runtime's static quantiser."""
    lo, hi = min(float(lo), 0.0), max(float(hi), 0.0)
    scale = np.float32(max((hi - lo) / 255.0, 1e-8))
    zp = int(np.clip(np.rint(-lo / float(scale)), 0, 255))
    return scale, zp


def _preprocess(bgr: np.ndarray) -> np.ndarray:
    # ImageSession::forward's normalisation (predict_onnx.rs:126-137), f32 op for op
    x = bgr[:, :, ::-1].astype(np.float32) * np.float32(1.0) / np.float32(255.0)
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32)
    inv = np.float32(1.0) / np.array([0.229, 0.224, 0.225], dtype=np.float32)
    return np.ascontiguousarray(((x - mean) * inv).transpose(2, 0, 1))


def quantize_fcn(model, calib_hw=(96, 128), n_calib: int = 3, aux: bool = False, static_hw=None) -> bytes:
    """Return the bytes of a QOperator-format ONNX model of ``model`` (a ``synth.build_fcn`` network).

    ``static_hw=(h, w)``: fixed input shape ``[1, 3, h, w]`` and a constant ``sizes`` input of the final Resize instead of the
    Shape/Slice/Concat subgraph (importers without dynamic-shape support, e.g. cv2.dnn, need this)."""
    import torch

    model.eval()
    with torch.no_grad():
        steps = _layers(model, aux)
        h, w = calib_hw
        xs = np.stack([_preprocess(synth.synth_frame(w, h, index=i, seed=977)) for i in range(n_calib)])
        t = _float_forward(steps, torch.from_numpy(xs))
        rng = {k: (float(v.min()), float(v.max())) for k, v in t.items()}

    q = {k: _u8_params(*v) for k, v in rng.items()}      # tensor -> (scale f32, zero point)
    q["pool"] = q["stem"]                                  # MaxPool keeps its input's quantisation
    inits, nodes = {}, []

    def scalar(name, scale, zp):
        inits[name + "_scale"] = np.array(scale, dtype=np.float32)
        inits[name + "_zp"] = np.array(zp, dtype=np.uint8)
        return [name + "_scale", name + "_zp"]

    for k, (s, z) in q.items():
        scalar(k, s, z)
    nodes.append(W.node("QuantizeLinear", ["input", "x_scale", "x_zp"], ["x"], name="quantize_input"))
    for st in steps:
        if st[0] == "conv":
            _, name, src, dst, wt, b, stride, pad, dil, _relu = st
            wt = wt.numpy().astype(np.float64)
            amax = np.maximum(np.abs(wt).reshape(wt.shape[0], -1).max(1), 1e-12)
            w_scale = (amax / 127.0).astype(np.float32)
            wq = np.clip(np.rint(wt / w_scale.astype(np.float64)[:, None, None, None]), -127, 127).astype(np.int8)
            x_scale = q[src][0]
            bq = np.rint(b.numpy().astype(np.float64) / (np.float64(x_scale) * w_scale.astype(np.float64))).astype(np.int64)
            bq = np.clip(bq, -(2**24) + 1, 2**24 - 1).astype(np.int32)
            inits[name + "_w"] = wq
            inits[name + "_w_scale"] = w_scale
            inits[name + "_w_zp"] = np.zeros(wt.shape[0], dtype=np.int8)
            inits[name + "_b"] = bq
            k = wt.shape[2]
            nodes.append(W.node(
                "QLinearConv",
                [src, src + "_scale", src + "_zp", name + "_w", name + "_w_scale", name + "_w_zp", dst + "_scale", dst + "_zp", name + "_b"],
                [dst], name=name, kernel_shape=[k, k], strides=[stride, stride], pads=[pad] * 4, dilations=[dil, dil], group=1))
        elif st[0] == "pool":
            nodes.append(W.node("MaxPool", [st[1]], [st[2]], name="maxpool", kernel_shape=[3, 3], strides=[2, 2], pads=[1, 1, 1, 1]))
        elif st[0] == "add":
            _, a, bb_, out = st
            nodes.append(W.node("QLinearAdd", [a, a + "_scale", a + "_zp", bb_, bb_ + "_scale", bb_ + "_zp", out + "_scale", out + "_zp"],
                                [out], name=out, domain="com.microsoft"))
        elif st[0] == "head":
            _, low, hname = st
            nodes.append(W.node("DequantizeLinear", [low, low + "_scale", low + "_zp"], [low + "_f"], name=hname + "_dequantize"))
            if static_hw is not None:
                kk = int(model.classifier[4].weight.shape[0])
                inits[hname + "_sizes"] = np.array([1, kk, static_hw[0], static_hw[1]], dtype=np.int64)
                nodes.append(W.node("Resize", [low + "_f", "", "", hname + "_sizes"], [hname], name=hname + "_resize", mode="linear",
                                    coordinate_transformation_mode="half_pixel"))
                continue
            # resize to the network input's H x W: sizes = concat(shape(low)[0:2], shape(input)[2:4])
            for nm, v in (("c0", [0]), ("c2", [2]), ("c4", [4])):
                inits[nm] = np.array(v, dtype=np.int64)
            nodes.append(W.node("Shape", ["input"], [hname + "_ishape"]))
            nodes.append(W.node("Slice", [hname + "_ishape", "c2", "c4", "c0"], [hname + "_hw"]))
            nodes.append(W.node("Shape", [low + "_f"], [hname + "_lshape"]))
            nodes.append(W.node("Slice", [hname + "_lshape", "c0", "c2", "c0"], [hname + "_nc"]))
            nodes.append(W.node("Concat", [hname + "_nc", hname + "_hw"], [hname + "_sizes"], axis=0))
            nodes.append(W.node("Resize", [low + "_f", "", "", hname + "_sizes"], [hname], name=hname + "_resize", mode="linear",
                                coordinate_transformation_mode="half_pixel"))
    heads = [s[2] for s in steps if s[0] == "head"]
    k = int(model.classifier[4].weight.shape[0])
    return W.model(
        nodes, inits,
        [W.value_info("input", W.FLOAT, ["batch", 3, "height", "width"] if static_hw is None else [1, 3, static_hw[0], static_hw[1]])],
        [W.value_info(hn, W.FLOAT, ["batch", k, "height", "width"] if static_hw is None else [1, k, static_hw[0], static_hw[1]]) for hn in heads],
        opsets=(("", 12), ("com.microsoft", 1)), producer="infur_b200.quantize")


def ensure_fixture(kind: str = "fcn_tiny_int8", seed: int = 0) -> str:
    """``fcn_tiny_int8`` / ``fcn50_int8``: the quantised stand-in for the zoo's int8 file; (re)generated when missing."""
    base = kind[: -len("_int8")]
    path = synth.fixture_path(kind, seed)
    if not os.path.exists(path):
        model = synth.build_fcn(seed=seed, layers=synth._LAYERS[base])
        data = quantize_fcn(model)
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, path)
    return path
