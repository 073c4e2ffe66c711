"""Oracle for quantised (QOperator-format) models: integer-exact interpreter of the ONNX graph on the CPU.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Stands in for ``session.run`` (infur/src/predict_onnx.rs:138)
on the kind of model the reference's own tests load, ``fcn-resnet50-12-int8.onnx`` (infur-test-gen/build.rs:89-91,
predict_onnx.rs:350-381).  The arithmetic lives in ONNX Runtime (un-vendored; unpinned git master of onnxruntime-rs,
Cargo.toml:20-22); it is restated here from the published operator definitions (ONNX ``QuantizeLinear`` /
``QLinearConv`` / ``DequantizeLinear`` / ``MaxPool`` / ``Resize``, and the com.microsoft contrib op ``QLinearAdd``):

  QuantizeLinear    y = saturate(rne(x / y_scale) + y_zp)                                  (f32 division)
  QLinearConv       acc = sum (x - x_zp) * (w - w_zp) + B   (int32; padding contributes x == x_zp)
                    y = saturate(rne(f32(acc) * ((x_scale * w_scale[c]) / y_scale)) + y_zp)   (f32, MLAS requantisation)
  QLinearAdd        C = saturate(rne((A_scale / C_scale) * f32(A - A_zp) + (B_scale / C_scale) * f32(B - B_zp)) + C_zp)
  DequantizeLinear  y = f32(x - x_zp) * x_scale

rne = round half to even.  PARITY UNPINNED against ONNX Runtime itself: neither ORT nor the zoo file exists here and the
reference pins only output shapes; the known-answer tests in tests/test_oracle_qlinear.py are the examples of the ONNX
operator specification.  ``run`` also records the largest |accumulator| of every convolution: the product carries the
integers in fp16 operands with an f32 accumulator, which is exact while that stays below 2^24.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import onnx_min
from .upsample import upsample_bilinear

F32 = np.float32


def _qrange(zp: np.ndarray):
    return (-128, 127) if zp.dtype == np.int8 else (0, 255)


def _scalar(a):
    return np.asarray(a).reshape(-1)[0]


def quantize_linear(x: np.ndarray, scale, zp) -> np.ndarray:
    zp = np.asarray(zp)
    lo, hi = _qrange(zp)
    q = np.rint(x.astype(F32) / F32(_scalar(scale))) + F32(int(_scalar(zp)))
    return np.clip(q, lo, hi).astype(zp.dtype)


def dequantize_linear(q: np.ndarray, scale, zp) -> np.ndarray:
    return (q.astype(np.int32) - int(_scalar(zp))).astype(F32) * F32(_scalar(scale))


def requantize(acc: np.ndarray, mult: np.ndarray, zp) -> np.ndarray:
    """acc: integer accumulators [N][C][H][W] (any exact dtype); mult: f32 [C]."""
    zp = np.asarray(zp)
    lo, hi = _qrange(zp)
    v = acc.astype(F32) * mult.astype(F32)[None, :, None, None]
    return np.clip(np.rint(v) + F32(int(_scalar(zp))), lo, hi).astype(zp.dtype)


def qlinear_conv(x, x_scale, x_zp, w, w_scale, w_zp, y_scale, y_zp, bias=None, stride=1, pad=0, dil=1, stats=None):
    """x: [N][C][H][W] u8/s8, w: [O][C][kh][kw] u8/s8; returns the quantised output."""
    xi = torch.from_numpy(x.astype(np.float64) - float(_scalar(x_zp)))     # exact integers; zero padding == x_zp
    wz = np.asarray(w_zp).astype(np.float64).reshape(-1)
    wi = torch.from_numpy(w.astype(np.float64) - (wz[:, None, None, None] if wz.size > 1 else wz[0]))
    acc = F.conv2d(xi, wi, None, stride, pad, dil).numpy()                # f64 sums of integers < 2^53: exact
    if bias is not None:
        acc = acc + bias.astype(np.float64)[None, :, None, None]
    if stats is not None:
        stats.append(float(np.abs(acc).max()))
    ws = np.asarray(w_scale, dtype=F32).reshape(-1)
    mult = (F32(_scalar(x_scale)) * ws) / F32(_scalar(y_scale))
    if mult.size == 1:
        mult = np.repeat(mult, w.shape[0])
    return requantize(acc, mult, y_zp)


def qlinear_add(a, a_scale, a_zp, b, b_scale, b_zp, c_scale, c_zp) -> np.ndarray:
    c_zp = np.asarray(c_zp)
    lo, hi = _qrange(c_zp)
    ra = F32(_scalar(a_scale)) / F32(_scalar(c_scale))
    rb = F32(_scalar(b_scale)) / F32(_scalar(c_scale))
    va = ra * (a.astype(np.int32) - int(_scalar(a_zp))).astype(F32)
    vb = rb * (b.astype(np.int32) - int(_scalar(b_zp))).astype(F32)
    return np.clip(np.rint(va + vb) + F32(int(_scalar(c_zp))), lo, hi).astype(c_zp.dtype)


def max_pool(x: np.ndarray, k: int, s: int, p: int) -> np.ndarray:
    y = F.max_pool2d(torch.from_numpy(x.astype(np.float32)), k, s, p).numpy()   # padding is -inf: never wins
    return y.astype(x.dtype)


def run(graph, x_nchw: np.ndarray) -> dict:
    """Interpret ``graph`` (an ``onnx_min.Graph`` or a path) on one f32 NCHW input.  Returns every named tensor plus
    ``"__max_abs_acc__"`` (largest |int32 accumulator + bias| over all QLinearConv nodes)."""
    g = graph if isinstance(graph, onnx_min.Graph) else onnx_min.load(graph)
    env = dict(g.inits)
    env[g.inputs[0][0]] = np.ascontiguousarray(x_nchw, dtype=F32)
    env[""] = None
    stats = []
    for n in g.nodes:
        i = [env[k] for k in n.inputs]
        a = n.attrs
        if n.op == "QuantizeLinear":
            out = quantize_linear(i[0], i[1], i[2] if len(i) > 2 else np.uint8(0))
        elif n.op == "DequantizeLinear":
            out = dequantize_linear(i[0], i[1], i[2] if len(i) > 2 else np.uint8(0))
        elif n.op == "QLinearConv":
            assert a.get("group", 1) == 1
            st, pd, dl = a.get("strides", [1, 1]), a.get("pads", [0, 0, 0, 0]), a.get("dilations", [1, 1])
            assert st[0] == st[1] and dl[0] == dl[1] and len(set(pd)) == 1
            out = qlinear_conv(i[0], i[1], i[2], i[3], i[4], i[5], i[6], i[7], i[8] if len(i) > 8 else None, st[0], pd[0], dl[0], stats)
        elif n.op == "QLinearAdd":
            out = qlinear_add(*i[:8])
        elif n.op == "MaxPool":
            ks, st, pd = a["kernel_shape"], a.get("strides", [1, 1]), a.get("pads", [0, 0, 0, 0])
            out = max_pool(i[0], ks[0], st[0], pd[0])
        elif n.op == "Relu":
            if i[0].dtype.kind != "f":
                raise NotImplementedError("Relu on a quantised tensor is not valid ONNX (quantisers fold it into the clamp)")
            out = np.maximum(i[0], 0)
        elif n.op == "Shape":
            out = np.array(i[0].shape, dtype=np.int64)
        elif n.op == "Slice":
            starts, ends = i[1], i[2]
            axes = i[3] if len(i) > 3 and i[3] is not None else np.arange(len(starts))
            sl = [slice(None)] * i[0].ndim
            for s0, e0, ax in zip(starts, ends, axes):
                sl[int(ax)] = slice(int(s0), int(e0))
            out = i[0][tuple(sl)]
        elif n.op == "Concat":
            out = np.concatenate(i, axis=a.get("axis", 0))
        elif n.op == "Resize":
            assert a.get("mode") == "linear" and a.get("coordinate_transformation_mode", "half_pixel") in ("half_pixel", "pytorch_half_pixel")
            sizes = [int(v) for v in i[3]]
            out = np.stack([upsample_bilinear(img, sizes[2], sizes[3]) for img in i[0]])
        else:
            raise NotImplementedError(n.op)
        env[n.outputs[0]] = out
    env["__max_abs_acc__"] = max(stats) if stats else 0.0
    return env


def lowres_name(graph) -> str:
    """Name of the tensor feeding the first output's Resize (the de-quantised low-resolution logits)."""
    g = graph if isinstance(graph, onnx_min.Graph) else onnx_min.load(graph)
    out0 = g.outputs[0][0]
    for n in g.nodes:
        if n.op == "Resize" and n.outputs[0] == out0:
            return n.inputs[0]
    raise ValueError("no Resize feeds output '%s'" % out0)
