"""Quantised-model oracle (oracle/qlinear.py) against the examples of the ONNX operator specification, the fixture
writer -> reader round trip, and the product's C++ lowering of QOperator graphs (CPU only: onnx_describe needs no GPU)."""
import ctypes as C
import os

import numpy as np
import pytest

from infur_b200 import _lib as L
from infur_b200 import onnx_write as W
from infur_b200 import quantize
from oracle import onnx_min, qlinear


def test_quantize_linear_spec_example():
    # ONNX QuantizeLinear example: scale 2, zero point 128
    x = np.array([0, 2, 3, 1000, -254, -1000], dtype=np.float32)
    y = qlinear.quantize_linear(x, np.float32(2), np.uint8(128))
    assert y.tolist() == [128, 129, 130, 255, 1, 0]      # 3 / 2 = 1.5 rounds to even (2)


def test_dequantize_linear_spec_example():
    x = np.array([0, 3, 128, 255], dtype=np.uint8)
    y = qlinear.dequantize_linear(x, np.float32(2), np.uint8(128))
    assert y.tolist() == [-256.0, -250.0, 0.0, 254.0]


def test_qlinear_conv_spec_example_row():
    # first row of the ONNX QLinearConv example (1x1 kernel w = 0, w_zp = 255: y = -(x - 132) * 255 * s + 123 with
    # s = x_scale * w_scale / y_scale = 1 / 255)
    x = np.array([[[[255, 174, 162, 25, 203, 168, 58]]]], dtype=np.uint8)
    w = np.array([[[[0]]]], dtype=np.uint8)
    y = qlinear.qlinear_conv(x, np.float32(0.00369204697), np.uint8(132), w, np.array([0.00172794575], np.float32), np.array([255], np.uint8),
                             np.float32(0.00162681262), np.uint8(123))
    assert y[0, 0, 0].tolist() == [0, 81, 93, 230, 52, 87, 197]


def test_qlinear_conv_padding_bias_and_saturation():
    # 3x3, pad 1: the padded border contributes (x_zp - x_zp) = 0; bias is int32 in units of x_scale * w_scale
    x = np.full((1, 1, 3, 3), 12, dtype=np.uint8)
    w = np.ones((1, 1, 3, 3), dtype=np.int8)
    y = qlinear.qlinear_conv(x, np.float32(1), np.uint8(10), w, np.array([1], np.float32), np.array([0], np.int8), np.float32(1), np.uint8(0),
                             bias=np.array([5], np.int32), pad=1)
    assert y[0, 0].tolist() == [[13, 17, 13], [17, 23, 17], [13, 17, 13]]     # 2 * taps inside + 5
    y = qlinear.qlinear_conv(x, np.float32(1), np.uint8(10), w, np.array([1], np.float32), np.array([0], np.int8), np.float32(0.05), np.uint8(0), pad=1)
    assert y.max() == 255 and y.dtype == np.uint8                               # saturates
    y = qlinear.qlinear_conv(x, np.float32(1), np.uint8(10), -w, np.array([1], np.float32), np.array([0], np.int8), np.float32(1), np.uint8(3), pad=1)
    assert y[0, 0].tolist() == [[0, 0, 0], [0, 0, 0], [0, 0, 0]]                # negative sums clamp at 0 (the folded ReLU)


def test_qlinear_add_hand_computed():
    a = np.array([[[[10, 200, 130]]]], dtype=np.uint8)
    b = np.array([[[[0, 255, 7]]]], dtype=np.uint8)
    # (a - 128) * 0.5 / 0.25 + (b - 0) * 0.125 / 0.25 = 2 (a - 128) + b / 2, + zero point 3
    y = qlinear.qlinear_add(a, np.float32(0.5), np.uint8(128), b, np.float32(0.125), np.uint8(0), np.float32(0.25), np.uint8(3))
    # -236 saturates to 0; 144 + 127.5 saturates to 255; 4 + 3.5 = 7.5 rounds half to even = 8, + 3 = 11
    assert y[0, 0, 0].tolist() == [0, 255, 11]


def _describe(path_or_bytes, tmp_path):
    lib = L.load()
    if isinstance(path_or_bytes, (bytes, bytearray)):
        p = os.path.join(tmp_path, "m.onnx")
        with open(p, "wb") as f:
            f.write(path_or_bytes)
    else:
        p = path_or_bytes
    buf = C.create_string_buffer(1 << 16)
    need = C.c_size_t()
    rc = lib.infur_b200_onnx_describe(p.encode(), buf, len(buf), C.byref(need))
    return rc, buf.value.decode()


def test_fixture_round_trip_and_lowering(tmp_path):
    path = quantize.ensure_fixture("fcn_tiny_int8")
    g = onnx_min.load(path)
    ops = [n.op for n in g.nodes]
    assert ops.count("QLinearConv") == 19 and ops.count("QLinearAdd") == 4 and ops[0] == "QuantizeLinear"
    assert g.inputs == [("input", 1, ["batch", 3, "height", "width"])]
    assert [n.domain for n in g.nodes if n.op == "QLinearAdd"] == ["com.microsoft"] * 4
    rc, text = _describe(path, str(tmp_path))
    assert rc == 0, text
    assert "quantised(zp=" in text and text.count(" conv ") == 19 and text.count(" add[") == 4 and text.count(" deq") == 1
    assert "head out" in text and "classes=21" in text
    assert "plan: int8" in text      # all tensors u8, every non-stem convolution input has zero point 0: runs natively in int8


def _tiny_graph(mid_quantize=False, bad_scale=False, float_conv=False, u8_weights=False, relu=False):
    """input -> QuantizeLinear -> QLinearConv(3->64 7x7 s2) -> DequantizeLinear -> Resize."""
    inits = {
        "xs": np.array(0.02, np.float32), "xz": np.array(100, np.uint8),
        "w": np.ones((64, 3, 7, 7), np.int8), "ws": np.full(64, 0.01, np.float32), "wz": np.zeros(64, np.int8),
        "ys": np.array(0.1, np.float32), "yz": np.array(0, np.uint8), "b": np.zeros(64, np.int32),
        "xs2": np.array(0.03, np.float32),
        "c0": np.array([0], np.int64), "c2": np.array([2], np.int64), "c4": np.array([4], np.int64),
    }
    if u8_weights:   # per-tensor u8 weights around zero point 128, scalar scale: the other weight convention QLinearConv allows
        inits["w"] = np.full((64, 3, 7, 7), 131, np.uint8)
        inits["ws"] = np.array(0.01, np.float32)
        inits["wz"] = np.array(128, np.uint8)
    nodes = [W.node("QuantizeLinear", ["input", "xs", "xz"], ["x"])]
    src = "x"
    if mid_quantize:
        nodes += [W.node("DequantizeLinear", ["x", "xs", "xz"], ["xf"]), W.node("QuantizeLinear", ["xf", "xs", "xz"], ["x2"])]
        src = "x2"
    nodes.append(W.node("QLinearConv", [src, "xs2" if bad_scale else "xs", "xz", "w", "ws", "wz", "ys", "yz", "b"], ["y"], kernel_shape=[7, 7],
                        strides=[2, 2], pads=[3, 3, 3, 3]))
    ysrc = "y"
    if relu:
        nodes.append(W.node("Relu", ["y"], ["yr"]))
        ysrc = "yr"
    nodes += [
        W.node("DequantizeLinear", [ysrc, "ys", "yz"], ["yf"]),
        W.node("Shape", ["input"], ["ish"]), W.node("Slice", ["ish", "c2", "c4", "c0"], ["hw"]),
        W.node("Shape", ["yf"], ["lsh"]), W.node("Slice", ["lsh", "c0", "c2", "c0"], ["nc"]),
        W.node("Concat", ["nc", "hw"], ["sizes"], axis=0),
        W.node("Resize", ["yf", "", "", "sizes"], ["out"], mode="linear", coordinate_transformation_mode="half_pixel"),
    ]
    return W.model(nodes, inits, [W.value_info("input", W.FLOAT, ["n", 3, "h", "w"])], [W.value_info("out", W.FLOAT, ["n", 64, "h", "w"])],
                   opsets=(("", 12), ("com.microsoft", 1)))


def test_lowering_accepts_minimal_quantised_graph(tmp_path):
    rc, text = _describe(_tiny_graph(), str(tmp_path))
    assert rc == 0 and "q[0,255] deq" in text, text
    env = qlinear.run(onnx_min.load(_tiny_graph()), np.zeros((1, 3, 16, 16), np.float32))
    assert env["out"].shape == (1, 64, 16, 16)


def test_lowering_rejects_requantising_edges(tmp_path):
    rc, text = _describe(_tiny_graph(bad_scale=True), str(tmp_path))
    assert rc == L.E_MODEL_LOAD and "differ from its producer" in text, text
    rc, text = _describe(_tiny_graph(mid_quantize=True), str(tmp_path))
    assert rc == L.E_MODEL_LOAD, text


def test_golden_qlinear_vectors():
    """Committed fixtures (tests/golden/make_golden.py) pin the quantised restatement across refactors."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "qlinear.npz"))
    q_in = qlinear.quantize_linear(g["x"], np.float32(0.023), np.uint8(121))
    assert (q_in == g["q_in"]).all()
    q_conv = qlinear.qlinear_conv(q_in, np.float32(0.023), np.uint8(121), g["w"], g["w_scale"], np.zeros(64, np.int8), np.float32(0.031), np.uint8(131),
                                  g["bias"], 1, 2, 2)
    assert (q_conv == g["q_conv"]).all() and 0 < q_conv.min() + 1 and len(np.unique(q_conv)) > 100
    q_add = qlinear.qlinear_add(q_conv, np.float32(0.031), np.uint8(131), g["q_res"], np.float32(0.019), np.uint8(0), np.float32(0.027), np.uint8(0))
    assert (q_add == g["q_add"]).all()
    assert (qlinear.dequantize_linear(q_add, np.float32(0.027), np.uint8(0)) == g["deq"]).all()


def test_quantised_fixture_tracks_its_float_original():
    """Sanity of the fixture quantiser (infur_b200/quantize.py): the int8 stand-in's class map, computed by the integer oracle,
    agrees with the fp32 network it was quantised from on most pixels (random-init weights leave many near-ties)."""
    from infur_b200 import synth
    from oracle import fcn, preprocess_f32
    path, model = synth.ensure_fixture("fcn_tiny")
    g = onnx_min.load(quantize.ensure_fixture("fcn_tiny_int8"))
    bgr = synth.synth_frame(128, 96, 1)
    q = qlinear.run(g, preprocess_f32(bgr)[None])["out"][0].argmax(0)
    f = fcn.pipeline(model, bgr, 1.0)["logits"].argmax(0)
    assert (q == f).mean() > 0.9


def test_lowering_weight_conventions_and_relu(tmp_path):
    rc, text = _describe(_tiny_graph(u8_weights=True), str(tmp_path))
    assert rc == 0 and "q[0,255] deq" in text, text
    env = qlinear.run(onnx_min.load(_tiny_graph(u8_weights=True)), np.ones((1, 3, 16, 16), np.float32))
    assert env["y"].dtype == np.uint8 and env["y"].max() > 0          # (131 - 128) * positive inputs
    rc, text = _describe(_tiny_graph(relu=True), str(tmp_path))
    assert rc == L.E_MODEL_LOAD and "quantised tensor" in text, text


def _two_conv_graph(mid_zp=0, w_dtype=np.int8):
    """stem (3->64 7x7 s2) -> QLinearConv 64->64 1x1 -> DequantizeLinear -> Resize; the tensor between has zero point mid_zp."""
    inits = {
        "xs": np.array(0.02, np.float32), "xz": np.array(100, np.uint8),
        "w0": np.ones((64, 3, 7, 7), np.int8), "ws0": np.full(64, 0.01, np.float32), "wz0": np.zeros(64, np.int8), "b0": np.zeros(64, np.int32),
        "ms": np.array(0.05, np.float32), "mz": np.array(mid_zp, np.uint8),
        "w1": (np.full((64, 64, 1, 1), 250, np.uint8) if w_dtype == np.uint8 else np.ones((64, 64, 1, 1), np.int8)),
        "ws1": np.array(0.01, np.float32), "wz1": np.array(0, w_dtype), "b1": np.zeros(64, np.int32),
        "ys": np.array(0.1, np.float32), "yz": np.array(7, np.uint8),
        "c0": np.array([0], np.int64), "c2": np.array([2], np.int64), "c4": np.array([4], np.int64),
    }
    nodes = [
        W.node("QuantizeLinear", ["input", "xs", "xz"], ["x"]),
        W.node("QLinearConv", ["x", "xs", "xz", "w0", "ws0", "wz0", "ms", "mz", "b0"], ["m"], kernel_shape=[7, 7], strides=[2, 2], pads=[3, 3, 3, 3]),
        W.node("QLinearConv", ["m", "ms", "mz", "w1", "ws1", "wz1", "ys", "yz", "b1"], ["y"], kernel_shape=[1, 1]),
        W.node("DequantizeLinear", ["y", "ys", "yz"], ["yf"]),
        W.node("Shape", ["input"], ["ish"]), W.node("Slice", ["ish", "c2", "c4", "c0"], ["hw"]),
        W.node("Shape", ["yf"], ["lsh"]), W.node("Slice", ["lsh", "c0", "c2", "c0"], ["nc"]),
        W.node("Concat", ["nc", "hw"], ["sizes"], axis=0),
        W.node("Resize", ["yf", "", "", "sizes"], ["out"], mode="linear", coordinate_transformation_mode="half_pixel"),
    ]
    return W.model(nodes, inits, [W.value_info("input", W.FLOAT, ["n", 3, "h", "w"])], [W.value_info("out", W.FLOAT, ["n", 64, "h", "w"])],
                   opsets=(("", 12), ("com.microsoft", 1)))


def test_int8_plan_eligibility_is_reported(tmp_path):
    rc, text = _describe(_two_conv_graph(), str(tmp_path))
    assert rc == 0 and "plan: int8" in text, text
    rc, text = _describe(_two_conv_graph(mid_zp=3), str(tmp_path))      # a convolution input whose zero point is not 0
    assert rc == 0 and "plan: fp16-carried" in text and "zero point 3" in text, text
    rc, text = _describe(_two_conv_graph(w_dtype=np.uint8), str(tmp_path))   # u8 weights 250 with zero point 0 do not fit s8
    assert rc == 0 and "plan: fp16-carried" in text and "do not fit int8" in text, text
    env = qlinear.run(onnx_min.load(_two_conv_graph(mid_zp=3)), np.zeros((1, 3, 16, 16), np.float32))   # the oracle runs either
    assert env["out"].shape == (1, 64, 16, 16)


def test_two_head_quantised_model_lowers(tmp_path):
    """The zoo model has two outputs (`out`, `aux`; infur/src/gui.rs:229-233 prints both): the quantised form of both heads lowers,
    gets an int8 plan, and the oracle produces both."""
    from infur_b200 import synth
    from oracle import preprocess_f32
    data = quantize.quantize_fcn(synth.build_fcn(seed=0, layers=(1, 1, 1, 1)), aux=True)
    rc, text = _describe(data, str(tmp_path))
    assert rc == 0 and "head out" in text and "head aux" in text and text.count(" deq") == 2 and "plan: int8" in text, text
    env = qlinear.run(onnx_min.load(data), preprocess_f32(synth.synth_frame(64, 48, 0))[None])
    assert env["out"].shape == env["aux"].shape == (1, 21, 48, 64)


def test_onnx_writer_reader_round_trip():
    """The fixture writer (infur_b200/onnx_write.py) and the oracle's independent reader (oracle/onnx_min.py) agree on every
    tensor type, attribute kind and graph field the quantised fixtures use."""
    inits = {
        "f": np.arange(6, dtype=np.float32).reshape(2, 3) * 0.5, "u": np.array([0, 7, 255], np.uint8), "s": np.array([-128, 0, 127], np.int8),
        "i": np.array([-(2**31), 5, 2**31 - 1], np.int32), "l": np.array([-3, 2**40], np.int64), "scalar": np.array(0.25, np.float32),
    }
    nodes = [W.node("QLinearAdd", ["a", "b"], ["c"], name="n0", domain="com.microsoft"),
             W.node("Conv", ["c", "f"], ["d"], kernel_shape=[3, 3], strides=[2, 2], group=1, auto_pad="NOTSET", alpha=0.5)]
    data = W.model(nodes, inits, [W.value_info("a", W.FLOAT, ["n", 3, "h", 7])], [W.value_info("d", W.UINT8, [1, 2])],
                   opsets=(("", 12), ("com.microsoft", 1)))
    g = onnx_min.load(data)
    for k, v in inits.items():
        assert g.inits[k].dtype == v.dtype and g.inits[k].shape == v.shape and (g.inits[k] == v).all(), k
    assert [n.op for n in g.nodes] == ["QLinearAdd", "Conv"] and g.nodes[0].domain == "com.microsoft" and g.nodes[0].name == "n0"
    assert g.nodes[1].inputs == ["c", "f"] and g.nodes[1].outputs == ["d"]
    assert g.nodes[1].attrs == {"kernel_shape": [3, 3], "strides": [2, 2], "group": 1, "auto_pad": "NOTSET", "alpha": 0.5}
    assert g.inputs == [("a", 1, ["n", 3, "h", 7])] and g.outputs == [("d", 2, [1, 2])]
