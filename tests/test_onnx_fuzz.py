"""The product's ONNX reader + lowering (csrc/onnx_reader.cpp) must reject damaged files with an error code, never crash:
``Model::control`` keeps the previous model when a load fails (infur/src/predict_onnx.rs:289-308), so a bad file is a
recoverable event for the host.  Mutated copies of small float and quantised graphs go through infur_b200_onnx_describe in
a child process (a crash would take the child down, not the test run).  CPU only."""
import os
import subprocess
import sys

import numpy as np

from infur_b200 import onnx_write as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import ctypes as C, random, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from infur_b200 import _lib as L
from test_oracle_qlinear import _tiny_graph
from test_onnx_fuzz import tiny_float_graph
lib = L.load()
base = [_tiny_graph(), _tiny_graph(mid_quantize=True), tiny_float_graph()]
rnd = random.Random({seed})
buf = C.create_string_buffer(1 << 16); need = C.c_size_t()
path = {path!r}
codes = {{}}
for it in range({count}):
    b = bytearray(rnd.choice(base))
    mode = rnd.randrange(4)
    if mode == 0:
        b = b[: rnd.randrange(len(b))]
    elif mode == 1:
        for _ in range(rnd.randrange(1, 8)):
            b[rnd.randrange(len(b))] = rnd.randrange(256)
    elif mode == 2:
        i = rnd.randrange(len(b)); b[i:i] = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 16)))
    else:
        i = rnd.randrange(len(b)); del b[i:i + rnd.randrange(1, 64)]
    open(path, "wb").write(bytes(b))
    rc = lib.infur_b200_onnx_describe(path.encode(), buf, len(buf), C.byref(need))
    assert rc in (0, L.E_MODEL_LOAD, L.E_MODEL_INPUT_FORMAT), rc
    codes[rc] = codes.get(rc, 0) + 1
print("FUZZ-OK", sorted(codes.items()))
"""


def tiny_float_graph() -> bytes:
    """input -> Conv(3->64 7x7 s2) -> Relu -> Conv(64->64 1x1) -> Resize to the input size."""
    rng = np.random.default_rng(0)
    inits = {
        "w0": rng.standard_normal((64, 3, 7, 7)).astype(np.float32), "b0": np.zeros(64, np.float32),
        "w1": rng.standard_normal((64, 64, 1, 1)).astype(np.float32),
        "c0": np.array([0], np.int64), "c2": np.array([2], np.int64), "c4": np.array([4], np.int64),
    }
    nodes = [
        W.node("Conv", ["input", "w0", "b0"], ["a"], kernel_shape=[7, 7], strides=[2, 2], pads=[3, 3, 3, 3]),
        W.node("Relu", ["a"], ["r"]),
        W.node("Conv", ["r", "w1"], ["y"], kernel_shape=[1, 1]),
        W.node("Shape", ["input"], ["ish"]), W.node("Slice", ["ish", "c2", "c4", "c0"], ["hw"]),
        W.node("Shape", ["y"], ["lsh"]), W.node("Slice", ["lsh", "c0", "c2", "c0"], ["nc"]),
        W.node("Concat", ["nc", "hw"], ["sizes"], axis=0),
        W.node("Resize", ["y", "", "", "sizes"], ["out"], mode="linear", coordinate_transformation_mode="half_pixel"),
    ]
    return W.model(nodes, inits, [W.value_info("input", W.FLOAT, ["n", 3, "h", "w"])], [W.value_info("out", W.FLOAT, ["n", 64, "h", "w"])])


def test_hand_written_float_graph_lowers(tmp_path):
    from test_oracle_qlinear import _describe
    rc, text = _describe(tiny_float_graph(), str(tmp_path))
    assert rc == 0 and text.count(" conv ") == 2 and "relu" in text and "quantised" not in text, text


def test_damaged_files_are_rejected_not_crashed_on(tmp_path):
    code = CHILD.format(root=ROOT, tests=os.path.join(ROOT, "tests"), seed=20240, count=600, path=str(tmp_path / "m.onnx"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FUZZ-OK" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-2000:])


def _float_graph_with(conv_attrs=None, pool_attrs=None, w_shape=(64, 64, 1, 1)) -> bytes:
    rng = np.random.default_rng(1)
    inits = {"w0": rng.standard_normal((64, 3, 7, 7)).astype(np.float32), "w1": np.zeros(w_shape, np.float32),
             "c0": np.array([0], np.int64), "c2": np.array([2], np.int64), "c4": np.array([4], np.int64)}
    ca = dict(kernel_shape=[int(w_shape[2]), int(w_shape[3])])
    ca.update(conv_attrs or {})
    pa = dict(kernel_shape=[3, 3], strides=[2, 2], pads=[1, 1, 1, 1])
    pa.update(pool_attrs or {})
    nodes = [
        W.node("Conv", ["input", "w0"], ["a"], kernel_shape=[7, 7], strides=[2, 2], pads=[3, 3, 3, 3]),
        W.node("Relu", ["a"], ["r"]),
        W.node("MaxPool", ["r"], ["p"], **pa),
        W.node("Conv", ["p", "w1"], ["y"], **ca),
        W.node("Shape", ["input"], ["ish"]), W.node("Slice", ["ish", "c2", "c4", "c0"], ["hw"]),
        W.node("Shape", ["y"], ["lsh"]), W.node("Slice", ["lsh", "c0", "c2", "c0"], ["nc"]),
        W.node("Concat", ["nc", "hw"], ["sizes"], axis=0),
        W.node("Resize", ["y", "", "", "sizes"], ["out"], mode="linear", coordinate_transformation_mode="half_pixel"),
    ]
    return W.model(nodes, inits, [W.value_info("input", W.FLOAT, ["n", 3, "h", "w"])], [W.value_info("out", W.FLOAT, ["n", 64, "h", "w"])])


def test_non_positive_geometry_is_a_load_error_not_a_crash(tmp_path):
    """zero / negative strides, kernel sizes, dilations and negative pads used to reach build_plan's `eh / s` (SIGFPE);
    now they fail the load with INFUR_E_MODEL_LOAD, which leaves the previous model active (predict_onnx.rs:289-308)."""
    from infur_b200 import _lib as L
    from test_oracle_qlinear import _describe
    rc, text = _describe(_float_graph_with(), str(tmp_path))
    assert rc == 0, text
    bad = [
        _float_graph_with(conv_attrs=dict(strides=[0, 0])), _float_graph_with(conv_attrs=dict(strides=[-2, -2])),
        _float_graph_with(conv_attrs=dict(dilations=[0, 0])), _float_graph_with(conv_attrs=dict(pads=[-1, -1, -1, -1])),
        _float_graph_with(pool_attrs=dict(strides=[0, 0])), _float_graph_with(pool_attrs=dict(kernel_shape=[0, 0])),
        _float_graph_with(pool_attrs=dict(pads=[-1, -1, -1, -1])), _float_graph_with(conv_attrs=dict(strides=[1 << 40, 1 << 40])),
    ]
    for i, b in enumerate(bad):
        rc, text = _describe(b, str(tmp_path))
        assert rc == L.E_MODEL_LOAD, (i, rc, text)
