"""CPU: the C-ABI library loads, exports every symbol include/infur_b200.h declares, parses ONNX files on the
host, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from infur_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "infur_b200.h")).read()
    declared = set(re.findall(r"\b(infur_b200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"infur_b200_config", "infur_b200_out", "infur_b200_slot", "infur_b200_handle", "infur_b200_conv_desc", "infur_b200_result",
                 "infur_b200_device_out"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    assert lib.infur_b200_abi_version() == 2


def test_default_config(lib):
    cfg = L.Config()
    lib.infur_b200_default_config(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(L.Config) and cfg.max_batch == 8 and cfg.ring_depth == 3 and cfg.conv_impl == L.CONV_TCGEN05


def test_create_without_gpu_fails_loudly(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.infur_b200_create(None, C.byref(h))
    assert rc == L.E_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.infur_b200_last_error(None)


def describe(lib, path):
    need = C.c_size_t()
    buf = C.create_string_buffer(1 << 16)
    rc = lib.infur_b200_onnx_describe(path.encode(), buf, len(buf), C.byref(need))
    return rc, buf.value.decode()


def test_onnx_lowering_fcn50(lib, fcn50):
    path, model = fcn50
    rc, text = describe(lib, path)
    assert rc == 0, text
    lines = text.splitlines()
    assert lines[0] == "inputs: input dtype=Float layout=NCHW color=RGB" and lines[1] == "outputs: out aux"
    convs = [l for l in lines if " conv " in l]
    assert len(convs) == 57 and sum("maxpool" in l for l in lines) == 1
    assert sum("+t" in l for l in convs) == 16 and sum(l.endswith("relu") for l in convs) == 51   # 16 residual adds, 51 ReLUs
    assert "3->64 k7 s2 p3 d1 relu" in convs[0]
    assert any("2048->512 k3 s1 p1 d1 relu" in l for l in convs) and any("512->512 k3 s1 p4 d4 relu" in l for l in convs)
    assert lines[-2].startswith("head out") and lines[-1].startswith("head aux") and "classes=21" in lines[-2]


def test_onnx_errors(lib, tmp_path):
    p = tmp_path / "junk.onnx"
    p.write_bytes(b"\x00\x01\x02 definitely not protobuf")
    rc, text = describe(lib, str(p))
    assert rc == L.E_MODEL_LOAD
    rc, text = describe(lib, str(tmp_path / "nope.onnx"))
    assert rc == L.E_MODEL_LOAD and "cannot open" in text


def test_onnx_input_format_errors(lib, tmp_path):
    """infer_img_pre_proc (predict_onnx.rs:223-265): rank != 4 and no colour dimension are input-format errors."""
    import torch

    from infur_b200 import synth

    class M(torch.nn.Module):
        def forward(self, x):
            return torch.relu(x)

    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes
    import warnings

    for shape, msg in (((1, 3, 8), "only 4 dimensions supported"), ((1, 4, 8, 8), "couldn't locate model's color input")):
        f = tmp_path / ("m%d.onnx" % len(shape))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.onnx.export(M(), (torch.zeros(*shape),), str(f), opset_version=12, dynamo=False, input_names=["input"], output_names=["out"])
        rc, text = describe(lib, str(f))
        assert rc == L.E_MODEL_INPUT_FORMAT and msg in text, text


def test_describe_other_input_conventions(lib):
    """The loader accepts the NHWC / Uint8 input conventions of infer_img_pre_proc (predict_onnx.rs:223-265) and reports
    the colour order Model::control would pick (:296-301); no GPU needed."""
    import ctypes as C

    from infur_b200 import synth

    for kind, want in (("fcn_tiny_u8_nhwc", "dtype=Uint8 layout=NHWC color=BGR"), ("fcn_tiny_f32_nhwc", "dtype=Float layout=NHWC color=RGB")):
        path, _ = synth.ensure_fixture(kind)
        need = C.c_size_t()
        lib.infur_b200_onnx_describe(path.encode(), None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value)
        assert lib.infur_b200_onnx_describe(path.encode(), buf, need.value, C.byref(need)) == 0
        assert want in buf.value.decode()


def test_header_is_plain_c99(lib, tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (-pedantic) and a C program must link and run against
    the library without any C++ or CUDA toolchain on the consumer's side."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "cabi.c"
    src.write_text('#include "include/infur_b200.h"\n'
                   "int main(void) { infur_b200_config c; infur_b200_default_config(&c);\n"
                   "  return (infur_b200_abi_version() == 2 && c.max_batch > 0) ? 0 : 1; }\n")
    exe = tmp_path / "cabi"
    libdir = os.path.join(root, "infur_b200", "lib")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", root, str(src), "-o", str(exe), "-L", libdir,
                        "-linfur_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_rust_ffi_matches_header(lib):
    """rust/infur-b200/src/ffi.rs (the Rust shim cannot be compiled here: no cargo) declares the same functions and the same
    struct fields, in the same order, as the ctypes binding -- which test_c99 / the GPU tests check against the header itself."""
    src = open(os.path.join(ROOT, "rust", "infur-b200", "src", "ffi.rs")).read()
    fns = set(re.findall(r"pub fn (infur_b200_[a-z0-9_]+)\(", src))
    assert fns and fns <= set(L.SYMBOLS), fns - set(L.SYMBOLS)
    assert {"infur_b200_create", "infur_b200_scale_control", "infur_b200_model_load", "infur_b200_is_dirty", "infur_b200_advance",
            "infur_b200_submit", "infur_b200_wait"} <= fns
    assert int(re.search(r"ABI_VERSION: i32 = (\d+)", src).group(1)) == L.ABI_VERSION == lib.infur_b200_abi_version()

    def rust_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, src, re.S).group(1)
        return re.findall(r"pub ([a-z0-9_]+):", body)

    for rname, ct in (("Config", L.Config), ("Out", L.Out), ("Result_", L.Result)):
        assert rust_fields(rname) == [f[0] for f in ct._fields_], rname
    # status codes
    for k, v in re.findall(r"pub const (E_[A-Z_]+): i32 = (\d+);", src):
        assert getattr(L, k) == int(v), k


def test_no_contracted_packed_fma_in_the_library(lib):
    """ptxas 12.9 fuses ``mul.rn.f32x2`` + ``add.rn.f32x2`` into one FFMA2 in spite of the explicit rounding modifiers (csrc/ptx.cuh,
    ``mul_f32x2_sep``), which silently drops a rounding of QLinearAdd.  No kernel of the library wants a packed FMA, so its SASS must not
    contain one; the tcgen05 / TMA / TMEM mnemonics the hot path is built on must be there."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "FFMA2" not in sass
    for mnemonic in ("UTCHMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert not re.search(r"\s(HMMA|IMMA)\.", sass)              # no legacy mma.sync path (UTCHMMA / UTCIMMA are the tcgen05 forms)
