"""GPU: quantised (QOperator-format) models.  Integer work: the bar is bit-exact.

The product keeps quantised activations as the integers q - zero_point in fp16, the weights as w_q - w_zero_point, and
requantises in the convolution epilogue (csrc/conv_tc.cu); while |accumulator| < 2^24 the fp16 x fp16 -> f32 tensor-core
sum IS the int32 accumulator, so every QLinearConv / QLinearAdd output must equal the oracle's (oracle/qlinear.py) exactly,
through every kernel variant.  Whole model: the low-resolution logits (after DequantizeLinear) are compared bit for bit;
behind them the path is the float one (bilinear Resize in f32 -> ColorCode), held to the same bar as test_gpu_pipeline.py.
"""
import numpy as np
import pytest

import oracle
from infur_b200 import _lib as L
from infur_b200 import processors as P
from infur_b200 import quantize, synth
from oracle import onnx_min, qlinear

pytestmark = pytest.mark.gpu

# n, h, w, cin, cout, k, stride, pad, dil, residual, impl, f32 head
QCASES = [
    (1, 8, 16, 64, 64, 1, 1, 0, 1, False, L.CONV_TCGEN05, False),
    (2, 30, 40, 64, 256, 1, 1, 0, 1, True, L.CONV_TCGEN05, False),
    (1, 30, 40, 256, 512, 1, 1, 0, 1, True, L.CONV_TCGEN05_PAIR, False),
    (1, 30, 40, 512, 256, 1, 1, 0, 1, False, L.CONV_TCGEN05_PAIR, False),
    (1, 30, 40, 64, 64, 3, 1, 1, 1, False, L.CONV_TCGEN05, False),
    (1, 30, 40, 64, 128, 3, 1, 1, 1, False, L.CONV_TCGEN05_HALO, False),
    (1, 33, 47, 128, 128, 3, 1, 2, 2, False, L.CONV_TCGEN05_HALO, False),
    (1, 60, 80, 128, 128, 3, 2, 1, 1, False, L.CONV_TCGEN05, False),
    (1, 60, 80, 256, 512, 1, 2, 0, 1, False, L.CONV_TCGEN05, False),
    (2, 61, 75, 3, 64, 7, 2, 3, 1, False, L.CONV_TCGEN05, False),
    (1, 30, 40, 512, 21, 1, 1, 0, 1, False, L.CONV_TCGEN05, True),
    (1, 30, 40, 2048, 512, 3, 1, 1, 1, False, L.CONV_TCGEN05, False),
]


def _case_id(c):
    return "n%d_%dx%d_c%d-%d_k%d_s%d_p%d_d%d%s_impl%d%s" % (c[:9] + ("_add" if c[9] else "", c[10], "_deq" if c[11] else ""))


@pytest.mark.parametrize("case", QCASES, ids=_case_id)
def test_qlinear_conv_bit_exact(handle, case):
    _run_qcase(handle, case, x_zp=117)


# the same layers as an int8 plan runs them: u8 tensors in HBM, tcgen05.mma.kind::i8 with s32 accumulators (the RGB stem keeps
# fp16-carried operands and writes u8).  Input zero point 0, as for every post-ReLU tensor of the network.
I8CASES = [c[:10] + (L.CONV_TCGEN05_I8, c[11]) for c in QCASES if c[10] == L.CONV_TCGEN05] + [
    (1, 30, 40, 256, 512, 1, 1, 0, 1, True, L.CONV_TCGEN05_I8, False),
    (1, 33, 47, 128, 128, 3, 1, 2, 2, False, L.CONV_TCGEN05_I8, False),
    (2, 17, 23, 1024, 256, 1, 1, 0, 1, False, L.CONV_TCGEN05_I8, False),
    (1, 30, 40, 256, 512, 1, 1, 0, 1, True, L.CONV_TCGEN05_I8_PAIR, False),
    (1, 30, 40, 512, 256, 1, 1, 0, 1, False, L.CONV_TCGEN05_I8_PAIR, False),
    (1, 33, 47, 256, 256, 3, 1, 2, 2, False, L.CONV_TCGEN05_I8_PAIR, False),
    (2, 31, 45, 512, 512, 3, 1, 4, 4, False, L.CONV_TCGEN05_I8_PAIR, False),
    (1, 30, 40, 256, 512, 1, 1, 0, 1, True, L.CONV_TCGEN05_I8_PAIR_DEEP, False),    # eight epilogue chunk buffers (round 2)
    (3, 17, 23, 64, 256, 1, 1, 0, 1, True, L.CONV_TCGEN05_I8_PAIR_DEEP, False),     # small K: |acc| < 2^22, the I2F-free conversion
]


@pytest.mark.parametrize("case", I8CASES, ids=_case_id)
def test_qlinear_conv_native_int8_bit_exact(handle, case):
    _run_qcase(handle, case, x_zp=117 if case[3] == 3 else 0)


def _run_qcase(handle, case, x_zp):
    n, h, w, cin, cout, k, stride, pad, dil, res, impl, f32 = case
    rng = np.random.default_rng(abs(hash(case[:10])) % 2**31)
    i8 = impl in (L.CONV_TCGEN05_I8, L.CONV_TCGEN05_I8_PAIR, L.CONV_TCGEN05_I8_PAIR_DEEP)
    y_zp, r_zp, c_zp = 128 if (res or f32) else 0, 37 if i8 else 0, 0
    xq = rng.integers(0, 256, size=(n, cin, h, w), dtype=np.uint8)
    wq = rng.integers(-127, 128, size=(cout, cin, k, k), dtype=np.int8)
    bq = rng.integers(-20000, 20000, size=cout, dtype=np.int32)
    x_scale, y_scale = np.float32(0.021), np.float32(0.043)
    w_scale = (rng.random(cout).astype(np.float32) + np.float32(0.5)) * np.float32(y_scale / x_scale / (40.0 * np.sqrt(cin * k * k)))
    stats = []
    yq = qlinear.qlinear_conv(xq, x_scale, np.uint8(x_zp), wq, w_scale, np.zeros(cout, np.int8), y_scale, np.uint8(y_zp), bq, stride, pad, dil, stats)
    assert stats[0] < (2**31 if i8 else 2**24)   # fp16-carried operands: exact below 2^24; native int8: int32
    quant = {"qmul": (x_scale * w_scale) / y_scale, "q_lo": -y_zp, "q_hi": 255 - y_zp, "q_zout": y_zp}
    expect = yq.astype(np.int32) - y_zp
    rq = None
    if res:
        rq = rng.integers(0, 256, size=yq.shape, dtype=np.uint8)
        r_scale, c_scale = np.float32(0.017), np.float32(0.031)
        cq = qlinear.qlinear_add(yq, y_scale, np.uint8(y_zp), rq, r_scale, np.uint8(r_zp), c_scale, np.uint8(c_zp))
        quant.update(q_ra=y_scale / c_scale, q_rb=r_scale / c_scale, q_lo2=-c_zp, q_hi2=255 - c_zp, q_zres=r_zp, q_zout=c_zp)
        expect = cq.astype(np.int32) - c_zp
    if f32:
        quant["q_deq"] = y_scale
        expect = qlinear.dequantize_linear(yq, y_scale, np.uint8(y_zp))
    xc = (xq.astype(np.int32) - x_zp).transpose(0, 2, 3, 1)              # NHWC, centred: what the engine stores
    wc = wq.astype(np.int32).transpose(0, 2, 3, 1)
    rc = (rq.astype(np.int32) - r_zp).transpose(0, 2, 3, 1) if res else None
    y = handle.conv_test(xc, wc, bq.astype(np.float32), rc, stride, pad, dil, relu=False, impl=impl, f32_out=f32, quant=quant)
    got = y.astype(np.float32).transpose(0, 3, 1, 2)
    assert got.shape == expect.shape
    bad = got != expect.astype(np.float32)
    assert not bad.any(), f"{bad.sum()} of {bad.size} outputs differ; first at {np.argwhere(bad)[0]}: got {got[bad][0]} want {expect[bad][0]}"
    assert len(np.unique(expect)) > 32           # the case exercises the whole range, not a saturated constant


@pytest.fixture(scope="module")
def tiny_int8():
    return quantize.ensure_fixture("fcn_tiny_int8")


@pytest.fixture(params=["int8 plan", "fp16-carried"])
def plan_kind(request):
    """Quantised models run as an int8 plan (u8 tensors, native int8 MMA) by default; INFUR_B200_I8=0, read when a model is
    loaded, keeps them on fp16-carried tensors (also the fallback for models an int8 plan cannot express).  Both are exact."""
    import os
    old = os.environ.get("INFUR_B200_I8")
    os.environ["INFUR_B200_I8"] = "1" if request.param == "int8 plan" else "0"
    yield request.param
    if old is None:
        os.environ.pop("INFUR_B200_I8", None)
    else:
        os.environ["INFUR_B200_I8"] = old


def test_quantised_model_info(handle, tiny_int8):
    handle.model_load(tiny_int8)
    info = handle.model_info()
    assert info.input_names == ["input"] and info.output_names == ["out"] and info.input0_dtype == "Float"


@pytest.mark.parametrize("w,h", [(128, 96), (320, 240), (200, 136), (194, 130)])   # the last: odd feature-map widths (97, 49, 25)
def test_quantised_model_lowres_bit_exact(handle, tiny_int8, plan_kind, w, h):
    handle.model_load("")
    handle.model_load(tiny_int8)
    assert (" int8 " in handle.plan_text(1, w, h)) == (plan_kind == "int8 plan")
    g = onnx_min.load(tiny_int8)
    frame = synth.synth_frame(w, h, 3)
    env = qlinear.run(g, oracle.preprocess_f32(frame)[None])
    assert env["__max_abs_acc__"] < 2**24
    want = env[qlinear.lowres_name(g)][0]
    got = handle.model_lowres(frame)
    assert got.shape == want.shape
    assert (got == want).all(), f"{(got != want).sum()} of {got.size} low-resolution logits differ (max {np.abs(got - want).max()})"


def test_quantised_pipeline_vs_oracle(handle, tiny_int8, plan_kind):
    from test_gpu_pipeline import check_against_oracle
    g = onnx_min.load(tiny_int8)
    pipe = P.GpuPipeline(handle)
    pipe.control(("Model", ""))
    pipe.control(("Model", tiny_int8))
    pipe.control(("Scale", 0.5))
    frame = synth.synth_frame(640, 480, 5)
    out = pipe.advance(P.Frame(3, frame))
    scaled = oracle.scale_nearest(frame, 0.5)
    env = qlinear.run(g, oracle.preprocess_f32(scaled)[None])
    logits = env["out"][0]
    klass, rgba = oracle.color_code_image(logits)
    ref = {"class_map": klass.astype(np.uint8), "logits": logits, "decoded_rgba": rgba}
    assert out.id == 3 and out.size == [320, 240]
    assert (out.buffer == oracle.frame_rgba(scaled)).all()
    check_against_oracle(out.class_map, out.decoded_buffer, ref, 0.999)
    # in fact nothing differs: the logits before Resize are the oracle's bit for bit, and the Resize + ColorCode kernel is
    # bit-exact given equal logits (tests/test_gpu_stages.py) -- the whole path is integer / LUT / exactly-ordered f32 work
    assert (out.class_map == ref["class_map"]).all(), f"{(out.class_map != ref['class_map']).sum()} class-map pixels differ"
    assert (out.decoded_buffer == ref["decoded_rgba"]).all()
    pipe.control(("Scale", 1.0))


def test_quantised_batch_matches_single(handle, tiny_int8):
    handle.model_load(tiny_int8)
    frames = np.stack([synth.synth_frame(320, 240, i) for i in range(8)])
    res = handle.advance_batch(frames, ids=list(range(1, 9)))
    one = handle.advance(frames[5], id=6)
    assert (res[5]["class_map"] == one["class_map"]).all() and (res[5]["decoded_rgba"] == one["decoded_rgba"]).all()


def test_quantised_fcn50_config1(handle, plan_kind):
    """configs[0] analogue on the kind of model the reference's tests load (int8 FCN-ResNet50, 320x240): the low-resolution
    logits of all 53 quantised convolutions + 16 quantised adds are bit-exact, the class map follows."""
    from test_gpu_pipeline import check_against_oracle
    path = quantize.ensure_fixture("fcn50_int8")
    g = onnx_min.load(path)
    handle.model_load("")
    handle.model_load(path)
    handle.scale_control(1.0)
    frame = synth.synth_frame(320, 240, 7)
    env = qlinear.run(g, oracle.preprocess_f32(frame)[None])
    assert env["__max_abs_acc__"] < 2**24
    got = handle.model_lowres(frame)
    want = env[qlinear.lowres_name(g)][0]
    assert (got == want).all(), f"{(got != want).sum()} of {got.size} low-resolution logits differ"
    out = handle.advance(frame, id=1)
    logits = env["out"][0]
    klass, rgba = oracle.color_code_image(logits)
    ref = {"class_map": klass.astype(np.uint8), "logits": logits, "decoded_rgba": rgba}
    check_against_oracle(out["class_map"], out["decoded_rgba"], ref, 0.999)
    assert (out["class_map"] == ref["class_map"]).all() and (out["decoded_rgba"] == ref["decoded_rgba"]).all()   # end to end bit-identical
    assert len(np.unique(out["class_map"])) >= 8


def test_quantised_fcn50_1080p_full_size(handle):
    """BASELINE's frame size: one 1080p frame through the quantised FCN-ResNet50, low-resolution logits bit-exact against the
    integer oracle (about 20 s of CPU work), and frame 0 of a batch of 8 identical to the single-frame result."""
    path = quantize.ensure_fixture("fcn50_int8")
    g = onnx_min.load(path)
    handle.model_load(path)
    handle.scale_control(1.0)
    frames = np.stack([synth.synth_frame(1920, 1080, i) for i in range(2)])
    env = qlinear.run(g, oracle.preprocess_f32(frames[0])[None])
    assert env["__max_abs_acc__"] < 2**24
    got = handle.model_lowres(frames[0])
    want = env[qlinear.lowres_name(g)][0]
    assert got.shape == want.shape == (21, 135, 240)
    assert (got == want).all(), f"{(got != want).sum()} of {got.size} low-resolution logits differ"
    batch = np.ascontiguousarray(np.resize(frames, (8, 1080, 1920, 3)))
    res = handle.advance_batch(batch, ids=list(range(1, 9)), want=("class_map", "decoded_rgba"))
    one = handle.advance(frames[0], id=1, want=("class_map", "decoded_rgba"))
    assert (res[0]["class_map"] == one["class_map"]).all() and (res[2]["decoded_rgba"] == one["decoded_rgba"]).all()
    klass, rgba = oracle.color_code_image(env["out"][0])
    assert (one["class_map"] == klass).all() and (one["decoded_rgba"] == rgba).all()   # 1080p: class map and mask bit-identical end to end


def test_golden_qlinear_on_gpu(handle):
    """The committed golden vectors (tests/golden/qlinear.npz): QLinearConv (3x3, dilation 2) + QLinearAdd through the CUDA path."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "qlinear.npz"))
    xs, ys, rs, cs = np.float32(0.023), np.float32(0.031), np.float32(0.019), np.float32(0.027)
    quant = {"qmul": (xs * g["w_scale"]) / ys, "q_lo": -131, "q_hi": 255 - 131, "q_ra": ys / cs, "q_rb": rs / cs, "q_lo2": 0, "q_hi2": 255}
    xc = (g["q_in"].astype(np.int32) - 121).transpose(0, 2, 3, 1)
    wc = g["w"].astype(np.int32).transpose(0, 2, 3, 1)
    rc = g["q_res"].astype(np.int32).transpose(0, 2, 3, 1)
    for impl in (L.CONV_TCGEN05, L.CONV_TCGEN05_HALO):
        if impl == L.CONV_TCGEN05_HALO:   # the halo variant has no residual input: check the convolution alone
            y = handle.conv_test(xc, wc, g["bias"].astype(np.float32), None, 1, 2, 2, impl=impl, quant={k: quant[k] for k in ("qmul", "q_lo", "q_hi")})
            assert (y.astype(np.int32).transpose(0, 3, 1, 2) + 131 == g["q_conv"]).all()
        else:
            y = handle.conv_test(xc, wc, g["bias"].astype(np.float32), rc, 1, 2, 2, impl=impl, quant=quant)
            assert (y.astype(np.int32).transpose(0, 3, 1, 2) == g["q_add"]).all()


def _fused_add_differs(ra, rb, a_c, b_c):
    """Positions where a fused multiply-add (a * ra + round(b * rb), ONE rounding) would round differently from QLinearAdd's
    three separately rounded f32 operations.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 although both carry
    explicit rounding modifiers (csrc/ptx.cuh mul_f32x2_sep): this is what such a kernel would compute."""
    F = np.float32
    vb = (F(rb) * b_c.astype(F)).astype(F)
    fused = np.rint((a_c.astype(np.float64) * np.float64(F(ra)) + vb.astype(np.float64)).astype(F))
    unfused = np.rint(((F(ra) * a_c.astype(F)).astype(F) + vb).astype(F))
    return fused != unfused


def _scale_pairs_where_fusion_matters(count, a_c, b_c):
    rng = np.random.default_rng(2024)
    out = []
    while len(out) < count:
        ra, rb = np.float32(rng.uniform(0.2, 3.0)), np.float32(rng.uniform(0.2, 3.0))
        if _fused_add_differs(ra, rb, a_c, b_c).sum() >= 3:
            out.append((ra, rb))
    return out


# impl, relu, zero point of the sum, scale multiplier on (ra, rb)
ADD_SWEEP = [
    (L.CONV_TCGEN05, False, 0, 1.0),                  # fp16-carried tensors
    (L.CONV_TCGEN05_I8, False, 0, 1.0),               # integer tail + small accumulators (ConvTcGeom::q_tail 2)
    (L.CONV_TCGEN05_I8, True, 37, 1.0),               # ... with a ReLU above a non-zero zero point: the byte-wise floor
    (L.CONV_TCGEN05_I8_PAIR, False, 3, 1.0),
    (L.CONV_TCGEN05_I8_PAIR_DEEP, True, 0, 1.0),
    (L.CONV_TCGEN05_I8, False, 0, 8192.0),            # 255 * (ra + rb) >= 2^22: the generic tail (f32 clamp before the rounding add)
]


@pytest.mark.parametrize("impl,relu,c_zp,mult", ADD_SWEEP, ids=lambda v: str(v))
def test_qlinear_add_every_byte_pair(handle, impl, relu, c_zp, mult):
    """QLinearConv -> QLinearAdd over EVERY (conv output byte, residual byte) pair, with scale pairs picked so that an
    implementation that fuses the multiply into the add gets at least three of the 65 536 pairs wrong."""
    n, h, w, cin, cout = 1, 16, 32, 64, 256
    y_zp, r_zp = 128, 41
    i8 = impl != L.CONV_TCGEN05
    # x = 0 everywhere: accumulator + bias = bias[c] = c - 128 -> the conv output of channel c is the byte c at every pixel;
    # the residual byte at pixel p, channel c is (p + 7 c) mod 256: all 256 values meet every channel
    xq = np.zeros((n, cin, h, w), np.uint8)
    wq = np.ones((cout, cin, 1, 1), np.int8)
    bq = (np.arange(cout) - 128).astype(np.int32)
    pix = np.arange(h * w).reshape(1, 1, h, w)
    rq = ((pix + 7 * np.arange(cout).reshape(1, cout, 1, 1)) % 256).astype(np.uint8)
    one = np.float32(1.0)
    yq = qlinear.qlinear_conv(xq, one, np.uint8(0), wq, np.full(cout, one), np.zeros(cout, np.int8), one, np.uint8(y_zp), bq, 1, 0, 1, [])
    assert (yq[0, :, 0, 0] == np.arange(cout)).all()
    a_c = yq.astype(np.int32) - y_zp
    b_c = rq.astype(np.int32) - r_zp
    pairs = _scale_pairs_where_fusion_matters(3, a_c, b_c)
    for ra, rb in pairs:
        ra, rb = np.float32(ra * mult), np.float32(rb * mult)
        cq = qlinear.qlinear_add(yq, ra, np.uint8(y_zp), rq, rb, np.uint8(r_zp), one, np.uint8(c_zp))
        if relu:
            cq = np.maximum(cq, np.uint8(c_zp))
        quant = {"qmul": np.full(cout, one), "q_lo": -y_zp, "q_hi": 255 - y_zp, "q_ra": ra, "q_rb": rb, "q_lo2": -c_zp, "q_hi2": 255 - c_zp,
                 "q_zres": r_zp, "q_zout": c_zp}
        xc = xq.astype(np.int32).transpose(0, 2, 3, 1)
        wc = wq.astype(np.int32).transpose(0, 2, 3, 1)
        rc = b_c.transpose(0, 2, 3, 1)
        y = handle.conv_test(xc, wc, bq.astype(np.float32), rc, 1, 0, 1, relu=relu, impl=impl, quant=quant)
        got = y.astype(np.int32).transpose(0, 3, 1, 2) + c_zp
        bad = got != cq.astype(np.int32)
        assert not bad.any(), f"ra {ra} rb {rb}: {bad.sum()} outputs differ; first at {np.argwhere(bad)[0]}: got {got[bad][0]} want {cq[bad][0]}"
        if mult == 1.0:   # the case discriminates: a kernel with the fused form differs from the oracle on these inputs
            assert _fused_add_differs(ra, rb, a_c, b_c).sum() >= 3
    assert len(pairs) == 3


@pytest.mark.parametrize("cin,impl", [(64, L.CONV_TCGEN05_I8), (1024, L.CONV_TCGEN05_I8), (64, L.CONV_TCGEN05_I8_PAIR)],
                         ids=["small_acc_int_tail", "generic_tail", "pair_int_tail"])
def test_qlinear_conv_relu_over_nonzero_zero_point(handle, cin, impl):
    """QLinearConv + ReLU whose output tensor has a non-zero zero point (no residual): the stored byte is floored at the zero
    point -- in the integer tail that is the byte-wise max after the saturating pack (ConvTcGeom::q_floor)."""
    n, h, w, cout, y_zp = 2, 17, 23, 256, 23
    rng = np.random.default_rng(cin)
    xq = rng.integers(0, 256, size=(n, cin, h, w), dtype=np.uint8)
    wq = rng.integers(-127, 128, size=(cout, cin, 1, 1), dtype=np.int8)
    bq = rng.integers(-20000, 20000, size=cout, dtype=np.int32)
    x_scale, y_scale = np.float32(0.021), np.float32(0.043)
    w_scale = (rng.random(cout).astype(np.float32) + np.float32(0.5)) * np.float32(y_scale / x_scale / (25.0 * np.sqrt(cin)))
    yq = qlinear.qlinear_conv(xq, x_scale, np.uint8(0), wq, w_scale, np.zeros(cout, np.int8), y_scale, np.uint8(y_zp), bq, 1, 0, 1, [])
    expect = np.maximum(yq, np.uint8(y_zp)).astype(np.int32)
    assert (yq < y_zp).mean() > 0.1 and (yq == 255).any() and len(np.unique(yq)) > 200          # both ends of the range are exercised
    quant = {"qmul": (x_scale * w_scale) / y_scale, "q_lo": -y_zp, "q_hi": 255 - y_zp, "q_zout": y_zp}
    y = handle.conv_test(xq.astype(np.int32).transpose(0, 2, 3, 1), wq.astype(np.int32).transpose(0, 2, 3, 1), bq.astype(np.float32), None, 1, 0, 1,
                         relu=True, impl=impl, quant=quant)
    got = y.astype(np.int32).transpose(0, 3, 1, 2) + y_zp
    assert (got == expect).all()
