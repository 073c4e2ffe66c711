"""CPU: the oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §8c).  The oracle is the checker for the GPU tests, so it is pinned first."""
import numpy as np
import pytest

import oracle
from oracle import colorcode, scale


def test_color_2():
    """decode_predict.rs:93-97: color_code(2, 0.5) == from_rgba_unmultiplied(25, 225, 255, 127)."""
    c = oracle.color_code(2, 0.5)
    ref = oracle.color32_from_rgba_unmultiplied(25, 225, 255, 127)
    assert (c == ref).all()
    assert c[3] == 127


def test_decode_0to1():
    """decode_predict.rs:99-116: linspace(0,1) over [22,24,32] -> class 21 everywhere, alpha monotone, last 255, 32x24."""
    k, h, w = 22, 24, 32
    hm = np.linspace(0.0, 1.0, k * h * w, dtype=np.float32).reshape(k, h, w)
    klass, rgba = oracle.color_code_image(hm)
    assert rgba.shape == (24, 32, 4)
    assert (klass == 21).all()
    a = rgba[..., 3].ravel().astype(np.int32)
    assert (np.diff(a) >= 0).all()
    assert a[-1] == 255
    # the reference test rebuilds each pixel from its own alpha: color_code(21, a/255) round-trips
    back = oracle.color_code(np.full(h * w, 21), a.astype(np.float32) / np.float32(255.0)).reshape(h, w, 4)
    assert (back == rgba).all()


def test_alpha_roundtrip_all_bytes():
    a = np.arange(256, dtype=np.float32)
    assert (colorcode.alpha_u8(a / np.float32(255.0)) == np.arange(256)).all()


def test_argmax_semantics():
    """decode_predict.rs:67-77: strict '>', first max wins, all <= 0 or NaN -> class 0 / alpha 0."""
    hm = np.zeros((3, 1, 4), dtype=np.float32)
    hm[:, 0, 0] = [0.5, 0.5, 0.2]        # tie -> first
    hm[:, 0, 1] = [-1.0, -2.0, 0.0]      # nothing positive -> class 0, conf 0
    hm[:, 0, 2] = [np.nan, 0.3, np.nan]  # NaN never wins
    hm[:, 0, 3] = [0.1, 7.0, 7.0]
    k, c = colorcode.argmax_conf(hm)
    assert k[0].tolist() == [0, 0, 1, 1]
    assert c[0].tolist() == [0.5, 0.0, np.float32(0.3), 7.0]
    _, rgba = oracle.color_code_image(hm)
    assert rgba[0, 1].tolist() == [0, 0, 0, 0]
    assert rgba[0, 3, 3] == 255


def test_scale_from_size0():
    """processing.rs:288-295"""
    with pytest.raises(oracle.ScaleError) as e:
        oracle.scale_nearest(np.zeros((10, 0, 3), np.uint8), 0.99)
    assert e.value.kind == "ZeroSizeIn"


def test_scale_to_size0():
    """processing.rs:296-303"""
    with pytest.raises(oracle.ScaleError) as e:
        oracle.scale_nearest(np.zeros((10, 10, 3), np.uint8), 0.00000001)
    assert e.value.kind == "ZeroSizeOut"


def test_valid_scale():
    with pytest.raises(oracle.ScaleError):
        oracle.valid_scale(0.0)
    with pytest.raises(oracle.ScaleError):
        oracle.valid_scale(-1.0)
    assert np.isnan(oracle.valid_scale(float("nan")))  # NaN passes (processing.rs:161)


@pytest.mark.parametrize("w,h,f,ew,eh", [(1280, 720, 0.5, 640, 360), (640, 480, 0.5, 320, 240), (1280, 720, 1.0, 1280, 720),
                                         (1280, 720, 2.0, 2560, 1440), (640, 480, 0.1, 64, 48)])
def test_app_sizes(w, h, f, ew, eh):
    """Sizes asserted by the pipeline tests app.rs:181-216."""
    assert oracle.scaled_size(w, h, f) == (ew, eh)


def test_nearest_matches_pil_and_cv2_on_dyadic():
    """Independent implementations of centre-aligned nearest agree on the configs' factors."""
    cv2 = pytest.importorskip("cv2")
    from PIL import Image

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    for f in (0.5, 2.0, 0.25):
        o = oracle.scale_nearest(img, f)
        nh, nw = o.shape[:2]
        assert (np.asarray(Image.fromarray(img).resize((nw, nh), Image.NEAREST)) == o).all()
        assert (cv2.resize(img, (nw, nh), interpolation=cv2.INTER_NEAREST_EXACT) == o).all()
    assert (oracle.scale_nearest(img, 0.5) == img[1::2, 1::2]).all()  # src = 2x+1


def test_preprocess_matches_lut_and_torchvision_formula():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    x = oracle.preprocess_f32(img)
    lut = oracle.norm_lut()
    assert (x[0] == lut[0][img[..., 2]]).all() and (x[1] == lut[1][img[..., 1]]).all() and (x[2] == lut[2][img[..., 0]]).all()
    approx = (img[..., ::-1].astype(np.float64) / 255.0 - [0.485, 0.456, 0.406]) / [0.229, 0.224, 0.225]
    assert np.abs(x.transpose(1, 2, 0) - approx).max() < 1e-5


def test_upsample_matches_torch_interpolate():
    """The final Resize is F.interpolate(bilinear, align_corners=False) in torchvision's FCN."""
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(2)
    x = rng.standard_normal((5, 30, 40)).astype(np.float32)
    ref = F.interpolate(torch.from_numpy(x)[None], size=(240, 320), mode="bilinear", align_corners=False)[0].numpy()
    got = oracle.upsample_bilinear(x, 240, 320)
    assert np.abs(got - ref).max() < 2e-6
    x2 = rng.standard_normal((3, 17, 23)).astype(np.float32)
    ref2 = F.interpolate(torch.from_numpy(x2)[None], size=(131, 179), mode="bilinear", align_corners=False)[0].numpy()
    assert np.abs(oracle.upsample_bilinear(x2, 131, 179) - ref2).max() < 1e-5  # non-dyadic ratio: f32 rounding of the source coordinate


def test_infer_seg_model_shapes(tiny):
    """predict_onnx.rs:370-381: an all-zero 320x240 image gives outputs of [21,240,320]."""
    from oracle import fcn

    _, model = tiny
    r = fcn.pipeline(model, np.zeros((240, 320, 3), np.uint8), 1.0)
    assert r["logits"].shape == (21, 240, 320)
    assert r["class_map"].shape == (240, 320) and r["decoded_rgba"].shape == (240, 320, 4)


def test_fp16_emulation_tracks_fp32(tiny):
    from infur_b200 import synth
    from oracle import fcn

    _, model = tiny
    f = synth.synth_frame(160, 120, 3)
    a = fcn.pipeline(model, f, 1.0)
    b = fcn.pipeline(model, f, 1.0, emulate_fp16=True)
    assert (a["class_map"] == b["class_map"]).mean() > 0.97


def test_golden_vectors():
    """Committed fixtures (tests/golden/make_golden.py) pin the oracle across refactors."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stages.npz"))
    assert (oracle.scale_nearest(g["frame"], 0.5) == g["scaled_half"]).all()
    assert (oracle.scale_nearest(g["frame"], 0.37) == g["scaled_037"]).all()
    assert (oracle.preprocess_f32(g["scaled_half"]) == g["pre_half"]).all()
    assert (oracle.color_lut() == g["color_lut"]).all()
    up = oracle.upsample_bilinear(g["lowres"], 48, 64)
    assert (up == g["upsampled"]).all()
    k, rgba = oracle.color_code_image(up)
    assert (k == g["class_map"]).all() and (rgba == g["decoded"]).all()


def test_bilinear_scale_extension_tracks_cv2():
    """The opt-in bilinear Scale (not a reference mode) is the plain half-pixel bilinear: within 1 grey level of
    cv2.INTER_LINEAR (which rounds differently: fixed-point weights) for up- and down-scaling, same sizes/errors as nearest."""
    import cv2

    from infur_b200 import synth

    img = synth.synth_frame(127, 93, 1)
    for f in (0.5, 0.37, 2.0, 1.5):
        out = oracle.scale_bilinear(img, f)
        nw, nh = oracle.scaled_size(127, 93, f)
        assert out.shape == (nh, nw, 3)
        ref = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
        assert np.abs(out.astype(np.int32) - ref.astype(np.int32)).max() <= 1
    assert (oracle.scale_bilinear(img, 1.0) == img).all()
    with pytest.raises(oracle.ScaleError):
        oracle.scale_bilinear(np.zeros((10, 0, 3), np.uint8), 0.5)
