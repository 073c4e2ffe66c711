"""GPU parity of the integer / LUT stages through the C ABI: bit-exact against the oracle."""
import os

import numpy as np
import pytest

import oracle
from infur_b200 import processors as P
from infur_b200 import synth

pytestmark = pytest.mark.gpu


def test_color_lut_exact(handle):
    assert (handle.color_lut() == oracle.color_lut()).all()


def test_color_2(handle):
    """decode_predict.rs:93-97 through the GPU ColorCode."""
    hm = np.zeros((3, 1, 1), np.float32)
    hm[2] = 0.5
    cls, rgba = handle.color_code(hm)
    assert cls[0, 0] == 2
    assert (rgba[0, 0] == oracle.color32_from_rgba_unmultiplied(25, 225, 255, 127)).all()


def test_decode_0to1(handle):
    """decode_predict.rs:99-116 on the GPU: K = 22 exercises klass % 20."""
    k, h, w = 22, 24, 32
    hm = np.linspace(0.0, 1.0, k * h * w, dtype=np.float32).reshape(k, h, w)
    rgba = P.ColorCode(handle).advance(hm)
    assert rgba.shape == (24, 32, 4)
    cls, _ = handle.color_code(hm)
    assert (cls == 21).all()
    a = rgba[..., 3].ravel().astype(np.int32)
    assert (np.diff(a) >= 0).all() and a[-1] == 255
    _, ref = oracle.color_code_image(hm)
    assert (rgba == ref).all()


def test_color_code_edge_cases(handle):
    rng = np.random.default_rng(3)
    hm = (rng.standard_normal((21, 37, 53)) * 0.8).astype(np.float32)
    hm[:, 0, 0] = -1.0              # nothing positive
    hm[:, 0, 1] = np.nan            # all NaN
    hm[:, 0, 2] = 0.25              # ties: first wins
    hm[:, 0, 3] = np.inf
    hm[5, 0, 4] = 1e30
    cls, rgba = handle.color_code(hm)
    k, ref = oracle.color_code_image(hm)
    assert (cls == k).all() and (rgba == ref).all()


@pytest.mark.parametrize("w,h,f", [(640, 480, 0.5), (1280, 720, 0.5), (320, 240, 2.0), (127, 93, 0.37), (640, 480, 0.73), (64, 48, 0.99),
                                   (33, 21, 0.1), (200, 100, 1.0)])
def test_scale_parity(handle, w, h, f):
    img = synth.synth_frame(w, h, 1)
    s = P.Scale(handle)
    s.control(f)
    out = s.advance(P.Frame(7, img))
    assert out.id == 7
    ref = oracle.scale_nearest(img, f)
    assert out.img.shape == ref.shape
    assert (out.img == ref).all()
    assert not s.is_dirty()
    s.control(1.0)


def test_scale_errors_and_dirty(handle):
    """processing.rs:288-303 and the dirty semantics of :220-233."""
    s = P.Scale(handle)
    s.control(1.0)
    with pytest.raises(P.ValidScaleError) as e:
        s.control(0.0)
    assert "Cannot scale by negative number" in str(e.value)
    with pytest.raises(P.ValidScaleError):
        s.control(-2.0)
    s.control(0.99)
    assert s.is_dirty()
    with pytest.raises(P.ScaleProcError) as e:
        s.advance(P.Frame(0, np.zeros((10, 0, 3), np.uint8)))
    assert e.value.kind == "ZeroSizeIn"
    assert not s.is_dirty()          # advance clears dirty before failing (processing.rs:233)
    s.control(0.00000001)
    with pytest.raises(P.ScaleProcError) as e:
        s.advance(P.Frame(0, np.zeros((10, 10, 3), np.uint8)))
    assert e.value.kind == "ZeroSizeOut"
    s.control(0.5)
    assert s.is_dirty()
    s.control(0.5)
    assert not s.is_dirty()          # dirty = new != old
    s.control(0.25)
    assert s.is_dirty()
    assert s.advance(None, "keep") == "keep"
    assert not s.is_dirty()
    s.control(1.0)
    z = s.advance(P.Frame(3, np.zeros((0, 10, 3), np.uint8)))   # unit scale clones even an empty frame
    assert z.img.shape == (0, 10, 3)


def test_preprocess_exact(handle):
    img = synth.synth_frame(321, 123, 2)
    assert (handle.preprocess(img) == oracle.preprocess_f32(img)).all()
    ramp = np.arange(256, dtype=np.uint8)[None, :, None].repeat(3, 2).repeat(2, 0)
    assert (handle.preprocess(ramp) == oracle.preprocess_f32(ramp)).all()


@pytest.mark.parametrize("lh,lw,oh,ow,k", [(30, 40, 240, 320, 21), (6, 8, 48, 64, 21), (17, 23, 131, 179, 21), (135, 240, 1080, 1920, 21),
                                            (9, 5, 67, 33, 3), (12, 12, 90, 90, 40)])
def test_upsample_color_exact(handle, lh, lw, oh, ow, k):
    rng = np.random.default_rng(lh * 1000 + lw)
    low = (rng.standard_normal((k, lh, lw)) * 2.0).astype(np.float32)
    frame = synth.synth_frame(ow, oh, 0)
    r = handle.upsample_color(low, oh, ow, frame_bgr=frame, want_logits=(oh * ow < 200000))
    up = oracle.upsample_bilinear(low, oh, ow)
    kk, rgba = oracle.color_code_image(up)
    if r["logits"] is not None:
        assert (r["logits"] == up).all()
    assert (r["class_map"] == (kk & 0xFF)).all()
    assert (r["decoded_rgba"] == rgba).all()
    assert (r["blended_rgba"] == oracle.blend_over(rgba, oracle.frame_rgba(frame))).all()


def test_golden_stages(handle):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stages.npz"))
    handle.scale_control(0.5)
    assert (handle.scale_advance(g["frame"]) == g["scaled_half"]).all()
    handle.scale_control(0.37)
    assert (handle.scale_advance(g["frame"]) == g["scaled_037"]).all()
    handle.scale_control(1.0)
    assert (handle.preprocess(g["scaled_half"]) == g["pre_half"]).all()
    r = handle.upsample_color(g["lowres"], 48, 64, want_logits=True)
    assert (r["logits"] == g["upsampled"]).all() and (r["class_map"] == g["class_map"]).all() and (r["decoded_rgba"] == g["decoded"]).all()


@pytest.mark.parametrize("w,h,f", [(640, 480, 0.5), (127, 93, 0.37), (320, 240, 2.0), (200, 136, 0.73)])
def test_bilinear_scale_extension_exact(w, h, f):
    """INFUR_RESIZE_BILINEAR (opt-in; the reference only has Nearest): bit-exact against its oracle definition, alone and
    as the first stage of the fused path."""
    from infur_b200 import _lib as L

    frame = synth.synth_frame(w, h, 2)
    with P.Handle(max_batch=1, resize_mode=L.RESIZE_BILINEAR) as hd:
        hd.scale_control(f)
        ref = oracle.scale_bilinear(frame, f)
        assert (hd.scale_advance(frame) == ref).all()
        r = hd.advance(frame, 1, want=("scaled_bgr", "frame_rgba"))
        assert (r["scaled_bgr"] == ref).all() and (r["frame_rgba"] == oracle.frame_rgba(ref)).all()
