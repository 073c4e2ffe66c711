"""GPU: opt-in extensions that answer the reference README's TODO list (README.md:73-80) and are NOT reference behaviour:
softmax confidence (README.md:76).  Each has its own oracle definition (oracle/colorcode.py) and stays off by default."""
import numpy as np
import pytest

import oracle
from infur_b200 import _lib as L
from infur_b200 import processors as P
from infur_b200 import synth
from oracle import fcn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,lh,lw,oh,ow", [(21, 30, 40, 240, 320), (21, 17, 25, 136, 200), (22, 12, 16, 96, 128), (5, 9, 9, 70, 66)])
def test_softmax_confidence(lib, k, lh, lw, oh, ow):
    rng = np.random.default_rng(k * 100 + lh)
    low = (rng.standard_normal((k, lh, lw)) * 3.0).astype(np.float32)
    low[:, 0, 0] = -5.0                      # all classes tie below zero: the raw mode gives class 0 / alpha 0, softmax gives 1/K
    low[:, 1, 1] = np.linspace(-30, 40, k)   # one class dominates: p -> 1, alpha saturates
    with P.Handle(device=0, max_batch=1, confidence=L.CONF_SOFTMAX) as h:
        got = h.upsample_color(low, oh, ow)
    logits = oracle.upsample_bilinear(low, oh, ow)
    p = oracle.softmax_confidence(logits)
    k_ref, _ = oracle.color_code_image(p)
    a_ref = oracle.colorcode.alpha_u8(p.max(axis=0))
    # same class except where two probabilities tie to the last bit (expf implementations differ by an ulp or two)
    top2 = np.sort(p, axis=0)[-2:]
    tie = (top2[1] - top2[0]) < 1e-5
    assert ((got["class_map"] == k_ref) | tie).all()
    assert (got["class_map"] == k_ref).mean() > 0.999
    da = np.abs(got["decoded_rgba"][..., 3].astype(int) - a_ref.astype(int))
    assert da.max() <= 1, f"alpha differs by {da.max()}"
    lut = oracle.color_lut()
    assert (got["decoded_rgba"] == lut[got["class_map"] % 20, got["decoded_rgba"][..., 3]]).all()
    assert abs(int(got["decoded_rgba"][0, 0, 3]) - int(255.0 / k)) <= 1


def test_softmax_pipeline_keeps_classes_of_positive_logits(lib, tiny):
    """Through the whole path: wherever the raw-mode winner has a positive logit, the softmax mode picks the same class (softmax is
    monotone); it differs only where every logit is <= 0, which the reference's scan from (0, 0.0) maps to class 0."""
    path, model = tiny
    frame = synth.synth_frame(320, 240, 4)
    with P.Handle(device=0, max_batch=1) as h0, P.Handle(device=0, max_batch=1, confidence=L.CONF_SOFTMAX) as h1:
        for h in (h0, h1):
            h.model_load(path)
        raw = h0.advance(frame, 1)
        soft = h1.advance(frame, 1)
    positive = raw["decoded_rgba"][..., 3] > 0
    assert (raw["class_map"][positive] == soft["class_map"][positive]).all()
    assert (soft["decoded_rgba"][..., 3] >= 255 // 21).all()          # p(winner) >= 1/K
    assert (soft["frame_rgba"] == raw["frame_rgba"]).all()
