"""Small pipeline runs for compute-sanitizer (odd sizes, scale 0.5 / 2.0, blend, aux, the ring; float, quantised fp16-carried
and int8-plan models):

    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infur_b200 import processors as P, quantize, synth  # noqa: E402

path, _ = synth.ensure_fixture("fcn_tiny")
qpath = quantize.ensure_fixture("fcn_tiny_int8")
with P.Handle(max_batch=2, blend=True, compute_aux=True) as h:
    for label, p, env in (("float", path, None), ("quantised, int8 plan", qpath, "1"), ("quantised, fp16-carried", qpath, "0")):
        if env is not None:
            os.environ["INFUR_B200_I8"] = env
        h.model_load("")
        h.model_load(p)
        for (w, hh, f) in ((97, 65, 1.0), (160, 120, 0.5), (64, 48, 2.0)):
            h.scale_control(f)
            fr = np.stack([synth.synth_frame(w, hh, i) for i in range(2)])
            r = h.advance_batch(fr, ids=[1, 2], want=("frame_rgba", "class_map", "decoded_rgba", "blended_rgba", "logits_f32"))
            print(label, w, hh, f, r[0]["class_map"].shape, int(r[0]["class_map"].sum()))
    h.scale_control(1.0)
    t, view = h.ring_acquire(2, 97, 65); view[...] = 7; h.ring_submit(t); print(h.ring_wait(t)["n"])
print("ok")
