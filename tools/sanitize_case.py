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
            # the network path proper: cell-per-thread post kernel + frame / blend pass (no full-resolution logits)
            r = h.advance_batch(fr, ids=[1, 2], want=("frame_rgba", "class_map", "decoded_rgba", "blended_rgba"))
            print(label, "cells", int(r[1]["class_map"].sum()))
    h.scale_control(1.0)
    t, view = h.ring_acquire(2, 97, 65); view[...] = 7; h.ring_submit(t); print(h.ring_wait(t)["n"])
    # w % 128 == 0: the warp-staged pre kernel
    fr = np.stack([synth.synth_frame(256, 40, i) for i in range(2)])
    print("warp128", int(h.advance_batch(fr, want=("class_map",))[0]["class_map"].sum()))
# round 2: the fused bottleneck tail (needs FCN-ResNet50: the tiny network has no candidate block), CUDA graphs on, and a
# two-context handle (the same GPU twice when only one is visible) through the ring and the frame-level API
os.environ["INFUR_B200_B2B"] = "force"
path50, _ = synth.ensure_fixture("fcn50")
with P.Handle(max_batch=2, autotune=False) as h:
    h.model_load(path50)
    fr = np.stack([synth.synth_frame(72, 40, i) for i in range(2)])
    r = h.advance_batch(fr, want=("class_map", "decoded_rgba"))
    print("b2b", h.plan_text(2, 72, 40).count("conv_b2b_kernel"), int(r[0]["class_map"].sum()))
import torch  # noqa: E402

if torch.cuda.device_count() < 2:
    os.environ["INFUR_B200_ALLOW_DUP_DEVICES"] = "1"
devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
with P.Handle(devices=devs, max_batch=2, ring_depth=2, autotune=False) as h:
    h.model_load(path)
    tickets = [h.submit(synth.synth_frame(64, 48, i), id=i + 1) for i in range(5)]
    print("group", [int(h.wait(t)["class_map"].sum()) for t in tickets])
print("ok")
