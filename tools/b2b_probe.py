"""Bring-up probe for conv_b2b_kernel: bit-compare against the unfused plan at a small size, then time both at 8 x 1080p."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infur_b200 import processors as P, synth

path = synth.fixture_path("fcn50")
if not os.path.exists(path):
    synth.ensure_fixture("fcn50")
frames = np.stack([synth.synth_frame(320, 240, i) for i in range(2)])
res = {}
for name, env in (("sep", {"INFUR_B200_NO_B2B": "1"}), ("fused", {"INFUR_B200_B2B": "force"})):
    for k in ("INFUR_B200_NO_B2B", "INFUR_B200_B2B"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with P.Handle(max_batch=2, autotune=False) as h:
        h.model_load(path)
        res[name] = (h.model_lowres(frames[0]), h.advance_batch(frames, want=("class_map",)))
        if name == "fused":
            print("\n".join(ln[:200] for ln in h.plan_text(2, 320, 240).splitlines() if "b2b" in ln))
d = np.abs(res["sep"][0] - res["fused"][0])
print("lowres max |diff|", d.max(), "class maps equal", all((a["class_map"] == b["class_map"]).all() for a, b in zip(res["sep"][1], res["fused"][1])))
