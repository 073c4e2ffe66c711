"""Small runs of the round-2 fused kernels for compute-sanitizer (racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_fused.py
stem_pool_kernel's shared-memory exchange relies on one named barrier per convolution row with a double-buffered row; conv_b2b_kernel hands
the intermediate from the epilogue warps to the tensor core through an mbarrier."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infur_b200 import processors as P, synth  # noqa: E402

os.environ["INFUR_B200_B2B"] = "force"
os.environ["INFUR_B200_STEM_POOL"] = "force"
which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
path, _ = synth.ensure_fixture("fcn50" if which == "fcn50" else "fcn_tiny")
with P.Handle(max_batch=1, autotune=False, use_cuda_graph=False) as h:
    h.model_load(path)
    fr = synth.synth_frame(136, 72, 1)
    r = h.advance(fr, 1, want=("class_map",))
    print(which, "fused kernels in plan:", h.plan_text(1, 136, 72).count("fused"), int(r["class_map"].sum()))
print("ok")
