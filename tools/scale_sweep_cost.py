import sys, time, numpy as np
sys.path.insert(0, '.')
from infur_b200 import processors as P, synth
path = synth.fixture_path("fcn50")
synth.ensure_fixture("fcn50")
one = synth.synth_frame(1920,1080,0)
with P.Handle(max_batch=8) as h:
    h.model_load(path)
    t0=time.perf_counter(); h.advance(one,1,want=("class_map",)); print("first 1.0", time.perf_counter()-t0, h.plan_build_stats(), file=sys.stderr)
    for f in (0.9, 0.89, 0.8, 0.7, 0.5, 0.9):
        h.scale_control(f)
        t0=time.perf_counter(); h.advance(one,1,want=("class_map",)); t1=time.perf_counter()-t0
        print("factor", f, "first frame ms", 1e3*t1, h.plan_build_stats(), file=sys.stderr)
