import sys, numpy as np
sys.path.insert(0, "/root/repo")
from infur_b200 import processors as P, synth
path, _ = synth.ensure_fixture("fcn_tiny")
with P.Handle(max_batch=2, blend=True, compute_aux=True) as h:
    h.model_load(path)
    for (w, hh, f) in ((97, 65, 1.0), (160, 120, 0.5), (64, 48, 2.0)):
        h.scale_control(f)
        fr = np.stack([synth.synth_frame(w, hh, i) for i in range(2)])
        r = h.advance_batch(fr, ids=[1, 2], want=("frame_rgba", "class_map", "decoded_rgba", "blended_rgba", "logits_f32"))
        print(w, hh, f, r[0]["class_map"].shape, int(r[0]["class_map"].sum()))
    t, view = h.ring_acquire(2, 97, 65); view[...] = 7; h.ring_submit(t); print(h.ring_wait(t)["n"])
print("ok")
