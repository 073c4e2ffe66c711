"""Per-kernel table of one step of the hot path on the GPU (CUDA events around every launch).

    python tools/profile_step.py [--batch 8] [--width 1920] [--height 1080] [--iters 5] [--steps 0]

Prints one line per kernel of the plan: shape, tile choice, ms, achieved TFLOP/s and GB/s (algorithmic FLOPs and
bytes from infur_b200_plan_text).  With --steps K it also runs K back-to-back steps (for an ncu launch list).
"""
import argparse
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--kind", default="fcn50")
    ap.add_argument("--scale", type=float, default=1.0, help="Scale factor (configs[4]: 3840x2160 at 0.5 vs 1.0)")
    ap.add_argument("--cuda-profiler", action="store_true", help="bracket the back-to-back steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    a = ap.parse_args()
    import torch

    from infur_b200 import processors as P
    from infur_b200 import synth

    path = synth.fixture_path(a.kind)
    if not os.path.exists(path):
        if a.kind.endswith("_int8"):   # quantised (QOperator) stand-in for the zoo's int8 file
            from infur_b200 import quantize
            quantize.ensure_fixture(a.kind)
        else:
            synth.ensure_fixture(a.kind)
    B, W, H = a.batch, a.width, a.height
    frames = np.stack([synth.synth_frame(W, H, i) for i in range(min(B, 2))])
    frames = np.ascontiguousarray(np.resize(frames, (B, H, W, 3)))
    d = torch.from_numpy(frames).cuda()
    OW, OH = (W, H) if a.scale == 1.0 else (int(np.float32(W) * np.float32(a.scale)), int(np.float32(H) * np.float32(a.scale)))
    d_class = torch.empty((B, OH, OW), dtype=torch.uint8, device="cuda")
    d_rgba = torch.empty((B, OH, OW, 4), dtype=torch.uint8, device="cuda")
    with P.Handle(max_batch=B) as h:
        h.model_load(path)
        h.scale_control(a.scale)
        for _ in range(2):
            h.advance_device(d.data_ptr(), B, W, H, d_class.data_ptr(), d_rgba.data_ptr(), sync=True)
        lines = [ln for ln in h.plan_text(B, W, H).splitlines() if ln.startswith(("conv ", "maxpool "))]
        if a.iters > 0:
            ms = h.profile_ops(d.data_ptr(), B, W, H, iters=a.iters)
            tot_ms = tot_fl = 0.0
            for ln, t in zip(lines, ms):
                gf = float(re.search(r"GFLOP ([0-9.e+-]+)", ln).group(1)) if "GFLOP" in ln else 0.0
                mb = float(re.search(r"MB ([0-9.e+-]+)", ln).group(1))
                tot_ms += t; tot_fl += gf
                print(f"{t:8.3f} ms {gf / t if t else 0:8.1f} TF/s {mb / t if t else 0:8.1f} GB/s | {ln}")
            print(f"{ms[len(lines)]:8.3f} ms pre-kernel (Scale + normalise) | {ms[len(lines) + 1]:8.3f} ms post-kernel (upsample + ColorCode)")
            tot_ms += ms[len(lines)] + ms[len(lines) + 1]
            print(f"total {tot_ms:.3f} ms per step of {B} frames -> {B / tot_ms * 1e3:.1f} frames/s (whole path), {tot_fl / tot_ms:.1f} TFLOP/s")
        stream = torch.cuda.ExternalStream(h.compute_stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = max(a.steps, 3) if not a.cuda_profiler else max(a.steps, 1)
        if a.cuda_profiler:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        e0.record(stream)
        for _ in range(K):
            h.advance_device(d.data_ptr(), B, W, H, d_class.data_ptr(), d_rgba.data_ptr())
        e1.record(stream)
        stream.synchronize()
        if a.cuda_profiler:
            torch.cuda.profiler.stop()
        t = e0.elapsed_time(e1) / K
        print(f"back-to-back: {t:.3f} ms per step -> {B / t * 1e3:.1f} frames/s")


if __name__ == "__main__":
    main()
