#!/usr/bin/env python
"""Headline benchmark: 1080p frames/s through Scale -> FCN-ResNet50 -> ColorCode (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1080p|4k] [--single-host]

One *step* = one batch of ``--batch`` (default 8) synthetic BGR frames per GPU through the whole hot path.
``--config 1080p`` (default) is ``configs[2]``: 1920x1080, scale 1.0, batch 8, pinned ring; the single-frame ``configs[1]``
latency is reported beside it as ``single_frame``.  ``--config 4k`` is ``configs[4]``: 3840x2160 frames at Scale 0.5 (the
network then sees 1080p) and at Scale 1.0, both in one JSON line (``sweep``), with overlay parity against the CPU oracle.

Multi-GPU, two ways:
  * torchrun (what the driver launches for N > 1): one rank per GPU, one single-device handle each, frames sharded by rank
    (weak scaling, no collective on the frame path; the packed weights are NCCL-broadcast from rank 0 at load);
  * ``--single-host --gpus N``: ONE process, ONE handle over N GPUs (``cfg.num_devices = N``: worker thread per GPU and the NCCL
    weight broadcast live inside the library), ONE producer thread feeding every GPU's pinned ring round-robin -- ``configs[3]`` as
    written ("stream sharded round-robin across 8xB200"); this measures the host-feed ceiling.

Printed by rank 0 as ONE JSON line:
  value         frames/s, inputs already resident in HBM (``infur_b200_advance_device``), CUDA events on the library's compute
                stream, max over ranks; the timed region is repeated until it lasts >= ``--min-seconds`` (default 2 s)
  e2e           frames/s through the pinned ring (``ring_acquire`` / ``ring_submit`` / ``ring_wait``): host memcpy of every frame
                into the pinned slot, H2D, the path, D2H of class map + decoded RGBA + frame RGBA, all inside the timed region
  parity        computed IN THIS RUN on the bench's own frames: class-map exact-match rate against the fp32 CPU oracle and
                against the fp16-emulating oracle, largest oracle top-2 margin over mismatching pixels, RGBA max |diff|
  roofline      the tcgen05 implicit-GEMM conv kernels: algorithmic conv FLOPs / summed CUDA-event time of the conv launches,
                timed INSIDE the sustained loop (every 8th step carries an event after each kernel), against the measured
                sustained bf16 peak; `traffic` = DRAM bytes per conv layer from the newest ncu launch list under profiles/;
                `layerwise_floor_ms` = sum over the launches of max(tensor-core time, HBM time) at the measured peaks
  int8          the same measurements for the QOperator-quantised network (the kind of file configs[0] names), same run
  cpu_baseline  the oracle pipeline (PyTorch-CPU fp32 + numpy) on the box's host cores, bounded sample (N = 1 only)
``--impl reference`` times that CPU pipeline alone (the reference's onnxruntime path cannot be built here: no Rust, no
onnxruntime, no model file -- see DESIGN.md).
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# 2 x MACs of the 55 convs of the `out` head, per frame, by the size the NETWORK sees
FLOPS_NO_AUX = {(1920, 1080): 2189.025e9, (3840, 2160): 8756.099e9}
PROFILE_EVERY = 8   # every 8th step of the timed loop records an event after each kernel


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["source"] = "MEASURED_PEAKS.json"
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_frame_fp32(model, bgr: np.ndarray, factor: float = 1.0):
    """One frame through the oracle restatement of the reference path (Scale -> pre-process -> FCN-ResNet50 fp32, both heads and
    both full-resolution Resize ops as ONNX Runtime executes them -> ColorCode on `out` -> display buffer), PyTorch-CPU + numpy.
    Returns (class_map, decoded_rgba, frame_rgba, full-resolution logits)."""
    import torch
    import torch.nn.functional as F

    import oracle

    with torch.no_grad():
        scaled = oracle.scale_nearest(bgr, factor)
        x = torch.from_numpy(oracle.preprocess_f32(scaled)[None])
        feats = model.backbone(x)
        out = F.interpolate(model.classifier(feats["out"]), size=x.shape[-2:], mode="bilinear", align_corners=False)
        if model.aux_classifier is not None:
            F.interpolate(model.aux_classifier(feats["aux"]), size=x.shape[-2:], mode="bilinear", align_corners=False)
        logits = out[0].numpy()
        klass, rgba = oracle.color_code_image(logits)
        return klass.astype(np.uint8), rgba, oracle.frame_rgba(scaled), logits


def cpu_frame_int8(graph, bgr: np.ndarray):
    """One frame through the quantised restatement (oracle/qlinear.py): Scale 1.0 -> pre-process -> integer-exact QOperator
    interpreter (torch-CPU f64 convolutions of integers) -> Resize -> ColorCode."""
    import oracle
    from oracle import qlinear

    scaled = oracle.scale_nearest(bgr, 1.0)
    env = qlinear.run(graph, oracle.preprocess_f32(scaled)[None])
    logits = env[graph.outputs[0][0]][0]
    klass, rgba = oracle.color_code_image(logits)
    return klass.astype(np.uint8), rgba, oracle.frame_rgba(scaled), logits


def parity_report(got_class, got_rgba, ref_class, ref_rgba, ref_logits) -> dict:
    """Class-map exact-match rate, largest oracle top-2 margin over mismatching pixels, RGBA differences."""
    same = got_class == ref_class
    out = {"class_map_match": float(same.mean()), "pixels": int(same.size), "mismatches": int((~same).sum())}
    if ref_logits is not None and (~same).any():
        lg = np.partition(ref_logits, -2, axis=0)
        margin = lg[-1] - np.maximum(lg[-2], 0.0)   # ColorCode's scan starts from (0, 0.0): 0 competes too
        out["max_oracle_margin_on_mismatch"] = float(margin[~same].max())
    else:
        out["max_oracle_margin_on_mismatch"] = 0.0
    d = np.abs(got_rgba.astype(np.int32) - ref_rgba.astype(np.int32))
    out["rgba_max_abs_diff"] = int(d.max())
    out["rgba_max_abs_diff_same_class"] = int(d[same].max()) if same.any() else 0
    out["alpha_max_abs_diff_same_class"] = int(d[..., 3][same].max()) if same.any() else 0
    return out


def model_fixture(kind: str):
    """(path, label, dtype) of the benchmarked network: FCN-ResNet50 fp16 (BASELINE configs[1..4]) or its QOperator-quantised form
    (the kind of file configs[0] names)."""
    from infur_b200 import quantize, synth

    if kind == "int8":
        return quantize.ensure_fixture("fcn50_int8"), "FCN-ResNet50 int8 (QOperator: QLinearConv / QLinearAdd; seeded synthetic weights, statically quantised)", "int8"
    path = synth.fixture_path("fcn50")
    if not os.path.exists(path):
        synth.ensure_fixture("fcn50")
    return path, "FCN-ResNet50 (seeded synthetic weights in an opset-12 .onnx)", "f16"


def frame_size(args):
    return (3840, 2160) if args.config == "4k" else (1920, 1080)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from infur_b200 import synth

    W, H = frame_size(args)
    factor = 0.5 if args.config == "4k" else 1.0
    cores = os.cpu_count() or 1
    frames = np.stack([synth.synth_frame(W, H, i) for i in range(2)])
    import torch

    torch.set_num_threads(cores)
    int8 = args.model == "int8"
    if int8:
        from oracle import onnx_min
        graph = onnx_min.load(model_fixture("int8")[0])
    else:
        _, model = synth.ensure_fixture("fcn50")
        model.eval()

    def one(i):
        if int8:
            return cpu_frame_int8(graph, frames[i % len(frames)])
        return cpu_frame_fp32(model, frames[i % len(frames)], factor)

    for i in range(min(args.warmup, 2)):
        one(i)
    # bounded sample: one frame per step, at most --cpu-budget-s seconds of CPU work in total
    t0 = time.perf_counter()
    done = 0
    while done < args.steps:
        one(done)
        done += 1
        if time.perf_counter() - t0 > args.cpu_budget_s and done >= 3:
            break
    dt = time.perf_counter() - t0
    fps = done / dt
    what = "integer-exact QOperator interpreter (torch-CPU f64 convolutions)" if int8 else "torch-CPU fp32 FCN-ResNet50 both heads"
    sample = f"{done} single {W}x{H} frames (one frame per step; {args.steps} requested, bounded to {args.cpu_budget_s:.0f} s), {what} + numpy Scale/ColorCode"
    print(json.dumps({
        "impl": "reference", "metric": "1080p frames/sec through FCN-ResNet50", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "steps_timed": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8" if int8 else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args) + ", " + model_fixture(args.model)[1] + "; CPU restatement of the reference's "
                               "onnxruntime path (the reference itself cannot be built here: no Rust/onnxruntime/model file)", "frames_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def workload_name(args) -> str:
    if args.config == "4k":
        return "configs[4]: 3840x2160 synthetic stream, Scale 0.5 (headline value; the network sees 1920x1080) and Scale 1.0 (sweep)"
    return "configs[2]: 1080p synthetic stream, batch=8 frames per step per GPU, scale 1.0"


# ------------------------------------------------------------------------------------------------ ncu evidence under profiles/
def conv_traffic_from_profiles(kind: str):
    """((DRAM bytes of all conv-layer kernels, their launch count), file) from the newest ncu launch list of one step kept under profiles/ (``tools/profile_step.py`` under
    ``ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum``); None when there is none."""
    tag = "int8" if kind == "int8" else "f16"
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_launches_step_{tag}_b8_1080p.csv")), key=lambda p: (os.path.basename(p).split("_")[0], os.path.getmtime(p)))
    if not cands and kind != "int8":
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_launches_step_b8_1080p.csv")))
    if not cands:
        return None, None
    path = cands[-1]
    rows = list(csv.reader(open(path)))
    try:
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    except IndexError:
        return None, None
    hdr = rows[hi]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_launch = {}
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if not d["Metric Name"].startswith("dram__bytes"):
            continue
        name = d["Kernel Name"]
        if not any(t in name for t in ("conv_tc", "stem_tc", "conv_halo", "conv_b2b", "stem_pool")):   # every kernel a conv layer runs in
            continue
        per_launch[d["ID"]] = per_launch.get(d["ID"], 0.0) + float(d["Metric Value"].replace(",", "")) * scale.get(d["Metric Unit"], 0.0)
    if not per_launch:
        return None, None
    return (sum(per_launch.values()), len(per_launch)), os.path.relpath(path, ROOT)


# ------------------------------------------------------------------------------------------------ GPU arm
class Measure:
    """One handle + one model: device-resident throughput with in-loop per-kernel timing, end-to-end ring throughput."""

    def __init__(self, args, h, dev, rank, world, W, H, dist):
        self.args, self.h, self.dev, self.rank, self.world, self.W, self.H, self.dist = args, h, dev, rank, world, W, H, dist
        import torch

        self.torch = torch
        self.stream = torch.cuda.ExternalStream(h.compute_stream(), device=dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v: float) -> float:
        if self.world > 1:
            t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def device_resident(self, d_sets, B, ow, oh, sampler=None, profile=True):
        """K steps per round, rounds repeated until the timed region lasts >= --min-seconds on every rank."""
        torch, h, args = self.torch, self.h, self.args
        d_class = torch.empty((B, oh, ow), dtype=torch.uint8, device=self.dev)
        d_rgba = torch.empty((B, oh, ow, 4), dtype=torch.uint8, device=self.dev)
        caps = (d_class.numel(), d_rgba.numel(), 0)
        nsets = len(d_sets)

        def step(i, prof):
            if prof:
                h.profile_step(d_sets[i % nsets].data_ptr(), B, self.W, self.H)
            else:
                h.advance_device(d_sets[i % nsets].data_ptr(), B, self.W, self.H, d_class.data_ptr(), d_rgba.data_ptr(), caps=caps)

        for i in range(args.warmup):
            step(i, False)
        self.stream.synchronize()
        torch.cuda.synchronize()
        self.barrier()
        if sampler:
            sampler.start()
        l0 = h.launch_count()
        profiled = os.environ.get("INFUR_BENCH_PROFILE") == "1"   # ncu --profile-from-start off: only the timed region is captured
        if profiled:
            torch.cuda.profiler.start()
        total_ms, rounds, steps_done = 0.0, 0, 0
        while True:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for i in range(args.steps):
                step(steps_done + i, profile and (steps_done + i) % PROFILE_EVERY == PROFILE_EVERY - 1)
            e1.record(self.stream)
            self.stream.synchronize()
            total_ms += e0.elapsed_time(e1)
            rounds += 1
            steps_done += args.steps
            # every rank runs the same number of rounds: continue while the SLOWEST rank's total is below the minimum ... decided
            # on the max so that no rank stops early
            if self.max_over_ranks(total_ms) >= 1e3 * args.min_seconds or rounds >= 1000 or profiled:
                break
        torch.cuda.synchronize()
        if profiled:
            torch.cuda.profiler.stop()
        launches = h.launch_count() - l0
        op_ms, prof_steps = h.profile_collect() if profile else ([], 0)
        self.barrier()
        ms = self.max_over_ranks(total_ms)
        return {"ms_total": ms, "rounds": rounds, "steps_timed": steps_done, "ms_per_step": ms / steps_done,
                "fps": self.world * B * steps_done / (ms * 1e-3), "launches": int(launches), "op_ms": op_ms, "prof_steps": prof_steps}

    def ring(self, host_sets, B, min_seconds):
        """End to end through the pinned ring; repeated until >= min_seconds."""
        h, args = self.h, self.args
        depth = args.ring_depth
        nsets = len(host_sets)
        sink = np.zeros(3, dtype=np.int64)

        def consume(r):
            sink[0] += int(r["class_map"][0, 0, 0]); sink[1] += int(r["decoded_rgba"][B - 1, -1, -1, 3]); sink[2] += int(r["frame_rgba"][B - 1, -1, -1, 0])

        def run(nsteps):
            inflight = []
            for i in range(nsteps):
                if len(inflight) == depth:
                    consume(h.ring_wait(inflight.pop(0)))
                t, view = h.ring_acquire(B, self.W, self.H)
                np.copyto(view, host_sets[i % nsets])      # the frame source writes into pinned memory
                h.ring_submit(t)
                inflight.append(t)
            for t in inflight:
                consume(h.ring_wait(t))

        run(max(args.warmup, depth))
        self.torch.cuda.synchronize()
        self.barrier()
        total_s, steps_done = 0.0, 0
        while True:
            t0 = time.perf_counter()
            run(args.steps)
            self.torch.cuda.synchronize()
            total_s += time.perf_counter() - t0
            steps_done += args.steps
            if self.max_over_ranks(total_s) >= min_seconds or steps_done >= 1000 * args.steps:
                break
        self.barrier()
        s = self.max_over_ranks(total_s)
        return {"fps": self.world * B * steps_done / s, "seconds": s, "steps_timed": steps_done}


def roofline_block(kind, plan_lines, op_ms, prof_steps, B, flops_per_frame, pk, ms_per_step):
    op_lines = [ln for ln in plan_lines if ln.startswith("conv ") or ln.startswith("maxpool ")]
    conv_ms = sum(m for ln, m in zip(op_lines, op_ms) if ln.startswith("conv "))
    n_conv = sum(1 for ln in op_lines if ln.startswith("conv "))
    pool_ms = sum(m for ln, m in zip(op_lines, op_ms) if ln.startswith("maxpool "))
    pre_ms, post_ms = op_ms[len(op_lines)], op_ms[len(op_lines) + 1]
    flops = B * flops_per_frame
    achieved = flops / (conv_ms * 1e-3) / 1e12
    # int8 plan: MEASURED_PEAKS.json has no int8 figure; kind::i8 issues at exactly twice the kind::f16 MAC rate on this part
    # (tools/ubench_mma.cu, profiles/r1_ubench_mma.txt), so the denominator is twice the measured sustained bf16 peak
    mult = 2.0 if kind == "int8" else 1.0
    peak = pk["bf16_tflops_sustained"] * mult
    # both per conv LAYER of the plan (fused layers share a launch: the same denominator keeps the ratio meaningful)
    traffic_tot, traffic_src = conv_traffic_from_profiles(kind)
    traffic = traffic_tot[0] / n_conv if traffic_tot else None
    algo_bytes = 1e6 * sum(float(ln.split(" MB ")[1]) for ln in op_lines if ln.startswith("conv ")) / n_conv
    n_launched = sum(1 for ln in op_lines if ln.startswith("conv ") and "(fused into previous)" not in ln)
    # layer-wise roofline: every launch at the better of its tensor-core time and its HBM time (algorithmic FLOPs and bytes of the plan,
    # measured peaks) -- what this layer-by-layer algorithm could reach at best; only cross-layer fusion moves it
    floor_ms = 0.0
    for ln in op_lines:
        if ln.startswith("conv ") and " MB " in ln:
            gf = float(ln.split("GFLOP ")[1].split()[0]) if "GFLOP " in ln else 0.0
            mb = float(ln.split(" MB ")[1].split()[0])
            floor_ms += max(gf / (peak * 1e3) * 1e3, mb * 1e6 / (pk["hbm_gbs"] * 1e9) * 1e3)
    roof = {
        "kernel": "conv_tc_* / stem_tc (tcgen05 implicit-GEMM conv+bias+ReLU(+residual)), all %d conv layers of one step (%d launches: fused bottleneck tails share one)" % (n_conv, n_launched),
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s" if kind == "int8" else "TFLOP/s", "frac": achieved / peak,
        "basis": "conv launches timed with CUDA events INSIDE the sustained loop (every %dth step, %d steps averaged) / sustained peak" % (PROFILE_EVERY, prof_steps),
        "frac_vs_burst_peak": achieved / (pk["bf16_tflops"] * mult),
        "frac_whole_step": flops / (ms_per_step * 1e-3) / 1e12 / peak,
        "layerwise_floor_ms": floor_ms, "frac_of_layerwise_floor": floor_ms / conv_ms if conv_ms > 0 else None,
        "traffic": traffic, "traffic_source": traffic_src, "traffic_launches_in_profile": traffic_tot[1] if traffic_tot else None,
        "algorithmic_bytes_per_launch": algo_bytes,
        "traffic_over_algorithmic": (traffic / algo_bytes) if traffic else None,
        "flops_per_launch_avg": flops / n_conv, "ms_per_launch_avg": conv_ms / n_conv, "ms_all_conv_launches": conv_ms,
        "ms_profiled_step": sum(op_ms), "peak_source": pk["source"] + (" x2 (kind::i8 issues at twice the kind::f16 rate)" if kind == "int8" else ""),
    }

    def hbm(name, mb_per_frame, t_ms):
        gbs = B * mb_per_frame * 1e6 / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        return {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                "ms_per_launch": t_ms, "algorithmic_MB_per_frame": mb_per_frame}
    esz = 1 if kind == "int8" else 2
    other = [hbm("pre kernel (Scale 1.0 + u8->fp16 normalise, NHWC4)", 6.2208 + 16.5888, pre_ms),
             hbm("post kernel (bilinear x8 upsample + argmax + colour)", 2.7216 + 8.2944 + 2.0736, post_ms)]
    if pool_ms > 0.02:      # a fused stem + pool leaves only the event overhead of the skipped op
        other.insert(1, hbm("maxpool3s2 kernel (3x3/s2, NHWC)", (66.3552 + 16.5888) * esz / 2, pool_ms))
    return roof, other


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    from infur_b200 import processors as P
    from infur_b200 import sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H = frame_size(args)
    B = args.batch
    pk = peaks()
    if rank == 0:
        model_fixture("f16")
        if args.model in ("both", "int8"):
            model_fixture("int8")
    if world > 1:
        dist.barrier()

    nsets = 4 if args.config == "1080p" else 2
    base = np.stack([synth.synth_frame(W, H, rank * 64 + i) for i in range(B)])
    host_sets = [np.ascontiguousarray(np.roll(base, s, axis=0)) for s in range(nsets)]
    d_sets = [torch.from_numpy(x).to(dev) for x in host_sets]
    # CPU legs, parity, single-frame latency: N = 1 only for the default config (for N > 1 the other ranks would idle in a barrier and
    # make the driver's GPU-busy sample meaningless); configs[4] asks for overlay parity in the 8-GPU run, so --config 4k keeps them
    solo = world == 1 or args.config == "4k"

    def measure_model(kind, factors):
        path, label, dtype = model_fixture(kind)
        h = P.Handle(device=local_rank, max_batch=B, ring_depth=args.ring_depth)
        t0 = time.perf_counter()
        sharding.load_model_sharded(h, path, rank, world, dev)
        load_s = time.perf_counter() - t0
        m = Measure(args, h, dev, rank, world, W, H, dist)
        res = {"label": label, "dtype": dtype, "load_s": load_s, "sweep": {}}
        for factor in factors:
            h.scale_control(factor)
            ow, oh = int(np.float32(W) * np.float32(factor)), int(np.float32(H) * np.float32(factor))
            t0 = time.perf_counter()
            h.advance_device_query(B, W, H)           # builds (and autotunes) the plan
            build_ms, tuned = h.plan_build_stats()
            sampler = ClockSampler(local_rank) if rank == 0 else None
            r = m.device_resident(d_sets, B, ow, oh, sampler)
            r["clocks"] = sampler.stop() if sampler else None
            r["plan_build_ms"], r["plan_tuned_convs"] = build_ms, tuned
            r["e2e"] = m.ring(host_sets, B, args.min_seconds)
            r["plan"] = h.plan_text(B, W, H).splitlines()
            r["out_size"] = (ow, oh)
            res["sweep"][factor] = r
        res["handle"], res["measure"] = h, m
        return res

    factors = [0.5, 1.0] if args.config == "4k" else [1.0]
    kinds = ["f16", "int8"] if args.model == "both" else [args.model]
    if args.config == "4k" and args.model == "both":
        kinds = ["f16"]
    out_models = {}
    for kind in kinds:
        res = measure_model(kind, factors)
        h = res["handle"]
        rep = {}
        for factor, r in res["sweep"].items():
            ow, oh = r["out_size"]
            roof, other = (None, None)
            if rank == 0 and r["op_ms"]:
                roof, other = roofline_block(kind, r["plan"], r["op_ms"], r["prof_steps"], B, FLOPS_NO_AUX[(ow, oh)], pk, r["ms_per_step"])
            rep[factor] = {
                "scale": factor, "network_input": [ow, oh], "value": r["fps"], "ms_per_step": r["ms_per_step"], "steps_timed": r["steps_timed"],
                "timed_seconds": r["ms_total"] * 1e-3, "rounds": r["rounds"],
                "e2e": {"value": r["e2e"]["fps"], "unit": "frames/s", "h2d_bytes_per_step": world * B * W * H * 3,
                        "d2h_bytes_per_step": world * B * ow * oh * 9, "timed_seconds": r["e2e"]["seconds"],
                        "api": "infur_b200_ring_acquire/submit/wait; host memcpy into the pinned slot, H2D, path, D2H of class map + decoded RGBA + frame RGBA inside the timed region"},
                "gpu_launches": r["launches"] * world, "clocks": r["clocks"], "roofline": roof, "roofline_other_kernels": other,
                "plan_build_ms": r["plan_build_ms"], "plan_tuned_convs": r["plan_tuned_convs"],
            }
        extra = {}
        if rank == 0 and solo:
            # ---- parity, in this run, on the bench's own frames (N = 1 only)
            f0 = factors[0]
            h.scale_control(f0)
            got = h.advance_batch(base[:2], want=("class_map", "decoded_rgba", "frame_rgba"))
            par = {}
            cpu = None
            if not args.no_cpu_baseline:
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                if kind == "int8":
                    from oracle import onnx_min
                    graph = onnx_min.load(model_fixture("int8")[0])
                    t0 = time.perf_counter()
                    rc, rr, rf, rl = cpu_frame_int8(graph, base[0])
                    sec = time.perf_counter() - t0
                    par["vs_integer_oracle"] = parity_report(got[0]["class_map"], got[0]["decoded_rgba"], rc, rr, rl)
                    par["frame_rgba_equal"] = bool((got[0]["frame_rgba"] == rf).all())
                    par["frames_compared"] = 1
                    cpu = {"value": 1.0 / sec, "unit": "frames/s", "cores": cores, "kind": "port",
                           "sample": f"1 single {W}x{H} frame ({sec:.1f} s); oracle port: integer-exact QOperator interpreter (torch-CPU f64 "
                                     "convolutions) + numpy Scale/ColorCode"}
                else:
                    from oracle import fcn
                    _, model = synth.ensure_fixture("fcn50")
                    model.eval()
                    cpu_frame_fp32(model, base[1], f0)   # warm-up
                    times, refs = [], []
                    for i in range(args.cpu_frames):
                        t0 = time.perf_counter()
                        refs.append(cpu_frame_fp32(model, base[i % 2], f0))
                        times.append(time.perf_counter() - t0)
                    sec = float(np.median(times))
                    rc, rr, rf, rl = refs[0]
                    par["vs_fp32_oracle"] = parity_report(got[0]["class_map"], got[0]["decoded_rgba"], rc, rr, rl)
                    par["frame_rgba_equal"] = bool((got[0]["frame_rgba"] == rf).all())
                    if len(refs) > 1:
                        rc1, rr1, rf1, rl1 = refs[1]
                        par["vs_fp32_oracle_frame1"] = parity_report(got[1]["class_map"], got[1]["decoded_rgba"], rc1, rr1, rl1)
                    if args.config == "4k" and len(factors) > 1:   # the other arm of the sweep: one 4K frame at Scale 1.0
                        h.scale_control(factors[1])
                        got1 = h.advance_batch(base[:1], want=("class_map", "decoded_rgba", "frame_rgba"))
                        t0 = time.perf_counter()
                        qc, qr, qf, ql = cpu_frame_fp32(model, base[0], factors[1])
                        par["scale_%g_vs_fp32_oracle" % factors[1]] = parity_report(got1[0]["class_map"], got1[0]["decoded_rgba"], qc, qr, ql)
                        par["scale_%g_frame_rgba_equal" % factors[1]] = bool((got1[0]["frame_rgba"] == qf).all())
                        par["scale_%g_cpu_s_per_frame" % factors[1]] = time.perf_counter() - t0
                        del got1, qc, qr, qf, ql
                        h.scale_control(f0)
                    emu = fcn.pipeline(model, base[0], f0, emulate_fp16=True)
                    par["vs_fp16_emulating_oracle"] = parity_report(got[0]["class_map"], got[0]["decoded_rgba"], emu["class_map"], emu["decoded_rgba"], emu["logits"])
                    par["frames_compared"] = min(2, len(refs))
                    cpu = {"value": 1.0 / sec, "unit": "frames/s", "cores": cores, "kind": "port",
                           "sample": f"{args.cpu_frames} single {W}x{H} frames at Scale {f0} after 1 warm-up (median {sec:.2f} s/frame); oracle port: "
                                     "torch-CPU fp32 FCN-ResNet50 both heads + numpy Scale/ColorCode"}
                par["note"] = ("oracle = CPU restatement of the reference path (parity unpinned: the reference itself cannot run here); the fp16 "
                               "network is compared by exact-match rate + oracle margin on mismatches, the int8 network bit for bit")
            extra["parity"] = par or None
            extra["cpu_baseline"] = cpu
            # ---- configs[1]: one frame, synchronous advance, host buffers
            if args.config == "1080p":
                h.scale_control(1.0)
                one = np.ascontiguousarray(base[0])
                h.advance(one, 1)

                def latency(frame, into):
                    lat = []
                    for _ in range(9):
                        t0 = time.perf_counter()
                        h.advance(frame, 1, want=("class_map", "decoded_rgba"), into=into)
                        lat.append(time.perf_counter() - t0)
                    return 1e3 * float(np.median(lat))

                # (a) the caller's buffers are ordinary (pageable) memory, re-used across frames like the reference's `out` arguments
                into_pg = {"class_map": np.empty((H, W), np.uint8), "decoded_rgba": np.empty((H, W, 4), np.uint8)}
                ms_pageable = latency(one, into_pg)
                # (b) the caller's buffers come from infur_b200_host_alloc (page-locked): copies are direct DMA
                pin_in, pin_c, pin_d = P.PinnedArray((H, W, 3)), P.PinnedArray((H, W)), P.PinnedArray((H, W, 4))
                pin_in.array[...] = one
                ms_pinned = latency(pin_in.array, {"class_map": pin_c.array, "decoded_rgba": pin_d.array})
                same = bool((pin_c.array == into_pg["class_map"]).all() and (pin_d.array == into_pg["decoded_rgba"]).all())
                # device-resident single frame for comparison (no copies)
                d_c1 = torch.empty((H, W), dtype=torch.uint8, device=dev); d_d1 = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
                st = torch.cuda.ExternalStream(h.compute_stream(), device=dev)
                for _ in range(3):
                    h.advance_device(d_sets[0].data_ptr(), 1, W, H, d_c1.data_ptr(), d_d1.data_ptr(), sync=True)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(10):
                    h.advance_device(d_sets[0].data_ptr(), 1, W, H, d_c1.data_ptr(), d_d1.data_ptr())
                e1.record(st)
                st.synchronize()
                dev_ms = e0.elapsed_time(e1) / 10
                extra["single_frame"] = {"workload": "configs[1]: one 1080p frame, synchronous infur_b200_advance, host buffers in and out (class map + decoded RGBA), "
                                                     "buffers re-used across calls, CUDA graph replay",
                                         "ms": ms_pinned, "ms_pageable_buffers": ms_pageable, "ms_device_resident_back_to_back": dev_ms,
                                         "note": "ms: caller buffers from infur_b200_host_alloc (pinned); identical results: %s" % same,
                                         "fraction_of_conv_roofline": (FLOPS_NO_AUX[(W, H)] / (pk["bf16_tflops_sustained"] * 1e12 * (2.0 if kind == "int8" else 1.0))) / (ms_pinned * 1e-3)}
                pin_in.close(); pin_c.close(); pin_d.close()
                # a new Scale factor: cost before the first result (the reference's slider, gui.rs:278-285)
                h.scale_control(0.9)
                t0 = time.perf_counter()
                h.advance(one, 1, want=("class_map",))
                ms_new = 1e3 * (time.perf_counter() - t0)
                b_ms, tuned = h.plan_build_stats()
                h.scale_control(0.9)
                t0 = time.perf_counter()
                h.advance(one, 1, want=("class_map",))
                extra["new_scale_factor"] = {"first_frame_ms": ms_new, "plan_build_ms": b_ms, "convs_autotuned": tuned,
                                             "next_frame_ms": 1e3 * (time.perf_counter() - t0), "factor": 0.9}
        h.close()
        out_models[kind] = (res, rep, extra)

    out = None
    if rank == 0:
        kind0 = kinds[0]
        res, rep, extra = out_models[kind0]
        head = rep[factors[0]]
        dtype = "f16" if kind0 == "f16" else "int8 (u8 activations x s8 weights -> s32 on tcgen05.mma.kind::i8; the RGB stem on fp16-carried integers)"
        out = {
            "metric": "1080p frames/sec through FCN-ResNet50", "value": head["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "steps_timed": head["steps_timed"], "timed_seconds": head["timed_seconds"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload_name(args) + ", " + res["label"] + ", out head only, class map + premultiplied RGBA + frame RGBA out",
                       "frames_per_step_per_gpu": B, "width": W, "height": H, "ring_depth": args.ring_depth, "sharding": "frames by rank, no collective",
                       "launch": "torchrun, one rank (one single-device handle) per GPU" if world > 1 else "one process, one GPU",
                       "l2": "inputs cycle through %d x %d distinct frames (%.0f MB) and each step streams > 30 GB of activations: larger than L2" % (nsets, B, nsets * B * W * H * 3 / 1e6)},
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "roofline": head["roofline"],
            "roofline_other_kernels": head["roofline_other_kernels"], "plan_build_ms": head["plan_build_ms"],
            "cpu_baseline": extra.get("cpu_baseline"), "parity": extra.get("parity"),
        }
        for k in ("single_frame", "new_scale_factor"):
            if k in extra:
                out[k] = extra[k]
        if len(factors) > 1:
            out["sweep"] = {str(f): {k: v for k, v in rep[f].items() if k != "roofline_other_kernels"} for f in factors}
        for kind in kinds[1:]:
            r2, rep2, extra2 = out_models[kind]
            hd = rep2[factors[0]]
            out[kind] = {"model": r2["label"], "value": hd["value"], "ms_per_step": hd["ms_per_step"], "timed_seconds": hd["timed_seconds"], "e2e": hd["e2e"],
                         "gpu_launches": hd["gpu_launches"], "clocks": hd["clocks"], "roofline": hd["roofline"], "parity": extra2.get("parity"),
                         "cpu_baseline": extra2.get("cpu_baseline"), "single_frame": extra2.get("single_frame"),
                         "dtype": "int8 (u8 x s8 -> s32, tcgen05.mma.kind::i8)"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ one process, N GPUs
def run_single_host(args):
    """configs[3] as written: ONE host stream sharded round-robin over N GPUs behind ONE handle; one producer thread."""
    import torch

    from infur_b200 import processors as P
    from infur_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    N = args.gpus
    if torch.cuda.device_count() < N:
        raise SystemExit(f"bench.py --single-host: {N} GPUs requested, {torch.cuda.device_count()} visible")
    W, H = frame_size(args)
    factor = 0.5 if args.config == "4k" else 1.0
    B = args.batch
    path, label, dtype = model_fixture("int8" if args.model == "int8" else "f16")
    h = P.Handle(devices=list(range(N)), max_batch=B, ring_depth=args.ring_depth) if N > 1 else P.Handle(device=0, max_batch=B, ring_depth=args.ring_depth)
    t0 = time.perf_counter()
    h.model_load(path)              # packs on devices[0], ncclBroadcast inside the library
    load_s = time.perf_counter() - t0
    sums = [h.weights_checksum(i) for i in range(N)]
    h.scale_control(factor)
    nsets = 4
    base = np.stack([synth.synth_frame(W, H, i) for i in range(B)])
    host_sets = [np.ascontiguousarray(np.roll(base, s, axis=0)) for s in range(nsets)]
    depth_total = args.ring_depth * N
    sink = np.zeros(2, dtype=np.int64)
    # the frame source: --producers threads fill a slot's frames in parallel (a decoder pool); 1 = the reference's single "Proc" thread
    pool = None
    if args.producers > 1:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=args.producers)

    def fill(view, src):
        if pool is None:
            np.copyto(view, src)
            return
        parts = np.array_split(np.arange(B), args.producers)
        futs = [pool.submit(np.copyto, view[p[0]:p[-1] + 1], src[p[0]:p[-1] + 1]) for p in parts if len(p)]   # numpy releases the GIL while copying
        for f in futs:
            f.result()

    def consume(r):
        sink[0] += int(r["class_map"][0, 0, 0]); sink[1] += int(r["decoded_rgba"][B - 1, -1, -1, 3])

    def run(nsteps):
        inflight = []
        for i in range(nsteps):
            if len(inflight) == depth_total:
                consume(h.ring_wait(inflight.pop(0)))      # submission order
            t, view = h.ring_acquire(B, W, H)              # ticket t -> GPU (t - 1) % N
            fill(view, host_sets[i % nsets])
            h.ring_submit(t)
            inflight.append(t)
        for t in inflight:
            consume(h.ring_wait(t))

    run(max(args.warmup, 1) * N + depth_total)             # every GPU builds its plan and warms up
    l0 = h.launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    total_s, steps_done = 0.0, 0
    while True:
        t0 = time.perf_counter()
        run(args.steps * N)
        total_s += time.perf_counter() - t0
        steps_done += args.steps * N
        if total_s >= args.min_seconds:
            break
    clocks = sampler.stop()
    launches = h.launch_count() - l0
    fps = B * steps_done / total_s
    # identity with a one-GPU handle on the same frames (frames are independent: which GPU ran them must not matter)
    t, view = h.ring_acquire(B, W, H)
    np.copyto(view, host_sets[1])
    h.ring_submit(t)
    r = h.ring_wait(t)
    grp_class, grp_rgba, grp_dev = r["class_map"].copy(), r["decoded_rgba"].copy(), r["device"]
    t2, view = h.ring_acquire(B, W, H)
    np.copyto(view, host_sets[1])
    h.ring_submit(t2)
    r2 = h.ring_wait(t2)
    same = bool((grp_class == r2["class_map"]).all() and (grp_rgba == r2["decoded_rgba"]).all())
    other_dev = r2["device"]
    h.close()
    ow, oh = int(np.float32(W) * np.float32(factor)), int(np.float32(H) * np.float32(factor))
    print(json.dumps({
        "metric": "1080p frames/sec through FCN-ResNet50", "value": fps, "unit": "frames/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / steps_done, "steps_timed": steps_done, "timed_seconds": total_s, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic", "mode": "single-host",
        "config": {"workload": ("configs[3]: ONE synthetic stream sharded round-robin over the GPUs of one box behind ONE handle (cfg.num_devices = %d), "
                                "weights ncclBroadcast inside infur_b200_model_load, %d producer thread(s) copying frames into the pinned slots; " % (N, args.producers)) + workload_name(args) + ", " + label,
                   "frames_per_step_per_gpu": B, "width": W, "height": H, "ring_depth": args.ring_depth, "sharding": "ring ticket t -> devices[(t - 1) % N]"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": B * W * H * 3, "d2h_bytes_per_step": B * ow * oh * 9,
                "api": "infur_b200_ring_acquire/submit/wait on a multi-device handle; host memcpy into the pinned slot inside the timed region",
                "host_feed_GBps": fps * (W * H * 3 + ow * oh * 9) / 1e9, "producer_threads": args.producers},
        "gpu_launches": int(launches), "clocks": clocks, "model_load_s": load_s, "weight_checksums_equal": len(set(sums)) == 1,
        "same_result_on_two_gpus": {"equal": same, "devices": [grp_dev, other_dev]},
        "note": "value == e2e here: the only path through a multi-device handle is the host-buffer ring (no device-resident leg)",
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="both", choices=["both", "f16", "int8"],
                    help="both (default): the fp16 network of BASELINE configs[1..4] as the headline and the QOperator-quantised network as `int8` in the same line")
    ap.add_argument("--config", default="1080p", choices=["1080p", "4k"])
    ap.add_argument("--single-host", action="store_true", help="one process, one handle over --gpus GPUs, one producer thread (configs[3])")
    ap.add_argument("--producers", type=int, default=1, help="--single-host: threads that copy frames into the pinned slots (1 = one Proc thread)")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--ring-depth", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=3)
    ap.add_argument("--min-seconds", type=float, default=2.0, help="repeat the K-step timed loop until the timed region lasts at least this long")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=90.0, help="--impl reference: stop after this many seconds of timed CPU work")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.model == "both":
            args.model = "f16"
        run_reference(args, rank, world)
    elif args.single_host:
        run_single_host(args)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
