"""bench.py's command-line contract on a machine without a GPU: the B200 arm refuses to run (no CPU fallback), the
reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900, cwd=ROOT)


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-budget-s", "5")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "1080p frames/sec through FCN-ResNet50" and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d


def test_roofline_block_accounting():
    """bench.roofline_block on a hand-made plan: tensor-core fraction from the summed conv time, the layer-wise floor (every launch at
    the better of its tensor-core and HBM time), and DRAM traffic per conv LAYER -- fused kernels (conv_b2b, stem_pool) counted, the
    layer a fused tail swallowed sharing its launch -- from the newest ncu launch list under profiles/."""
    sys.path.insert(0, ROOT)
    import bench

    plan = [
        "conv stem [..] k7 ++ maxpool | tcgen05 fused stem + max-pool (stem_pool_kernel) | GFLOP 100 MB 655.26",     # HBM-bound: 0.1 ms
        "maxpool pool [..] | (fused into previous) | MB 0",
        "conv a [..] k3 ++ b 1x1 | tcgen05 fused 3x3 -> 1x1 (conv_b2b_kernel, cmid 64) | GFLOP 1369.7 MB 65.526",    # tensor-bound: 1 ms
        "conv b [..] k1 +res | (fused into previous) | GFLOP 0 MB 0",
        "conv c [..] k1 | tcgen05 tile 16x8px x N256 pair | GFLOP 2739.4 MB 655.26",                                 # tensor-bound: 2 ms
    ]
    op_ms = [0.2, 0.003, 1.5, 0.003, 2.5, 0.05, 0.1]          # per op, then pre and post kernel
    pk = {"bf16_tflops_sustained": 1369.7, "bf16_tflops": 1618.3, "hbm_gbs": 6552.6, "source": "test"}
    flops_per_frame = (100 + 1369.7 + 2739.4) * 1e9
    out = bench.roofline_block("f16", plan, op_ms, 3, 1, flops_per_frame, pk, 4.5)
    roof = out[0] if isinstance(out, tuple) else out
    conv_ms = 0.2 + 1.5 + 0.003 + 2.5
    assert abs(roof["ms_all_conv_launches"] - conv_ms) < 1e-9
    assert abs(roof["achieved"] - flops_per_frame / (conv_ms * 1e-3) / 1e12) < 1e-6 and abs(roof["frac"] - roof["achieved"] / 1369.7) < 1e-9
    assert abs(roof["layerwise_floor_ms"] - 3.1) < 1e-3 and abs(roof["frac_of_layerwise_floor"] - 3.1 / conv_ms) < 1e-3
    assert "4 conv layers" in roof["kernel"] and "3 launches" in roof["kernel"]
    assert abs(roof["algorithmic_bytes_per_launch"] - (655.26 + 65.526 + 655.26) * 1e6 / 4) < 1.0
    if roof["traffic"] is not None:        # per conv layer of THIS plan, from the kept launch list (fused kernels included)
        total, launches = bench.conv_traffic_from_profiles("f16")[0]
        assert launches >= 40 and abs(roof["traffic"] - total / 4) < 1.0
