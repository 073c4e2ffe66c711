"""GPU: the tcgen05 implicit-GEMM convolution against a plain PyTorch fp32 reference of the same op
(floating point: fp16 inputs, f32 accumulation; tolerance 2e-3 + 4e-3*|ref| for fp16 outputs, i.e. ~2 fp16 ulp
plus accumulation-order noise; 1e-3 + 1e-3*|ref| for f32 outputs), and against the CUDA-core validation kernel."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from infur_b200 import _lib as L

pytestmark = pytest.mark.gpu


def ref_conv(x, w, b, res, stride, pad, dil, relu):
    xt = torch.from_numpy(x.astype(np.float32)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(w.astype(np.float32)).permute(0, 3, 1, 2)
    y = F.conv2d(xt, wt, torch.from_numpy(b), stride, pad, dil).permute(0, 2, 3, 1)
    if res is not None:
        y = y + torch.from_numpy(res.astype(np.float32))
    if relu:
        y = torch.relu(y)
    return y.numpy()


CASES = [
    # n, h, w, cin, cout, k, stride, pad, dil, relu, res, f32
    (1, 8, 16, 64, 64, 1, 1, 0, 1, True, False, False),
    (2, 30, 40, 64, 256, 1, 1, 0, 1, True, True, False),
    (1, 30, 40, 256, 128, 1, 1, 0, 1, True, False, False),
    (1, 30, 40, 512, 2048, 1, 1, 0, 1, True, True, False),
    (1, 30, 40, 64, 64, 3, 1, 1, 1, True, False, False),
    (1, 30, 40, 128, 128, 3, 1, 2, 2, True, False, False),
    (1, 17, 23, 64, 64, 3, 1, 4, 4, True, False, False),
    (1, 60, 80, 128, 128, 3, 2, 1, 1, True, False, False),
    (2, 31, 45, 64, 64, 3, 2, 1, 1, True, False, False),
    (1, 60, 80, 256, 512, 1, 2, 0, 1, False, False, False),
    (1, 30, 40, 512, 21, 1, 1, 0, 1, False, False, True),
    (1, 64, 96, 3, 64, 7, 2, 3, 1, True, False, False),
    (2, 61, 75, 3, 64, 7, 2, 3, 1, True, False, False),
    (1, 30, 40, 2048, 512, 3, 1, 1, 1, True, False, False),
    (3, 5, 7, 64, 128, 3, 1, 1, 1, False, False, False),
    (1, 135, 240, 256, 1024, 1, 1, 0, 1, True, True, False),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_%dx%d_c%d-%d_k%d_s%d_p%d_d%d%s%s%s" % (c[:9] + ("_relu" if c[9] else "", "_res" if c[10] else "", "_f32" if c[11] else "")))
def test_conv_tc_vs_torch(handle, case):
    n, h, w, cin, cout, k, stride, pad, dil, relu, res, f32 = case
    rng = np.random.default_rng(abs(hash(case)) % 2**31)
    x = rng.standard_normal((n, h, w, cin)).astype(np.float16)
    wt = (rng.standard_normal((cout, k, k, cin)) * (2.0 / (cin * k * k)) ** 0.5).astype(np.float16)
    b = rng.standard_normal(cout).astype(np.float32)
    oh = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    ow = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    r = rng.standard_normal((n, oh, ow, cout)).astype(np.float16) if res else None
    ref = ref_conv(x, wt, b, r, stride, pad, dil, relu)
    y = handle.conv_test(x, wt, b, r, stride, pad, dil, relu, impl=L.CONV_TCGEN05, f32_out=f32)
    tol = (1e-3 + 1e-3 * np.abs(ref)) if f32 else (2e-3 + 4e-3 * np.abs(ref))
    err = np.abs(y.astype(np.float32) - ref)
    assert (err <= tol).all(), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"
    if n * oh * ow * cout * cin * k * k < 3e9:
        yv = handle.conv_test(x, wt, b, r, stride, pad, dil, relu, impl=L.CONV_VALIDATE, f32_out=f32)
        assert (np.abs(yv.astype(np.float32) - ref) <= tol).all()


PAIR_CASES = [
    # the CTA-pair (cta_group::2) variant: cout % 256 == 0; odd numbers of M tiles exercise the dummy half of the last pair
    (1, 30, 40, 256, 256, 1, 1, 0, 1, True, False),
    (3, 17, 23, 128, 512, 1, 1, 0, 1, True, True),
    (2, 30, 40, 128, 256, 3, 1, 2, 2, True, False),
    (1, 60, 80, 256, 512, 1, 2, 0, 1, False, False),
    (1, 9, 11, 64, 256, 3, 1, 1, 1, True, True),
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=lambda c: "pair_n%d_%dx%d_c%d-%d_k%d_s%d_p%d_d%d" % c[:9])
def test_conv_pair_bit_identical_to_single_cta(handle, case):
    """Every N-tile / CTA-pair choice the autotuner may make accumulates each output element's K blocks in the same
    order, so the variants must agree BIT FOR BIT (that is what makes the plan-time choice invisible in results)."""
    n, h, w, cin, cout, k, stride, pad, dil, relu, res = case
    rng = np.random.default_rng(abs(hash(case)) % 2**31)
    x = rng.standard_normal((n, h, w, cin)).astype(np.float16)
    wt = (rng.standard_normal((cout, k, k, cin)) * (2.0 / (cin * k * k)) ** 0.5).astype(np.float16)
    b = rng.standard_normal(cout).astype(np.float32)
    oh = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    ow = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    r = rng.standard_normal((n, oh, ow, cout)).astype(np.float16) if res else None
    ref = ref_conv(x, wt, b, r, stride, pad, dil, relu)
    y1 = handle.conv_test(x, wt, b, r, stride, pad, dil, relu, impl=L.CONV_TCGEN05)
    y2 = handle.conv_test(x, wt, b, r, stride, pad, dil, relu, impl=L.CONV_TCGEN05_PAIR)
    assert (y1.view(np.uint16) == y2.view(np.uint16)).all()
    assert (np.abs(y2.astype(np.float32) - ref) <= 2e-3 + 4e-3 * np.abs(ref)).all()
    if res:   # the deep-epilogue pair variant (eight chunk buffers: four stores + four residual chunks in flight)
        y3 = handle.conv_test(x, wt, b, r, stride, pad, dil, relu, impl=L.CONV_TCGEN05_PAIR_DEEP)
        assert (y1.view(np.uint16) == y3.view(np.uint16)).all()


HALO_CASES = [
    # the halo-patch variant (3x3 / stride 1 / pad = dilation): image edges, all three dilations of the network, ragged sizes
    (1, 30, 40, 64, 64, 1),
    (2, 30, 40, 128, 128, 2),
    (1, 17, 23, 128, 256, 4),
    (3, 5, 7, 64, 128, 1),
    (1, 33, 9, 64, 64, 2),
]


@pytest.mark.parametrize("case", HALO_CASES, ids=lambda c: "halo_n%d_%dx%d_c%d-%d_d%d" % c)
def test_conv_halo_bit_identical_to_tapwise(handle, case):
    """The halo variant fetches each activation patch once and reads the nine taps through shifted UMMA descriptors;
    K blocks are accumulated in the same (chunk-major) order as in the tap-wise kernel, so the outputs agree bit for bit."""
    n, h, w, cin, cout, d = case
    rng = np.random.default_rng(abs(hash(case)) % 2**31)
    x = rng.standard_normal((n, h, w, cin)).astype(np.float16)
    wt = (rng.standard_normal((cout, 3, 3, cin)) * (2.0 / (cin * 9)) ** 0.5).astype(np.float16)
    b = rng.standard_normal(cout).astype(np.float32)
    ref = ref_conv(x, wt, b, None, 1, d, d, True)
    y1 = handle.conv_test(x, wt, b, None, 1, d, d, True, impl=L.CONV_TCGEN05)
    y3 = handle.conv_test(x, wt, b, None, 1, d, d, True, impl=L.CONV_TCGEN05_HALO)
    assert (y1.view(np.uint16) == y3.view(np.uint16)).all()
    assert (np.abs(y3.astype(np.float32) - ref) <= 2e-3 + 4e-3 * np.abs(ref)).all()
