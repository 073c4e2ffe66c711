"""One convolution of the network's shape through infur_b200_conv_test on the GPU (for ncu captures and timing).

    python tools/conv_case.py NAME [--n 8]        NAME in CASES below (FCN-ResNet50 layers at 1080p: H/8 = 135, W/8 = 240)

Prints the kernel time of the last of 3 launches and the achieved TFLOP/s and GB/s (algorithmic).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name: (h, w, cin, cout, k, stride, pad, dil, relu, residual)
CASES = {
    "l1_conv3": (270, 480, 64, 256, 1, 1, 0, 1, True, True),
    "l1_down": (270, 480, 64, 256, 1, 1, 0, 1, False, False),
    "l2_conv3": (135, 240, 128, 512, 1, 1, 0, 1, True, True),
    "l1_conv1a": (270, 480, 64, 64, 1, 1, 0, 1, True, False),
    "l1_conv2": (270, 480, 64, 64, 3, 1, 1, 1, True, False),
    "l2_conv2": (135, 240, 128, 128, 3, 1, 1, 1, True, False),
    "l3_conv1": (135, 240, 1024, 256, 1, 1, 0, 1, True, False),
    "l3_conv2": (135, 240, 256, 256, 3, 1, 2, 2, True, False),
    "l3_conv3": (135, 240, 256, 1024, 1, 1, 0, 1, True, True),
    "l4_conv1": (135, 240, 2048, 512, 1, 1, 0, 1, True, False),
    "l4_conv2": (135, 240, 512, 512, 3, 1, 4, 4, True, False),
    "l4_conv3": (135, 240, 512, 2048, 1, 1, 0, 1, True, True),
    "l4_conv3_nores": (135, 240, 512, 2048, 1, 1, 0, 1, True, False),
    "l3_conv3_nores": (135, 240, 256, 1024, 1, 1, 0, 1, True, False),
    "l4_down": (135, 240, 1024, 2048, 1, 1, 0, 1, False, False),
    "cls0": (135, 240, 2048, 512, 3, 1, 1, 1, True, False),
    "stem": (1080, 1920, 3, 64, 7, 2, 3, 1, True, False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--impl", type=int, default=0, help="0 tcgen05, 2 CTA pair, 3 halo patch (3x3)")
    a = ap.parse_args()
    from infur_b200 import processors as P

    h, w, cin, cout, k, s, p, d, relu, res = CASES[a.name]
    n = a.n
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, h, w, cin), dtype=np.float32).astype(np.float16)
    wt = (rng.standard_normal((cout, k, k, cin), dtype=np.float32) * (2.0 / (cin * k * k)) ** 0.5).astype(np.float16)
    b = rng.standard_normal(cout).astype(np.float32)
    oh = (h + 2 * p - d * (k - 1) - 1) // s + 1
    ow = (w + 2 * p - d * (k - 1) - 1) // s + 1
    r = rng.standard_normal((n, oh, ow, cout), dtype=np.float32).astype(np.float16) if res else None
    with P.Handle(max_batch=n) as hd:
        y, ms = hd.conv_test(x, wt, b, r, s, p, d, relu, impl=a.impl, timed=True)
    fl = 2.0 * n * oh * ow * cout * k * k * cin
    by = x.nbytes + wt.nbytes + y.nbytes + (r.nbytes if res else 0)
    print(f"{a.name} impl{a.impl}: {ms:.3f} ms  {fl / ms * 1e-9:.1f} TFLOP/s  {by / ms * 1e-6:.1f} GB/s  (n={n})  checksum {float(np.abs(y[0, :4, :4].astype(np.float32)).sum()):.3f}")


if __name__ == "__main__":
    main()
