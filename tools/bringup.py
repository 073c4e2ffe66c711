"""GPU bring-up: every check in its own process (a trapped kernel kills only that check).
Usage on the GPU box:  python tools/bringup.py [name ...]   -> prints one line per check."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def _ref_conv(x, w, b, res, stride, pad, dil, relu):
    import torch
    import torch.nn.functional as F

    xt = torch.from_numpy(x.astype("float32")).permute(0, 3, 1, 2)
    wt = torch.from_numpy(w.astype("float32")).permute(0, 3, 1, 2)
    y = F.conv2d(xt, wt, torch.from_numpy(b), stride, pad, dil).permute(0, 2, 3, 1)
    if res is not None:
        y = y + torch.from_numpy(res.astype("float32"))
    if relu:
        y = torch.relu(y)
    return y.numpy()


def _conv_case(name, n, h, w, cin, cout, k, stride, pad, dil, relu=True, res=False, f32=False, impls=(0, 1)):
    import numpy as np

    from infur_b200 import processors as P

    rng = np.random.default_rng(hash(name) % 2**31)
    x = rng.standard_normal((n, h, w, cin)).astype(np.float16)
    wt = (rng.standard_normal((cout, k, k, cin)) * (2.0 / (cin * k * k)) ** 0.5).astype(np.float16)
    b = rng.standard_normal(cout).astype(np.float32)
    oh = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    ow = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    r = rng.standard_normal((n, oh, ow, cout)).astype(np.float16) if res else None
    ref = _ref_conv(x, wt, b, r, stride, pad, dil, relu)
    out = []
    with P.Handle(max_batch=8) as hd:
        for impl in impls:
            y, ms = hd.conv_test(x, wt, b, r, stride, pad, dil, relu, impl=impl, f32_out=f32, timed=True)
            err = np.abs(y.astype(np.float32) - ref)
            tol = 2e-3 + 4e-3 * np.abs(ref) if not f32 else 1e-3 + 1e-3 * np.abs(ref)
            bad = int((err > tol).sum())
            out.append(f"impl{impl}: maxerr {err.max():.4g} bad {bad}/{err.size} {ms:.3f}ms")
            if bad:
                idx = np.argwhere(err > tol)[:5]
                out.append(f"  first bad idx {idx.tolist()} got {[float(y[tuple(i)]) for i in idx]} ref {[float(ref[tuple(i)]) for i in idx]}")
    print(name, "|", " | ".join(out))


@check
def conv_1x1_small():
    _conv_case("1x1 64->64 16x8", 1, 8, 16, 64, 64, 1, 1, 0, 1)


@check
def conv_1x1_k256():
    _conv_case("1x1 256->128 30x40", 1, 30, 40, 256, 128, 1, 1, 0, 1)


@check
def conv_1x1_n256_res():
    _conv_case("1x1 64->256 +res 30x40 n2", 2, 30, 40, 64, 256, 1, 1, 0, 1, res=True)


@check
def conv_1x1_2048():
    _conv_case("1x1 512->2048 +res 30x40", 1, 30, 40, 512, 2048, 1, 1, 0, 1, res=True)


@check
def conv_3x3():
    _conv_case("3x3 64->64 p1 30x40", 1, 30, 40, 64, 64, 3, 1, 1, 1)


@check
def conv_3x3_d2():
    _conv_case("3x3 128->128 d2 p2 30x40", 1, 30, 40, 128, 128, 3, 1, 2, 2)


@check
def conv_3x3_d4():
    _conv_case("3x3 64->64 d4 p4 17x23", 1, 17, 23, 64, 64, 3, 1, 4, 4)


@check
def conv_3x3_s2():
    _conv_case("3x3 128->128 s2 p1 60x80", 1, 60, 80, 128, 128, 3, 2, 1, 1)


@check
def conv_3x3_s2_odd():
    _conv_case("3x3 64->64 s2 p1 31x45", 2, 31, 45, 64, 64, 3, 2, 1, 1)


@check
def conv_1x1_s2():
    _conv_case("1x1 256->512 s2 60x80", 1, 60, 80, 256, 512, 1, 2, 0, 1, relu=False)


@check
def pair_1x1():
    _conv_case("pair 1x1 256->256 30x40", 1, 30, 40, 256, 256, 1, 1, 0, 1, impls=(2, 0))


@check
def pair_1x1_res_n3():
    _conv_case("pair 1x1 128->512 +res 17x23 n3 (odd M tiles)", 3, 17, 23, 128, 512, 1, 1, 0, 1, res=True, impls=(2, 0))


@check
def pair_3x3_d2():
    _conv_case("pair 3x3 128->256 d2 p2 30x40 n2", 2, 30, 40, 128, 256, 3, 1, 2, 2, impls=(2, 0))


@check
def pair_big():
    _conv_case("pair 1x1 512->2048 +res 135x240", 1, 135, 240, 512, 2048, 1, 1, 0, 1, res=True, impls=(2, 0))


@check
def halo_3x3():
    _conv_case("halo 3x3 64->64 p1 30x40", 1, 30, 40, 64, 64, 3, 1, 1, 1, impls=(3, 0))


@check
def halo_3x3_d2():
    _conv_case("halo 3x3 128->128 d2 p2 30x40 n2", 2, 30, 40, 128, 128, 3, 1, 2, 2, impls=(3, 0))


@check
def halo_3x3_d4():
    _conv_case("halo 3x3 128->256 d4 p4 17x23", 1, 17, 23, 128, 256, 3, 1, 4, 4, impls=(3, 0))


@check
def conv_head_f32():
    _conv_case("1x1 512->21 f32 30x40", 1, 30, 40, 512, 21, 1, 1, 0, 1, relu=False, f32=True)


@check
def conv_stem():
    _conv_case("7x7 3->64 s2 p3 64x96", 1, 64, 96, 3, 64, 7, 2, 3, 1)


@check
def conv_stem_odd():
    _conv_case("7x7 3->64 s2 p3 61x75 n2", 2, 61, 75, 3, 64, 7, 2, 3, 1)


@check
def conv_big_k():
    _conv_case("3x3 2048->512 p1 30x40 (classifier.0)", 1, 30, 40, 2048, 512, 3, 1, 1, 1)


@check
def smoke_tiny():
    import __graft_entry__ as g

    g.smoke()


@check
def pipeline_fcn50():
    import numpy as np

    from infur_b200 import processors as P
    from infur_b200 import synth
    from oracle import fcn

    path, model = synth.ensure_fixture("fcn50")
    frame = synth.synth_frame(320, 240, 0)
    ref = fcn.pipeline(model, frame, 1.0, emulate_fp16=True)
    for impl in (1, 0):
        with P.Handle(max_batch=1, conv_impl=impl) as h:
            h.model_load(path)
            r = h.advance(frame, 1, want=("class_map", "decoded_rgba", "logits_f32"))
            m = (r["class_map"] == ref["class_map"]).mean()
            le = np.abs(r["logits_f32"] - ref["logits"]).max()
            print(f"fcn50 320x240 impl{impl}: class match {m:.5f} logits maxerr {le:.4g} (ref max {np.abs(ref['logits']).max():.3g})")


def main():
    names = sys.argv[1:] or list(CHECKS)
    if len(names) == 1 and names[0].startswith("--run="):
        CHECKS[names[0][6:]]()
        return
    for n in names:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), f"--run={n}"], capture_output=True, text=True, timeout=300)
            tail = (p.stdout.strip() + ("\n" + p.stderr.strip()[-1500:] if p.returncode else "")).strip()
            print(f"[{n}] rc={p.returncode}\n{tail}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{n}] TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
