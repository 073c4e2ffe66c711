"""GPU: several GPUs behind ONE handle (cfg.num_devices > 1; SURVEY.md 8(b)/(e), configs[3]) and the ring / frame-level API contracts.

One "Proc" thread owns the handle (infur/src/main.rs:36-40,105-112); inside the library one worker thread per GPU issues that
GPU's copies and kernels, `model_load` packs the weights on devices[0] and ncclBroadcasts the arena, ring ticket t runs on
devices[(t - 1) % n], frame id on devices[(id - 1) % n], results come back in submission order.  The bar is byte identity with the
single-device handle on the same frames (frames are independent units of work: no temporal state in Scale / Model / ColorCode).

With one visible GPU the same code paths run with that GPU listed twice (INFUR_B200_ALLOW_DUP_DEVICES=1: the NCCL broadcast
becomes a device-to-device copy, everything else -- workers, routing, ordering -- is identical); with >= 2 GPUs the real thing runs.
"""
import os

import numpy as np
import pytest

import oracle
from infur_b200 import _lib as L
from infur_b200 import processors as P
from infur_b200 import synth

pytestmark = pytest.mark.gpu


def _devices(n=2):
    import torch

    have = torch.cuda.device_count()
    if have >= n:
        return list(range(n))
    os.environ["INFUR_B200_ALLOW_DUP_DEVICES"] = "1"
    return [i % max(have, 1) for i in range(n)]


@pytest.fixture(scope="module")
def group(lib):
    h = P.Handle(devices=_devices(2), max_batch=4, ring_depth=3, blend=True)
    yield h
    h.close()


@pytest.fixture(scope="module")
def single(lib):
    h = P.Handle(device=0, max_batch=4, ring_depth=3, blend=True)
    yield h
    h.close()


def test_create_rejects_bad_device_lists(lib):
    with pytest.raises(P.InfurError) as e:
        P.Handle(devices=[0, 99])
    assert e.value.code == L.E_NO_DEVICE
    old = os.environ.pop("INFUR_B200_ALLOW_DUP_DEVICES", None)
    try:
        with pytest.raises(P.InfurError) as e:
            P.Handle(devices=[0, 0])
        assert e.value.code == L.E_INVALID_ARG
    finally:
        if old is not None:
            os.environ["INFUR_B200_ALLOW_DUP_DEVICES"] = old


def test_weights_reach_every_device(group, single, tiny):
    path, _ = tiny
    group.model_load(path)
    single.model_load(path)
    assert group.num_devices() == 2 and single.num_devices() == 1
    sums = [group.weights_checksum(i) for i in range(2)]
    assert sums[0] == sums[1] == single.weights_checksum(0)
    info = group.model_info()
    assert info.output_names == ["out", "aux"]


def test_failed_load_keeps_previous_model_on_every_device(group, tiny, tmp_path):
    path, _ = tiny
    group.model_load(path)
    before = [group.weights_checksum(i) for i in range(2)]
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\x08\x07garbage")
    with pytest.raises(P.ModelCmdError):
        group.model_load(str(bad))
    assert [group.weights_checksum(i) for i in range(2)] == before
    group.model_load("")
    assert group.model_info() is None
    group.model_load(path)


def test_ring_round_robin_identical_to_single_device(group, single, tiny):
    """configs[3] in miniature: one host stream, slots alternate between the GPUs, every result byte-identical to one GPU."""
    path, _ = tiny
    for h in (group, single):
        h.model_load(path)
        h.scale_control(1.0)
    frames = np.stack([synth.synth_frame(160, 120, i) for i in range(12)])
    want = single.advance_batch(frames[:4]) + single.advance_batch(frames[4:8]) + single.advance_batch(frames[8:12])
    tickets = []
    for s in range(3):
        t, view = group.ring_acquire(4, 160, 120)
        view[...] = frames[4 * s:4 * s + 4]
        group.ring_submit(t)
        tickets.append(t)
    devs = []
    for s, t in enumerate(tickets):       # submission order
        r = group.ring_wait(t)
        devs.append(r["device"])
        for i in range(4):
            w = want[4 * s + i]
            assert (r["class_map"][i] == w["class_map"]).all() and (r["decoded_rgba"][i] == w["decoded_rgba"]).all()
            assert (r["frame_rgba"][i] == w["frame_rgba"]).all()
            assert (r["blended_rgba"][i] == oracle.blend_over(w["decoded_rgba"], w["frame_rgba"])).all()
    if len(set(_devices(2))) == 2:
        assert devs == [0, 1, 0]          # ticket t -> devices[(t - 1) % 2], tickets are consecutive
    assert tickets == list(range(tickets[0], tickets[0] + 3))


def test_advance_batch_splits_by_frame_id(group, single, tiny):
    path, _ = tiny
    for h in (group, single):
        h.model_load(path)
        h.scale_control(0.5)
    frames = np.stack([synth.synth_frame(128, 96, 20 + i) for i in range(4)])
    a = group.advance_batch(frames, ids=[7, 8, 9, 10])
    b = single.advance_batch(frames, ids=[7, 8, 9, 10])
    for x, y in zip(a, b):
        assert x["id"] == y["id"] and x["out_w"] == 64 and x["out_h"] == 48
        assert (x["class_map"] == y["class_map"]).all() and (x["decoded_rgba"] == y["decoded_rgba"]).all() and (x["frame_rgba"] == y["frame_rgba"]).all()
    one = group.advance(frames[2], 9)
    assert (one["class_map"] == b[2]["class_map"]).all()
    for h in (group, single):
        h.scale_control(1.0)


@pytest.mark.parametrize("which", ["single", "group"])
def test_frame_level_submit_wait_in_order(which, group, single, tiny):
    """infur_b200_submit / infur_b200_wait: per-frame tickets, results in submission order, partial slots flushed by wait."""
    h = group if which == "group" else single
    path, _ = tiny
    for x in (h, single):
        x.model_load(path)
        x.scale_control(1.0)
    frames = [synth.synth_frame(96, 64, 40 + i) for i in range(11)]   # 11 frames, max_batch 4: the last slots stay partial
    want = [single.advance(f, i + 1) for i, f in enumerate(frames)]
    tickets = [h.submit(f, id=i + 1) for i, f in enumerate(frames)]
    assert tickets == sorted(tickets) and len(set(tickets)) == 11
    for i, t in enumerate(tickets):
        r = h.wait(t)
        assert r["id"] == i + 1 and r["has_decoded"]
        assert (r["class_map"] == want[i]["class_map"]).all() and (r["decoded_rgba"] == want[i]["decoded_rgba"]).all()
        assert (r["frame_rgba"] == want[i]["frame_rgba"]).all()
    with pytest.raises(P.InfurError) as e:
        h.wait(tickets[0])                # already waited
    assert e.value.code == L.E_TICKET
    # the ring is free again: a second stream of frames goes through
    t2 = [h.submit(f) for f in frames[:5]]
    h.flush()
    for i, t in enumerate(t2):
        assert (h.wait(t)["class_map"] == want[i]["class_map"]).all()
    h.flush()
    h.flush()                             # two more wait / flush calls: every lent slot is back in the ring


def test_scale_raised_between_acquire_and_submit(single, tiny):
    """the slot was sized at acquire for factor 0.5; submit runs with factor 2.0 and must grow the slot's
    output buffers instead of writing past them.  Scale changes mid-stream are the reference's normal control flow
    (gui.rs:278-285 -> app.rs:96-98)."""
    path, _ = tiny
    single.model_load(path)
    single.scale_control(0.5)
    frames = np.stack([synth.synth_frame(128, 96, 60 + i) for i in range(2)])
    t, view = single.ring_acquire(2, 128, 96)
    view[...] = frames
    single.scale_control(2.0)
    assert single.is_dirty()
    single.ring_submit(t)
    assert not single.is_dirty()
    r = single.ring_wait(t)
    assert (r["out_w"], r["out_h"]) == (256, 192)
    want = single.advance_batch(frames)
    for i in range(2):
        assert (r["class_map"][i] == want[i]["class_map"]).all() and (r["decoded_rgba"][i] == want[i]["decoded_rgba"]).all()
    single.scale_control(1.0)


def test_ring_release_and_failed_submit_free_the_slot(single, tiny):
    path, _ = tiny
    single.model_load(path)
    single.scale_control(1.0)
    held = [single.ring_acquire(1, 64, 48)[0] for _ in range(3)]
    with pytest.raises(P.InfurError) as e:
        single.ring_acquire(1, 64, 48)
    assert e.value.code == L.E_TICKET
    single.ring_release(held[0])
    with pytest.raises(P.InfurError):
        single.ring_release(held[0])       # no longer a live ticket
    # a submit that fails (Scale makes the output 0-sized) gives its slot back instead of leaking it
    single.scale_control(1e-8)
    with pytest.raises(P.ScaleProcError):
        single.ring_submit(held[1])
    single.scale_control(1.0)
    single.ring_release(held[2])
    again = [single.ring_acquire(1, 64, 48)[0] for _ in range(3)]   # all three slots are free again
    for t in again:
        single.ring_release(t)


def test_advance_device_capacities(single, tiny):
    """the device-resident entry point takes capacities, answers size queries, and reports has_decoded."""
    import torch

    path, _ = tiny
    single.model_load(path)
    single.scale_control(2.0)
    q = single.advance_device_query(2, 64, 48)
    assert (q["out_w"], q["out_h"], q["has_decoded"], q["num_classes"]) == (128, 96, True, 21)
    assert q["required"] == [2 * 128 * 96, 2 * 128 * 96 * 4, 2 * 128 * 96 * 4]
    dev = torch.device("cuda", 0)
    frames = np.stack([synth.synth_frame(64, 48, 70 + i) for i in range(2)])
    d_in = torch.from_numpy(frames).to(dev)
    small_c = torch.empty(2 * 64 * 48, dtype=torch.uint8, device=dev)
    small_d = torch.empty(2 * 64 * 48 * 4, dtype=torch.uint8, device=dev)
    with pytest.raises(P.InfurError) as e:   # sized for scale 1.0: too small for 2.0, nothing is written
        single.advance_device(d_in.data_ptr(), 2, 64, 48, small_c.data_ptr(), small_d.data_ptr())
    assert e.value.code == L.E_BUFFER_TOO_SMALL
    d_c = torch.zeros(q["required"][0], dtype=torch.uint8, device=dev)
    d_d = torch.zeros(q["required"][1], dtype=torch.uint8, device=dev)
    single.advance_device(d_in.data_ptr(), 2, 64, 48, d_c.data_ptr(), d_d.data_ptr(), sync=True, caps=(d_c.numel(), d_d.numel(), 0))
    want = single.advance_batch(frames)
    got_c = d_c.cpu().numpy().reshape(2, 96, 128)
    got_d = d_d.cpu().numpy().reshape(2, 96, 128, 4)
    for i in range(2):
        assert (got_c[i] == want[i]["class_map"]).all() and (got_d[i] == want[i]["decoded_rgba"]).all()
    with pytest.raises(P.InfurError) as e:
        single.advance_device(d_in.data_ptr(), 5, 64, 48, d_c.data_ptr(), d_d.data_ptr())   # n > max_batch
    assert e.value.code == L.E_INVALID_ARG
    single.model_load("")
    q = single.advance_device_query(2, 64, 48)
    assert q["has_decoded"] is False and q["required"] == [0, 0, 0]
    single.scale_control(1.0)


def test_class_legend(single, tiny):
    path, _ = tiny
    single.model_load("")
    assert single.class_legend() is None
    single.model_load(path)
    leg = single.class_legend()
    assert len(leg) == 21 and leg[0][1] == "__background__" and leg[15][1] == "person" and leg[20][1] == "tvmonitor"
    lut = oracle.color_lut()
    for i, _, rgb in leg:
        assert rgb == tuple(int(v) for v in lut[i % 20, 255, :3])   # COLORS_PALETTE[i % 20] (decode_predict.rs:9-34)


def test_tune_table_export_import_skips_measurements(lib, tiny):
    """infur_b200_tune_export / _import: a host can persist the autotune decisions next to the model; a fresh handle that imports
    them builds the same plan without measuring anything (gui.rs:91-103 persists its settings the same way)."""
    path, _ = tiny
    frame = synth.synth_frame(320, 240, 3)
    with P.Handle(device=0, max_batch=2) as h1:
        h1.model_load(path)
        ref = h1.advance(frame, 1)
        ms1, tuned1 = h1.plan_build_stats()
        table = h1.tune_export()
        plan1 = h1.plan_text(1, 320, 240)
    assert tuned1 > 0 and len(table.splitlines()) >= tuned1
    with P.Handle(device=0, max_batch=2) as h2:
        h2.tune_import(table)
        h2.model_load(path)
        got = h2.advance(frame, 1)
        ms2, tuned2 = h2.plan_build_stats()
        plan2 = h2.plan_text(1, 320, 240)
        with pytest.raises(P.InfurError) as e:
            h2.tune_import("64 256 1 1 1 0 1 0 20 100 0\n")        # block_n 100 is not a tile size
        assert e.value.code == L.E_INVALID_ARG
    assert tuned2 == 0          # nothing measured (ms2 < ms1 as a rule, but a wall-clock comparison of two ~10 ms builds is not an invariant)
    assert plan1 == plan2
    assert (got["class_map"] == ref["class_map"]).all() and (got["decoded_rgba"] == ref["decoded_rgba"]).all()
