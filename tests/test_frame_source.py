"""Host frame source against ff-video's read_frame semantics (ff-video/src/decoder.rs:150-165, processing.rs:116-139).
CPU part: read_exact + 1-based ids + FinishedNormally / ExactReadError classification.  GPU part: the same stream read
straight into pinned ring slots (infur_b200_ring_read) gives the results of advance_batch on the same frames."""
import numpy as np
import pytest

from infur_b200 import frame_source as FS
from infur_b200 import synth


def test_read_frame_ids_and_clean_finish():
    src = FS.spawn_synthetic(64, 48, 3)
    img = src.empty_image()
    for i in range(3):
        assert src.read_frame(img) == i + 1                      # decoder.rs:163-164: 1-based counter
        assert (img == synth.synth_frame(64, 48, i)).all()        # tight bgr24, nothing re-ordered
    with pytest.raises(FS.FinishedNormally):                      # child exited 0 -> FinishedNormally (decoder.rs:157-159)
        src.read_frame(img)
    assert src.frame_counter == 3
    src.close()


def test_failed_decoder_is_exact_read_error():
    src = FS.spawn_synthetic(64, 48, 2, exit_code=3)
    img = src.empty_image()
    assert src.read_frame(img) == 1 and src.read_frame(img) == 2
    with pytest.raises(FS.ExactReadError):                        # non-zero exit -> ExactReadError (decoder.rs:160)
        src.read_frame(img)
    src.close()


def test_truncated_last_frame():
    src = FS.spawn_synthetic(64, 48, 2, exit_code=1, truncate_bytes=100)
    img = src.empty_image()
    assert src.read_frame(img) == 1
    with pytest.raises(FS.ExactReadError):
        src.read_frame(img)
    assert src.frame_counter == 1                                 # a failed read does not advance the counter
    src.close()


@pytest.mark.gpu
def test_stream_into_pinned_ring(handle, tiny):
    path, _ = tiny
    handle.model_load(path)
    handle.scale_control(1.0)
    n_frames, batch, w, h = 11, 4, 160, 120
    frames = np.stack([synth.synth_frame(w, h, i) for i in range(n_frames)])
    ref = handle.advance_batch(frames[:4], ids=[1, 2, 3, 4]) + handle.advance_batch(frames[4:8], ids=[5, 6, 7, 8]) + \
        handle.advance_batch(frames[8:], ids=[9, 10, 11])
    src = FS.spawn_synthetic(w, h, n_frames)
    seen = []
    while True:
        ticket, ids, err = src.read_batch(handle, batch)
        if ticket is not None:
            handle.ring_submit(ticket)
            r = handle.ring_wait(ticket)
            assert r["n"] == len(ids)
            for j, fid in enumerate(ids):
                assert (r["class_map"][j] == ref[fid - 1]["class_map"]).all()
                assert (r["decoded_rgba"][j] == ref[fid - 1]["decoded_rgba"]).all()
            seen += ids
        if err is not None:
            assert isinstance(err, FS.FinishedNormally)
            break
    assert seen == list(range(1, n_frames + 1))                   # 4 + 4 + 3 frames, ids in stream order
    src.close()
