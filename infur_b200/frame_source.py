"""Host frame source: raw ``bgr24`` frames from a pipe, straight into pinned ring slots.

Mirror of the step immediately before the hot path: ``FFMpegDecoder::read_frame``
(ff-video/src/decoder.rs:150-165) as driven by ``VideoPlayer::advance`` (infur/src/processing.rs:116-139):

* the decoder process writes tightly packed ``w*h*3`` byte frames to its stdout (``-f image2pipe -pix_fmt bgr24
  -c:v rawvideo pipe:1``, decoder.rs:51-75);
* ``read_frame`` does ``read_exact`` of one frame; on failure the error is ``FinishedNormally`` when the child has
  exited with status 0 and ``ExactReadError`` otherwise (decoder.rs:156-162); on success the 1-based frame counter
  advances and is the frame's id (decoder.rs:163-164);
* ``VideoPlayer::advance`` closes the video on ``FinishedNormally`` and propagates the error (processing.rs:133-136).

ffmpeg is not part of this repo's environment; any process that writes raw frames works (the tests and
``python -m infur_b200.frame_source W H N`` use the synthetic generator of ``synth.py``).  Reading bytes is I/O, not
compute: this module has no GPU dependency of its own; ``read_batch`` hands the descriptor to
``infur_b200_ring_read`` so the bytes land in pinned memory without a staging copy.
"""
from __future__ import annotations

import os
import subprocess
import sys
from typing import List, Optional, Tuple

import numpy as np


class VideoProcError(Exception):
    """ff-video/src/error.rs:10-40"""


class FinishedNormally(VideoProcError):
    def __str__(self):
        return "finished normally"


class ExactReadError(VideoProcError):
    def __str__(self):
        return "couldn't read an entire image"


class RawVideoSource:
    """A running decoder process (or any readable fd) producing ``width x height`` bgr24 frames."""

    def __init__(self, width: int, height: int, proc: Optional[subprocess.Popen] = None, fd: Optional[int] = None):
        if (proc is None) == (fd is None):
            raise ValueError("give exactly one of proc / fd")
        self.width, self.height = width, height
        self.proc = proc
        self.fd = proc.stdout.fileno() if proc is not None else fd
        self.frame_counter = 0          # decoder.rs:163: ids are 1-based

    @property
    def frame_bytes(self) -> int:
        return self.width * self.height * 3

    def empty_image(self) -> np.ndarray:
        """decoder.rs:150-153"""
        return np.zeros((self.height, self.width, 3), np.uint8)

    def _classify(self) -> VideoProcError:
        # decoder.rs:157-161: Ok(Some(status)) with code 0 -> FinishedNormally, anything else -> ExactReadError
        if self.proc is not None:
            try:
                code = self.proc.wait(timeout=0.5)   # the reference polls try_wait; EOF normally means the child is gone
            except subprocess.TimeoutExpired:
                code = None
            return FinishedNormally() if code == 0 else ExactReadError()
        return FinishedNormally() if getattr(self, "_fd_clean_eof", False) else ExactReadError()

    def read_frame(self, image: np.ndarray) -> int:
        """``read_frame`` (decoder.rs:156-165): fill ``image`` (``[H][W][3]`` u8, contiguous) and return its id."""
        buf = memoryview(image).cast("B")
        if len(buf) != self.frame_bytes:
            raise ValueError("image buffer has the wrong size")
        have = 0
        while have < len(buf):
            n = os.readv(self.fd, [buf[have:]])
            if n == 0:
                self._fd_clean_eof = have == 0
                raise self._classify()
            have += n
        self.frame_counter += 1
        return self.frame_counter

    def read_batch(self, handle, batch: int) -> Tuple[Optional[int], List[int], Optional[VideoProcError]]:
        """Up to ``batch`` frames into the next pinned ring slot.  Returns (ticket or None, ids, end-of-stream error or None);
        the caller submits the ticket when it is not None and stops the stream when the error is not None."""
        ticket, _ = handle.ring_acquire(batch, self.width, self.height)
        got, partial, ended = handle.ring_read(ticket, self.fd)
        ids = list(range(self.frame_counter + 1, self.frame_counter + got + 1))
        self.frame_counter += got
        err = None
        if ended:
            self._fd_clean_eof = partial == 0
            err = self._classify()
        return (ticket if got else None), ids, err

    def close(self):
        if self.proc is not None:
            try:
                self.proc.stdout.close()
            except Exception:
                pass
            if self.proc.poll() is None:
                self.proc.terminate()
            self.proc.wait()


def spawn_synthetic(width: int, height: int, frames: int, exit_code: int = 0, truncate_bytes: int = 0) -> RawVideoSource:
    """A child process standing in for ``ffmpeg ... -f lavfi -i testsrc`` (infur-test-gen/build.rs:12-31): writes
    ``frames`` synthetic frames (``synth.synth_frame(w, h, i)``) to its stdout, optionally cutting the last one short,
    then exits with ``exit_code``."""
    cmd = [sys.executable, "-m", "infur_b200.frame_source", str(width), str(height), str(frames), str(exit_code), str(truncate_bytes)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, env=env, bufsize=0)
    return RawVideoSource(width, height, proc=proc)


def _main(argv):
    from infur_b200 import synth

    w, h, n = int(argv[1]), int(argv[2]), int(argv[3])
    code = int(argv[4]) if len(argv) > 4 else 0
    cut = int(argv[5]) if len(argv) > 5 else 0
    out = sys.stdout.buffer
    for i in range(n):
        data = synth.synth_frame(w, h, i).tobytes()
        if cut and i == n - 1:
            data = data[: len(data) - cut]
        out.write(data)
    out.flush()
    sys.exit(code)


if __name__ == "__main__":
    _main(sys.argv)
