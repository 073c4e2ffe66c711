"""CPU oracle for the infur hot path  Scale -> Model -> ColorCode.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU stand-in for the reference.  The product path
(``infur_b200/``) never imports this package and fails loudly when its CUDA
library is missing.

PARITY STATUS (see DESIGN.md "Oracle"):
  * pinned by the reference's own tests: the known-answer tests of
    ``infur/src/decode_predict.rs:93-116`` (``color_2``, ``decode_0to1``),
    ``infur/src/processing.rs:288-303`` (zero-size errors) and the output sizes
    asserted in ``infur/src/app.rs:174-252`` and
    ``infur/src/predict_onnx.rs:370-381``; all are ported in
    ``tests/test_oracle_kat.py``.
  * PARITY UNPINNED for numeric values: the reference cannot be built here (no
    Rust toolchain, no onnxruntime, no ffmpeg, no model file) and its tests pin
    no output value of ``fast_image_resize`` (v1, Nearest), ``onnxruntime``
    (unpinned git master) or ``epaint`` 0.19 ``Color32``.  Those three
    third-party algorithms are restated from their published behaviour; each
    function cites the reference call site it stands in for.
"""

from .scale import scaled_size, scale_nearest, scale_bilinear, ScaleError, valid_scale  # noqa: F401
from .preprocess import preprocess_f32, preprocess_u8, norm_lut  # noqa: F401
from .colorcode import (  # noqa: F401
    COLORS_PALETTE,
    color32_from_rgba_unmultiplied,
    color_code,
    color_code_image,
    color_lut,
    frame_rgba,
    softmax_confidence,
    blend_over,
)
from .upsample import bilinear_tables, upsample_bilinear  # noqa: F401
