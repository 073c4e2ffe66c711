"""ctypes binding of ``libinfur_b200.so`` (the C ABI declared in ``include/infur_b200.h``).

There is no Python or CPU fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C infur_b200/csrc``) importing
this module raises, and without a CUDA device ``infur_b200_create`` returns ``INFUR_E_NO_DEVICE``.
"""
from __future__ import annotations

import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libinfur_b200.so")

# status codes (include/infur_b200.h)
OK = 0
E_INVALID_ARG = 1
E_SCALE_NONPOSITIVE = 2
E_ZERO_SIZE_IN = 3
E_ZERO_SIZE_OUT = 4
E_MODEL_LOAD = 5
E_MODEL_INPUT_FORMAT = 6
E_SHAPE = 7
E_RUNTIME = 8
E_BUFFER_TOO_SMALL = 9
E_NO_DEVICE = 10
E_UNSUPPORTED = 11
E_TICKET = 12
E_STREAM_END = 13

RESIZE_NEAREST = 0
RESIZE_BILINEAR = 1
CONV_TCGEN05 = 0
CONV_VALIDATE = 1
CONV_TCGEN05_PAIR = 2   # conv_test only
CONV_TCGEN05_HALO = 3   # conv_test only
CONV_TCGEN05_I8 = 4     # conv_test only: int8 plan form of a quantised layer
CONV_TCGEN05_I8_PAIR = 5
CONV_TCGEN05_PAIR_DEEP = 6   # conv_test only: CTA pair, eight epilogue chunk buffers (needs a residual)
CONV_TCGEN05_I8_PAIR_DEEP = 7
LOAD_DEFAULT = 0
LOAD_SKIP_WEIGHTS = 1
CONF_RAW = 0
CONF_SOFTMAX = 1
ABI_VERSION = 2
MAX_DEVICES = 8


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("max_batch", C.c_int32), ("ring_depth", C.c_int32),
        ("resize_mode", C.c_int32), ("compute_aux", C.c_int32), ("blend", C.c_int32), ("conv_impl", C.c_int32),
        ("use_cuda_graph", C.c_int32), ("autotune", C.c_int32),
        ("num_devices", C.c_int32), ("devices", C.c_int32 * 8), ("frame_rgba", C.c_int32), ("confidence", C.c_int32),
    ]


class Out(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("scaled_bgr", C.c_void_p), ("scaled_bgr_cap", C.c_size_t),
        ("frame_rgba", C.c_void_p), ("frame_rgba_cap", C.c_size_t),
        ("class_map", C.c_void_p), ("class_map_cap", C.c_size_t),
        ("decoded_rgba", C.c_void_p), ("decoded_rgba_cap", C.c_size_t),
        ("blended_rgba", C.c_void_p), ("blended_rgba_cap", C.c_size_t),
        ("logits_f32", C.c_void_p), ("logits_cap", C.c_size_t),
        ("aux_logits_f32", C.c_void_p), ("aux_logits_cap", C.c_size_t),
        ("out_w", C.c_uint32), ("out_h", C.c_uint32), ("num_classes", C.c_uint32), ("has_decoded", C.c_int32),
        ("id", C.c_uint64), ("required", C.c_size_t * 7),
    ]


class Slot(C.Structure):
    _fields_ = [
        ("ticket", C.c_uint64), ("n", C.c_uint32), ("w", C.c_uint32), ("h", C.c_uint32),
        ("out_w", C.c_uint32), ("out_h", C.c_uint32), ("num_classes", C.c_uint32), ("has_decoded", C.c_int32), ("device", C.c_int32),
        ("bgr_in", C.c_void_p), ("class_map", C.c_void_p), ("decoded_rgba", C.c_void_p), ("blended_rgba", C.c_void_p), ("frame_rgba", C.c_void_p),
    ]


class Result(C.Structure):
    _fields_ = [
        ("ticket", C.c_uint64), ("id", C.c_uint64), ("out_w", C.c_uint32), ("out_h", C.c_uint32), ("num_classes", C.c_uint32),
        ("has_decoded", C.c_int32), ("device", C.c_int32),
        ("class_map", C.c_void_p), ("decoded_rgba", C.c_void_p), ("blended_rgba", C.c_void_p), ("frame_rgba", C.c_void_p),
    ]


class DeviceOut(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("d_class_map", C.c_void_p), ("class_map_cap", C.c_size_t),
        ("d_decoded_rgba", C.c_void_p), ("decoded_rgba_cap", C.c_size_t),
        ("d_blended_rgba", C.c_void_p), ("blended_rgba_cap", C.c_size_t),
        ("out_w", C.c_uint32), ("out_h", C.c_uint32), ("num_classes", C.c_uint32), ("has_decoded", C.c_int32),
        ("required", C.c_size_t * 3),
    ]


class ConvDesc(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("n", "h", "w", "cin", "cout", "kh", "kw", "stride", "pad", "dil")] + [
        ("relu", C.c_int32), ("impl", C.c_int32), ("qmul", C.c_void_p)] + [
        (k, C.c_float) for k in ("q_lo", "q_hi", "q_ra", "q_rb", "q_lo2", "q_hi2", "q_deq")] + [("q_zres", C.c_int32), ("q_zout", C.c_int32)]


# every symbol include/infur_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "infur_b200_default_config": (None, [C.POINTER(Config)]),
    "infur_b200_abi_version": (C.c_int32, []),
    "infur_b200_create": (C.c_int32, [C.POINTER(Config), C.POINTER(_H)]),
    "infur_b200_destroy": (None, [_H]),
    "infur_b200_last_error": (C.c_char_p, [_H]),
    "infur_b200_scale_control": (C.c_int32, [_H, C.c_float]),
    "infur_b200_model_load": (C.c_int32, [_H, C.c_char_p]),
    "infur_b200_model_load_bytes": (C.c_int32, [_H, C.c_void_p, C.c_size_t]),
    "infur_b200_model_load_opts": (C.c_int32, [_H, C.c_char_p, C.c_int32]),
    "infur_b200_model_info": (C.c_int32, [_H, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "infur_b200_model_weights_size": (C.c_int32, [_H, C.POINTER(C.c_size_t)]),
    "infur_b200_model_weights_export": (C.c_int32, [_H, C.c_void_p, C.c_size_t]),
    "infur_b200_model_weights_import": (C.c_int32, [_H, C.c_void_p, C.c_size_t]),
    "infur_b200_is_dirty": (C.c_int32, [_H]),
    "infur_b200_advance": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(Out)]),
    "infur_b200_advance_batch": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(Out)]),
    "infur_b200_ring_acquire": (C.c_int32, [_H, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Slot)]),
    "infur_b200_ring_read": (C.c_int32, [_H, C.c_uint64, C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_size_t)]),
    "infur_b200_ring_submit": (C.c_int32, [_H, C.c_uint64]),
    "infur_b200_ring_wait": (C.c_int32, [_H, C.c_uint64, C.POINTER(Slot)]),
    "infur_b200_host_alloc": (C.c_int32, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "infur_b200_host_free": (None, [C.c_void_p]),
    "infur_b200_ring_release": (C.c_int32, [_H, C.c_uint64]),
    "infur_b200_submit": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]),
    "infur_b200_flush": (C.c_int32, [_H]),
    "infur_b200_wait": (C.c_int32, [_H, C.c_uint64, C.POINTER(Result)]),
    "infur_b200_advance_device": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(DeviceOut), C.c_int32]),
    "infur_b200_num_devices": (C.c_int32, [_H]),
    "infur_b200_model_weights_checksum": (C.c_int32, [_H, C.c_int32, C.POINTER(C.c_uint64)]),
    "infur_b200_class_legend": (C.c_int32, [_H, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "infur_b200_profile_step": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "infur_b200_profile_collect": (C.c_int32, [_H, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "infur_b200_tune_export": (C.c_int32, [_H, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "infur_b200_tune_import": (C.c_int32, [_H, C.c_char_p]),
    "infur_b200_plan_build_stats": (C.c_int32, [_H, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "infur_b200_compute_stream": (C.c_void_p, [_H]),
    "infur_b200_launch_count": (C.c_uint64, [_H]),
    "infur_b200_scale_advance": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "infur_b200_model_advance": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                             C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    "infur_b200_model_lowres": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "infur_b200_preprocess": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]),
    "infur_b200_color_code": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "infur_b200_upsample_color": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "infur_b200_color_lut": (C.c_int32, [_H, C.c_void_p, C.c_size_t]),
    "infur_b200_conv_test": (C.c_int32, [_H, C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_float)]),
    "infur_b200_plan_text": (C.c_int32, [_H, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "infur_b200_profile_ops": (C.c_int32, [_H, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_float), C.c_int32,
                                           C.POINTER(C.c_int32)]),
    "infur_b200_onnx_describe": (C.c_int32, [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C infur_b200/csrc` (or __graft_entry__.build()). "
            "infur_b200 has no Python/CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
