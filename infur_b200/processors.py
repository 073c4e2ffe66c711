"""Host-side mirror of infur's plug-in surface over the C ABI.

The reference's stages all implement ``trait Processor`` (infur/src/processing.rs:23-60):
``control(cmd)``, ``advance(input, out)``, ``is_dirty()``, ``generate()``.  The classes here keep
those names, argument meaning and error behaviour so that the parity tests read like the
reference's own tests:

* :class:`Scale`       <- ``Scale``       (processing.rs:179-282)
* :class:`Model`       <- ``Model``       (predict_onnx.rs:146-346)
* :class:`ColorCode`   <- ``ColorCode``   (decode_predict.rs:38-84)
* :class:`GpuPipeline` <- the ``scale -> model -> decoder`` part of ``ProcessingApp`` (app.rs:53-158)

All arithmetic runs in ``libinfur_b200.so`` on the GPU; this file only marshals numpy buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib as L


# --------------------------------------------------------------------------- errors
class InfurError(Exception):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class ValidScaleError(InfurError):
    """processing.rs:145-168"""


class ScaleProcError(InfurError):
    """processing.rs:201-211; ``kind`` is ``"ZeroSizeIn"`` or ``"ZeroSizeOut"``"""

    @property
    def kind(self) -> str:
        return {L.E_ZERO_SIZE_IN: "ZeroSizeIn", L.E_ZERO_SIZE_OUT: "ZeroSizeOut"}.get(self.code, "Other")


class ModelCmdError(InfurError):
    """predict_onnx.rs:41-54"""


class ModelProcError(InfurError):
    """predict_onnx.rs:32-39"""


def _raise(code: int, msg: str):
    if code == L.E_SCALE_NONPOSITIVE:
        raise ValidScaleError(code, msg)
    if code in (L.E_ZERO_SIZE_IN, L.E_ZERO_SIZE_OUT):
        raise ScaleProcError(code, msg)
    if code in (L.E_MODEL_LOAD, L.E_MODEL_INPUT_FORMAT):
        raise ModelCmdError(code, msg)
    if code in (L.E_SHAPE, L.E_RUNTIME):
        raise ModelProcError(code, msg)
    raise InfurError(code, msg)


# --------------------------------------------------------------------------- data types
@dataclass
class Frame:
    """processing.rs:9-18 -- equality is on ``id`` only."""

    id: int
    img: np.ndarray  # [H][W][3] u8, B,G,R (image-ext/src/image_bgr.rs:7-11)

    def __eq__(self, other):
        return isinstance(other, Frame) and self.id == other.id


@dataclass
class ModelInfo:
    """predict_onnx.rs:56-62"""

    input_names: list
    input0_dtype: str
    output_names: list


@dataclass
class GUIFrame:
    """app.rs:65-69 -- ``buffer``/``decoded_buffer`` are ``[H][W][4]`` u8 RGBA (``ColorImage.pixels``)."""

    id: int
    buffer: np.ndarray
    decoded_buffer: Optional[np.ndarray]
    class_map: Optional[np.ndarray] = None
    blended: Optional[np.ndarray] = None

    @property
    def size(self):  # ColorImage.size = [w, h]
        return [self.buffer.shape[1], self.buffer.shape[0]]


class PinnedArray:
    """A numpy array over page-locked host memory from ``infur_b200_host_alloc`` (freed with the object)."""

    def __init__(self, shape, dtype=np.uint8):
        self.lib = L.load()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = self.lib.infur_b200_host_alloc(n, C.byref(p))
        if rc != L.OK:
            raise InfurError(rc, self.lib.infur_b200_last_error(None).decode())
        self._p = p
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(n, 1),))[:n].view(dtype).reshape(shape)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            self.lib.infur_b200_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_bgr(img: np.ndarray) -> np.ndarray:
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise TypeError("expected an [H][W][3] uint8 BGR image")
    return np.ascontiguousarray(img)


# --------------------------------------------------------------------------- handle
class Handle:
    """Owner of one ``infur_b200_handle``: one GPU, or -- ``devices=[...]`` -- several GPUs of the box behind one handle
    (one owner thread either way; the library runs one worker thread per GPU)."""

    def __init__(self, device: int = 0, max_batch: int = 8, ring_depth: int = 3, compute_aux: bool = False, blend: bool = False,
                 conv_impl: int = L.CONV_TCGEN05, resize_mode: int = L.RESIZE_NEAREST, devices=None, frame_rgba: bool = True,
                 use_cuda_graph: bool = True, confidence: int = L.CONF_RAW, autotune: bool = True):
        self.lib = L.load()
        cfg = L.Config()
        self.lib.infur_b200_default_config(C.byref(cfg))
        cfg.device, cfg.max_batch, cfg.ring_depth = device, max_batch, ring_depth
        cfg.compute_aux, cfg.blend, cfg.conv_impl = int(compute_aux), int(blend), conv_impl
        cfg.resize_mode = resize_mode
        cfg.frame_rgba, cfg.use_cuda_graph, cfg.confidence, cfg.autotune = int(frame_rgba), int(use_cuda_graph), confidence, int(autotune)
        if devices is not None:
            devices = list(devices)
            if not 1 <= len(devices) <= L.MAX_DEVICES:
                raise ValueError("devices must list 1..8 CUDA ordinals")
            cfg.num_devices = len(devices)
            for i, d in enumerate(devices):
                cfg.devices[i] = d
            cfg.device = devices[0]
        self.cfg = cfg
        self._h = C.c_void_p()
        rc = self.lib.infur_b200_create(C.byref(cfg), C.byref(self._h))
        if rc != L.OK:
            msg = self.lib.infur_b200_last_error(None).decode()
            self._h = None
            raise InfurError(rc, msg)

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self.lib.infur_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != L.OK:
            _raise(rc, self.lib.infur_b200_last_error(self._h).decode())

    # -- control
    def scale_control(self, factor: float):
        self._check(self.lib.infur_b200_scale_control(self._h, C.c_float(factor)))

    def is_dirty(self) -> bool:
        return bool(self.lib.infur_b200_is_dirty(self._h))

    def model_load(self, path: str, skip_weights: bool = False):
        self._check(self.lib.infur_b200_model_load_opts(self._h, path.encode(), L.LOAD_SKIP_WEIGHTS if skip_weights else L.LOAD_DEFAULT))

    def model_load_bytes(self, data: bytes):
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        self._check(self.lib.infur_b200_model_load_bytes(self._h, buf, len(data)))

    def model_info(self) -> Optional[ModelInfo]:
        need = C.c_size_t()
        rc = self.lib.infur_b200_model_info(self._h, None, 0, C.byref(need))
        if rc == L.E_INVALID_ARG:
            return None
        buf = C.create_string_buffer(need.value)
        self._check(self.lib.infur_b200_model_info(self._h, buf, need.value, C.byref(need)))
        name, dtype, outs = buf.value.decode().split("\t")
        return ModelInfo([name], dtype, outs.split(",") if outs else [])

    def weights_size(self) -> int:
        n = C.c_size_t()
        self._check(self.lib.infur_b200_model_weights_size(self._h, C.byref(n)))
        return n.value

    def weights_export(self, dptr: int, nbytes: int):
        self._check(self.lib.infur_b200_model_weights_export(self._h, C.c_void_p(dptr), nbytes))

    def weights_import(self, dptr: int, nbytes: int):
        self._check(self.lib.infur_b200_model_weights_import(self._h, C.c_void_p(dptr), nbytes))

    # -- advance (host buffers)
    def advance_batch(self, frames: np.ndarray, ids=None, want=("frame_rgba", "class_map", "decoded_rgba"), into=None) -> list:
        """``frames``: ``[N][H][W][3]`` u8 BGR.  Returns one dict per frame.  ``into``: optional list (one dict per frame) of
        preallocated output arrays by name (e.g. views of a :class:`PinnedArray`), re-used across calls like the reference's `out`."""
        frames = np.ascontiguousarray(frames)
        if frames.dtype != np.uint8 or frames.ndim != 4 or frames.shape[3] != 3:
            raise TypeError("expected [N][H][W][3] uint8")
        n, h, w = frames.shape[:3]
        ids_arr = (C.c_uint64 * n)(*(ids if ids is not None else range(1, n + 1)))
        outs = (L.Out * n)()
        for o in outs:
            o.struct_size = C.sizeof(L.Out)
        # first call without buffers: sizes (the C analogue of "stage re-allocates Out on size change")
        rc = self.lib.infur_b200_advance_batch(self._h, frames.ctypes.data, n, w, h, ids_arr, outs)
        self._check(rc)
        ow, oh, k, has = outs[0].out_w, outs[0].out_h, outs[0].num_classes, outs[0].has_decoded
        res = []
        keep = []
        spec = {"scaled_bgr": ((oh, ow, 3), np.uint8, True), "frame_rgba": ((oh, ow, 4), np.uint8, True),
                "class_map": ((oh, ow), np.uint8, has), "decoded_rgba": ((oh, ow, 4), np.uint8, has),
                "blended_rgba": ((oh, ow, 4), np.uint8, has), "logits_f32": ((k, oh, ow), np.float32, has),
                "aux_logits_f32": ((k, oh, ow), np.float32, has)}
        capname = {"logits_f32": "logits_cap", "aux_logits_f32": "aux_logits_cap"}
        for i in range(n):
            d = {"id": int(ids_arr[i]), "out_w": ow, "out_h": oh, "num_classes": k, "has_decoded": bool(has)}
            for name in want:
                shape, dt, ok = spec[name]
                if not ok:
                    d[name] = None
                    continue
                arr = into[i][name] if into is not None and name in into[i] else np.empty(shape, dtype=dt)
                if arr.shape != tuple(shape) or arr.dtype != dt or not arr.flags.c_contiguous:
                    raise ValueError(f"preallocated '{name}' must be a C-contiguous {dt.__name__} array of shape {tuple(shape)}")
                setattr(outs[i], name, arr.ctypes.data)
                setattr(outs[i], capname.get(name, name + "_cap"), arr.nbytes)
                d[name] = arr
                keep.append(arr)
            res.append(d)
        if ow * oh:
            self._check(self.lib.infur_b200_advance_batch(self._h, frames.ctypes.data, n, w, h, ids_arr, outs))
        return res

    def advance(self, img: np.ndarray, id: int = 0, want=("frame_rgba", "class_map", "decoded_rgba"), into=None) -> dict:
        return self.advance_batch(_check_bgr(img)[None], [id], want, into=[into] if into is not None else None)[0]

    # -- single stages
    def scale_advance(self, img: np.ndarray) -> np.ndarray:
        img = _check_bgr(img)
        h, w = img.shape[:2]
        ow, oh = C.c_uint32(), C.c_uint32()
        rc = self.lib.infur_b200_scale_advance(self._h, img.ctypes.data, w, h, None, 0, C.byref(ow), C.byref(oh))
        if rc not in (L.OK, L.E_BUFFER_TOO_SMALL):
            self._check(rc)
        out = np.empty((oh.value, ow.value, 3), dtype=np.uint8)
        if out.size:
            self._check(self.lib.infur_b200_scale_advance(self._h, img.ctypes.data, w, h, out.ctypes.data, out.nbytes, C.byref(ow), C.byref(oh)))
        return out

    def preprocess(self, img: np.ndarray) -> np.ndarray:
        img = _check_bgr(img)
        h, w = img.shape[:2]
        out = np.empty((3, h, w), dtype=np.float32)
        self._check(self.lib.infur_b200_preprocess(self._h, img.ctypes.data, w, h, out.ctypes.data, out.nbytes))
        return out

    def model_advance(self, img: np.ndarray, out=None):
        """``Model::advance`` alone: list of ``[K][H][W]`` f32 (``out`` and, with compute_aux, ``aux``); ``out`` untouched without a model."""
        img = _check_bgr(img)
        h, w = img.shape[:2]
        info = self.model_info()
        if info is None:
            return out
        k, has = C.c_uint32(), C.c_int32()
        # K is a property of the model: query it with empty buffers first
        rc = self.lib.infur_b200_model_advance(self._h, img.ctypes.data, w, h, None, 0, None, 0, C.byref(k), C.byref(has))
        self._check(rc)
        want_aux = bool(self.cfg.compute_aux) and len(info.output_names) > 1
        lg = np.empty((k.value, h, w), dtype=np.float32)
        aux = np.empty((k.value, h, w), dtype=np.float32) if want_aux else None
        self._check(self.lib.infur_b200_model_advance(self._h, img.ctypes.data, w, h, lg.ctypes.data, lg.nbytes,
                                                      aux.ctypes.data if want_aux else None, aux.nbytes if want_aux else 0,
                                                      C.byref(k), C.byref(has)))
        return [lg, aux] if want_aux else [lg]

    def model_lowres(self, img: np.ndarray) -> Optional[np.ndarray]:
        """Diagnostics: the ``out`` head's logits before the final Resize, ``[K][h/8][w/8]`` f32 (None without a model)."""
        img = _check_bgr(img)
        hgt, w = img.shape[:2]
        k, lw, lh = C.c_uint32(), C.c_uint32(), C.c_uint32()
        rc = self.lib.infur_b200_model_lowres(self._h, img.ctypes.data, w, hgt, None, 0, C.byref(k), C.byref(lw), C.byref(lh))
        if k.value == 0:
            self._check(rc)
            return None
        out = np.empty((k.value, lh.value, lw.value), dtype=np.float32)
        self._check(self.lib.infur_b200_model_lowres(self._h, img.ctypes.data, w, hgt, out.ctypes.data, out.size, C.byref(k), C.byref(lw), C.byref(lh)))
        return out

    def color_code(self, hm: np.ndarray):
        hm = np.ascontiguousarray(hm, dtype=np.float32)
        k, h, w = hm.shape
        rgba = np.zeros((h, w, 4), dtype=np.uint8)
        cls = np.zeros((h, w), dtype=np.uint8)
        self._check(self.lib.infur_b200_color_code(self._h, hm.ctypes.data, k, w, h, rgba.ctypes.data, cls.ctypes.data))
        return cls, rgba

    def upsample_color(self, lowres: np.ndarray, out_h: int, out_w: int, frame_bgr: Optional[np.ndarray] = None, want_logits: bool = False):
        lowres = np.ascontiguousarray(lowres, dtype=np.float32)
        k, lh, lw = lowres.shape
        cls = np.zeros((out_h, out_w), dtype=np.uint8)
        dec = np.zeros((out_h, out_w, 4), dtype=np.uint8)
        bl = np.zeros((out_h, out_w, 4), dtype=np.uint8) if frame_bgr is not None else None
        lg = np.zeros((k, out_h, out_w), dtype=np.float32) if want_logits else None
        fb = _check_bgr(frame_bgr) if frame_bgr is not None else None
        self._check(self.lib.infur_b200_upsample_color(
            self._h, lowres.ctypes.data, k, lw, lh, out_w, out_h, fb.ctypes.data if fb is not None else None, cls.ctypes.data,
            dec.ctypes.data, bl.ctypes.data if bl is not None else None, lg.ctypes.data if lg is not None else None))
        return {"class_map": cls, "decoded_rgba": dec, "blended_rgba": bl, "logits": lg}

    def color_lut(self) -> np.ndarray:
        lut = np.zeros((20, 256, 4), dtype=np.uint8)
        self._check(self.lib.infur_b200_color_lut(self._h, lut.ctypes.data, lut.nbytes))
        return lut

    # -- diagnostics
    def conv_test(self, x, w, bias, residual=None, stride=1, pad=0, dil=1, relu=False, impl=L.CONV_TCGEN05, f32_out=False, timed=False,
                  quant=None):
        """x: [N][H][W][Cin] fp16, w: [Cout][kh][kw][Cin] fp16, bias f32 [Cout], residual like the output.
        ``quant``: dict(qmul=[Cout] f32, q_lo, q_hi[, q_ra, q_rb, q_lo2, q_hi2][, q_deq]) runs the layer as a quantised one."""
        x = np.ascontiguousarray(x, dtype=np.float16)
        w = np.ascontiguousarray(w, dtype=np.float16)
        bias = np.ascontiguousarray(bias, dtype=np.float32)
        n, h, wd, cin = x.shape
        cout, kh, kw, _ = w.shape
        oh = (h + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        ow = (wd + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        d = L.ConvDesc(n, h, wd, cin, cout, kh, kw, stride, pad, dil, int(relu), impl)
        if quant is not None:
            qmul = np.ascontiguousarray(quant["qmul"], dtype=np.float32)
            assert qmul.shape == (cout,)
            d.qmul = qmul.ctypes.data
            for k in ("q_lo", "q_hi", "q_ra", "q_rb", "q_lo2", "q_hi2", "q_deq"):
                setattr(d, k, float(quant.get(k, 0.0)))
            d.q_zres, d.q_zout = int(quant.get("q_zres", 0)), int(quant.get("q_zout", 0))
        y = np.zeros((n, oh, ow, cout), dtype=np.float32 if f32_out else np.float16)
        res = np.ascontiguousarray(residual, dtype=np.float16) if residual is not None else None
        ms = C.c_float()
        self._check(self.lib.infur_b200_conv_test(
            self._h, C.byref(d), x.ctypes.data, w.ctypes.data, bias.ctypes.data, res.ctypes.data if res is not None else None,
            None if f32_out else y.ctypes.data, y.ctypes.data if f32_out else None, C.byref(ms) if timed else None))
        return (y, ms.value) if timed else y

    def plan_text(self, n: int, w: int, h: int) -> str:
        need = C.c_size_t()
        rc = self.lib.infur_b200_plan_text(self._h, n, w, h, None, 0, C.byref(need))
        if rc not in (L.OK, L.E_BUFFER_TOO_SMALL):
            self._check(rc)
        buf = C.create_string_buffer(need.value)
        self._check(self.lib.infur_b200_plan_text(self._h, n, w, h, buf, need.value, C.byref(need)))
        return buf.value.decode()

    def launch_count(self) -> int:
        return int(self.lib.infur_b200_launch_count(self._h))

    def compute_stream(self) -> int:
        return int(self.lib.infur_b200_compute_stream(self._h) or 0)

    # -- device-resident and ring paths
    def advance_device(self, d_bgr: int, n: int, w: int, h: int, d_class: int, d_decoded: int, d_blended: int = 0, sync: bool = False,
                       caps=None):
        """Device buffers in and out.  ``caps`` = (class_map, decoded, blended) capacities in bytes; default: exactly what an
        output of the INPUT size needs (scale 1.0) -- a larger Scale factor then fails with E_BUFFER_TOO_SMALL."""
        o = L.DeviceOut()
        o.struct_size = C.sizeof(L.DeviceOut)
        px = n * w * h
        caps = caps or (px, px * 4, px * 4)
        o.d_class_map, o.class_map_cap = d_class or None, caps[0]
        o.d_decoded_rgba, o.decoded_rgba_cap = d_decoded or None, caps[1]
        o.d_blended_rgba, o.blended_rgba_cap = d_blended or None, caps[2]
        self._check(self.lib.infur_b200_advance_device(self._h, C.c_void_p(d_bgr), n, w, h, C.byref(o), int(sync)))
        return o.out_w, o.out_h

    def advance_device_query(self, n: int, w: int, h: int) -> dict:
        """Size query of the device-resident path: nothing runs."""
        o = L.DeviceOut()
        o.struct_size = C.sizeof(L.DeviceOut)
        self._check(self.lib.infur_b200_advance_device(self._h, None, n, w, h, C.byref(o), 0))
        return {"out_w": o.out_w, "out_h": o.out_h, "num_classes": o.num_classes, "has_decoded": bool(o.has_decoded), "required": list(o.required)}

    def profile_step(self, d_bgr: int, n: int, w: int, h: int):
        self._check(self.lib.infur_b200_profile_step(self._h, C.c_void_p(d_bgr), n, w, h))

    def profile_collect(self):
        cap = 256
        ms = (C.c_float * cap)()
        cnt, steps = C.c_int32(), C.c_int32()
        self._check(self.lib.infur_b200_profile_collect(self._h, ms, cap, C.byref(cnt), C.byref(steps)))
        return [ms[i] for i in range(cnt.value)], steps.value

    def plan_build_stats(self):
        ms, tuned = C.c_float(), C.c_int32()
        self._check(self.lib.infur_b200_plan_build_stats(self._h, C.byref(ms), C.byref(tuned)))
        return ms.value, tuned.value

    def tune_export(self) -> str:
        need = C.c_size_t()
        self.lib.infur_b200_tune_export(self._h, None, 0, C.byref(need))
        buf = C.create_string_buffer(max(need.value, 1))
        self._check(self.lib.infur_b200_tune_export(self._h, buf, need.value, C.byref(need)))
        return buf.value.decode()

    def tune_import(self, text: str):
        self._check(self.lib.infur_b200_tune_import(self._h, text.encode()))

    def num_devices(self) -> int:
        return int(self.lib.infur_b200_num_devices(self._h))

    def weights_checksum(self, index: int = 0) -> int:
        v = C.c_uint64()
        self._check(self.lib.infur_b200_model_weights_checksum(self._h, index, C.byref(v)))
        return v.value

    def class_legend(self):
        """[(index, label, (r, g, b))] for the loaded model's classes, or None without a model."""
        need = C.c_size_t()
        rc = self.lib.infur_b200_class_legend(self._h, None, 0, C.byref(need))
        if rc == L.E_INVALID_ARG:
            return None
        buf = C.create_string_buffer(need.value)
        self._check(self.lib.infur_b200_class_legend(self._h, buf, need.value, C.byref(need)))
        out = []
        for ln in buf.value.decode().splitlines():
            i, label, rgb = ln.split("\t")
            out.append((int(i), label, tuple(int(v) for v in rgb.split(","))))
        return out

    def profile_ops(self, d_bgr: int, n: int, w: int, h: int, iters: int = 3):
        cap = 256
        ms = (C.c_float * cap)()
        cnt = C.c_int32()
        self._check(self.lib.infur_b200_profile_ops(self._h, C.c_void_p(d_bgr), n, w, h, iters, ms, cap, C.byref(cnt)))
        return [ms[i] for i in range(cnt.value)]

    def ring_acquire(self, n: int, w: int, h: int):
        s = L.Slot()
        self._check(self.lib.infur_b200_ring_acquire(self._h, n, w, h, C.byref(s)))
        view = np.ctypeslib.as_array(C.cast(s.bgr_in, C.POINTER(C.c_uint8)), shape=(n, h, w, 3)) if n * w * h else np.zeros((n, h, w, 3), np.uint8)
        return s.ticket, view

    def ring_read(self, ticket: int, fd: int):
        """Fill the slot from file descriptor ``fd`` (raw bgr24 frames).  Returns (frames_read, partial_bytes, ended)."""
        got, part = C.c_uint32(), C.c_size_t()
        rc = self.lib.infur_b200_ring_read(self._h, ticket, fd, C.byref(got), C.byref(part))
        if rc not in (L.OK, L.E_STREAM_END):
            self._check(rc)
        return got.value, part.value, rc == L.E_STREAM_END

    def ring_submit(self, ticket: int):
        self._check(self.lib.infur_b200_ring_submit(self._h, ticket))

    def ring_wait(self, ticket: int) -> dict:
        s = L.Slot()
        self._check(self.lib.infur_b200_ring_wait(self._h, ticket, C.byref(s)))
        n, oh, ow = s.n, s.out_h, s.out_w
        out = {"n": n, "out_w": ow, "out_h": oh, "has_decoded": bool(s.has_decoded), "num_classes": s.num_classes, "device": s.device,
               "class_map": None, "decoded_rgba": None, "blended_rgba": None, "frame_rgba": None}
        if s.has_decoded and n * oh * ow:
            out["class_map"] = np.ctypeslib.as_array(C.cast(s.class_map, C.POINTER(C.c_uint8)), shape=(n, oh, ow))
            out["decoded_rgba"] = np.ctypeslib.as_array(C.cast(s.decoded_rgba, C.POINTER(C.c_uint8)), shape=(n, oh, ow, 4))
            if s.blended_rgba:
                out["blended_rgba"] = np.ctypeslib.as_array(C.cast(s.blended_rgba, C.POINTER(C.c_uint8)), shape=(n, oh, ow, 4))
        if s.frame_rgba and n * oh * ow:
            out["frame_rgba"] = np.ctypeslib.as_array(C.cast(s.frame_rgba, C.POINTER(C.c_uint8)), shape=(n, oh, ow, 4))
        return out

    def ring_release(self, ticket: int):
        self._check(self.lib.infur_b200_ring_release(self._h, ticket))

    # -- frame-level asynchronous API (what a "Proc" thread calls per frame)
    def submit(self, img: np.ndarray, id: int = 0) -> int:
        img = _check_bgr(img)
        h, w = img.shape[:2]
        t = C.c_uint64()
        self._check(self.lib.infur_b200_submit(self._h, img.ctypes.data, w, h, id, C.byref(t)))
        return t.value

    def flush(self):
        self._check(self.lib.infur_b200_flush(self._h))

    def wait(self, ticket: int) -> dict:
        """Result of one submitted frame; the arrays are views into pinned library memory (valid until the second-next wait)."""
        r = L.Result()
        self._check(self.lib.infur_b200_wait(self._h, ticket, C.byref(r)))
        oh, ow = r.out_h, r.out_w
        out = {"id": r.id, "out_w": ow, "out_h": oh, "has_decoded": bool(r.has_decoded), "num_classes": r.num_classes, "device": r.device,
               "class_map": None, "decoded_rgba": None, "blended_rgba": None, "frame_rgba": None}

        def view(p, shape):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=shape) if p and oh * ow else None

        out["class_map"] = view(r.class_map, (oh, ow))
        out["decoded_rgba"] = view(r.decoded_rgba, (oh, ow, 4))
        out["blended_rgba"] = view(r.blended_rgba, (oh, ow, 4))
        out["frame_rgba"] = view(r.frame_rgba, (oh, ow, 4))
        return out


# --------------------------------------------------------------------------- Processor mirrors
class Processor:
    """trait Processor (processing.rs:23-60)."""

    def control(self, cmd):
        raise NotImplementedError

    def advance(self, inp, out):
        raise NotImplementedError

    def is_dirty(self) -> bool:
        raise NotImplementedError

    def generate(self):
        """processing.rs:53-59: ``advance(&(), &mut ())`` for source-like processors."""
        return self.advance(None, None)


class Scale(Processor):
    """``Scale`` (processing.rs:179-282): Command = f32, Input/Output = Option<Frame>.

    ``advance(input, out)`` returns the new ``out`` (Python has no ``&mut Option``)."""

    def __init__(self, handle: Handle):
        self.h = handle

    def control(self, cmd: float):
        self.h.scale_control(cmd)
        return self

    def is_dirty(self) -> bool:
        return self.h.is_dirty()

    def advance(self, inp: Optional[Frame], out: Optional[Frame] = None) -> Optional[Frame]:
        if inp is None:
            self.h.lib.infur_b200_scale_advance(self.h._h, None, 0, 0, None, 0, None, None)  # clears dirty (processing.rs:233-237)
            return out
        return Frame(inp.id, self.h.scale_advance(inp.img))


class Model(Processor):
    """``Model`` (predict_onnx.rs:146-346): Command = ModelCmd::Load(path), Input = BgrImage,
    Output = Vec<ArrayD<f32>> (here: list of ``[K][H][W]`` f32 arrays)."""

    def __init__(self, handle: Handle):
        self.h = handle

    def control(self, path: str):
        self.h.model_load(path)
        return self

    def is_dirty(self) -> bool:
        return False  # predict_onnx.rs:336-338

    def get_info(self) -> Optional[ModelInfo]:
        return self.h.model_info()

    def advance(self, img: np.ndarray, out: Optional[list] = None) -> Optional[list]:
        return self.h.model_advance(img, out)


class ColorCode(Processor):
    """``ColorCode`` (decode_predict.rs:38-84): Input = Array3<f32> ``[K][H][W]``, Output = Option<ColorImage>."""

    def __init__(self, handle: Handle):
        self.h = handle

    def control(self, cmd=None):
        return self

    def is_dirty(self) -> bool:
        return False  # decode_predict.rs:81-83

    def advance(self, hm: np.ndarray, out=None) -> np.ndarray:
        _, rgba = self.h.color_code(hm)
        return rgba


class GpuPipeline(Processor):
    """The ``scale -> model -> decoder`` section of ``ProcessingApp`` (app.rs:53-158) as ONE fused GPU call.

    ``control(("Scale", f))`` / ``control(("Model", path))`` mirror ``AppCmd::Scale`` / ``AppCmd::Model``
    (app.rs:39-51,91-105); ``advance(frame)`` mirrors app.rs:109-149 and returns a :class:`GUIFrame`."""

    def __init__(self, handle: Optional[Handle] = None, **kw):
        self.h = handle or Handle(**kw)
        self.scale = Scale(self.h)
        self.model = Model(self.h)
        self.decoder = ColorCode(self.h)

    def control(self, cmd):
        kind, arg = cmd
        if kind == "Scale":
            self.scale.control(arg)
        elif kind == "Model":
            self.model.control(arg)
        else:
            raise ValueError(f"unknown command {kind!r}")
        return self

    def is_dirty(self) -> bool:
        return self.scale.is_dirty()

    def advance(self, frame: Optional[Frame], out=None) -> Optional[GUIFrame]:
        if frame is None:
            return None
        want = ["frame_rgba", "class_map", "decoded_rgba"] + (["blended_rgba"] if self.h.cfg.blend else [])
        r = self.h.advance(frame.img, id=frame.id, want=tuple(want))
        return GUIFrame(frame.id, r["frame_rgba"], r["decoded_rgba"], r["class_map"], r.get("blended_rgba"))
