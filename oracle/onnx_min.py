"""Minimal ONNX file reader for the oracle (protobuf wire format decoded by hand; no ``onnx`` package here).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Independent of the product's C++ reader
(infur_b200/csrc/onnx_reader.cpp) and of the fixture writer (infur_b200/onnx_write.py): the parity tests feed the
same ``.onnx`` file to both readers.  Stands in for ONNX Runtime's model loading at ``with_model_from_file``
(infur/src/predict_onnx.rs:288-293).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 11: np.float64}


def _fields(buf: bytes):
    """Yield (field number, wire type, value) of one message; length-delimited values are bytes."""
    i, n = 0, len(buf)
    while i < n:
        key = 0
        shift = 0
        while True:
            b = buf[i]; i += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        f, w = key >> 3, key & 7
        if w == 0:
            v = 0
            shift = 0
            while True:
                b = buf[i]; i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield f, w, v
        elif w == 1:
            yield f, w, buf[i:i + 8]; i += 8
        elif w == 2:
            ln = 0
            shift = 0
            while True:
                b = buf[i]; i += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield f, w, buf[i:i + ln]; i += ln
        elif w == 5:
            yield f, w, buf[i:i + 4]; i += 4
        else:
            raise ValueError("unsupported wire type %d" % w)


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _unpack(buf: bytes):
    out, i = [], 0
    while i < len(buf):
        v = 0
        shift = 0
        while True:
            b = buf[i]; i += 1
            v |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        out.append(v)
    return out


def _tensor(buf: bytes):
    dims, dtype, name, raw, floats, i32, i64 = [], 0, "", None, [], [], []
    for f, w, v in _fields(buf):
        if f == 1:
            dims += [_signed(x) for x in (_unpack(v) if w == 2 else [v])]
        elif f == 2:
            dtype = v
        elif f == 4:
            floats += list(struct.unpack("<%df" % (len(v) // 4), v)) if w == 2 else [struct.unpack("<f", v)[0]]
        elif f == 5:
            i32 += [_signed(x) for x in (_unpack(v) if w == 2 else [v])]
        elif f == 7:
            i64 += [_signed(x) for x in (_unpack(v) if w == 2 else [v])]
        elif f == 8:
            name = v.decode()
        elif f == 9:
            raw = v
    dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=dt).copy()
    elif floats:
        arr = np.array(floats, dtype=dt)
    elif i64:
        arr = np.array(i64, dtype=dt)
    else:
        arr = np.array(i32, dtype=np.int64).astype(dt)
    return name, arr.reshape(dims)


@dataclass
class Node:
    op: str = ""
    name: str = ""
    domain: str = ""
    inputs: list = field(default_factory=list)
    outputs: list = field(default_factory=list)
    attrs: dict = field(default_factory=dict)


def _attr(buf: bytes):
    name, val = "", None
    ints = []
    for f, w, v in _fields(buf):
        if f == 1:
            name = v.decode()
        elif f == 2:
            val = struct.unpack("<f", v)[0]
        elif f == 3:
            val = _signed(v)
        elif f == 4:
            val = v.decode()
        elif f == 5:
            val = _tensor(v)[1]
        elif f == 8:
            ints += [_signed(x) for x in (_unpack(v) if w == 2 else [v])]
    return name, (ints if ints else val)


def _node(buf: bytes) -> Node:
    n = Node()
    for f, w, v in _fields(buf):
        if f == 1:
            n.inputs.append(v.decode())
        elif f == 2:
            n.outputs.append(v.decode())
        elif f == 3:
            n.name = v.decode()
        elif f == 4:
            n.op = v.decode()
        elif f == 5:
            k, a = _attr(v)
            n.attrs[k] = a
        elif f == 7:
            n.domain = v.decode()
    return n


def _value_info(buf: bytes):
    name, elem, dims = "", 0, []
    for f, w, v in _fields(buf):
        if f == 1:
            name = v.decode()
        elif f == 2:
            for f2, _, v2 in _fields(v):
                if f2 != 1:
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        elem = v3
                    elif f3 == 2:
                        for f4, _, v4 in _fields(v3):
                            if f4 != 1:
                                continue
                            d = None
                            for f5, _, v5 in _fields(v4):
                                d = v5 if f5 == 1 else v5.decode()
                            dims.append(d)
    return name, elem, dims


@dataclass
class Graph:
    nodes: list = field(default_factory=list)
    inits: dict = field(default_factory=dict)
    inputs: list = field(default_factory=list)    # (name, elem_type, dims) without initializers
    outputs: list = field(default_factory=list)


def load(path_or_bytes) -> Graph:
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    g = Graph()
    for f, w, v in _fields(data):
        if f != 7 or w != 2:
            continue
        for f2, _, v2 in _fields(v):
            if f2 == 1:
                g.nodes.append(_node(v2))
            elif f2 == 5:
                k, a = _tensor(v2)
                g.inits[k] = a
            elif f2 == 11:
                g.inputs.append(_value_info(v2))
            elif f2 == 12:
                g.outputs.append(_value_info(v2))
    g.inputs = [i for i in g.inputs if i[0] not in g.inits]
    return g
