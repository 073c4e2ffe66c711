"""Minimal ONNX writer (protobuf wire format by hand; the ``onnx`` package is not available here).

Fixture tooling only: ``quantize.py`` uses it to write QOperator-format (QLinearConv ...) models of the same shape
as the file the reference's tests download (``fcn-resnet50-12-int8.onnx``, infur-test-gen/build.rs:89-91).  The product
reads such files back through its own C++ reader (csrc/onnx_reader.cpp); the oracle through ``oracle/onnx_min.py``.

Field numbers are those of onnx.proto3 (ModelProto, GraphProto, NodeProto, AttributeProto, TensorProto,
ValueInfoProto, TypeProto, TensorShapeProto).
"""
from __future__ import annotations

import struct

import numpy as np

FLOAT, UINT8, INT8, INT32, INT64 = 1, 2, 3, 6, 7
_NP2ONNX = {np.dtype(np.float32): FLOAT, np.dtype(np.uint8): UINT8, np.dtype(np.int8): INT8, np.dtype(np.int32): INT32,
            np.dtype(np.int64): INT64}


def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _int(field: int, v: int) -> bytes:
    return _varint(field << 3) + _varint(v)


def _bytes(field: int, b: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(b)) + b


def _str(field: int, s: str) -> bytes:
    return _bytes(field, s.encode())


def tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.asarray(arr, order="C")   # (ascontiguousarray would turn a 0-d scalar into shape [1])
    out = b"".join(_int(1, int(d)) for d in arr.shape)
    out += _int(2, _NP2ONNX[arr.dtype]) + _str(8, name) + _bytes(9, arr.tobytes())
    return out


def attribute(name: str, v) -> bytes:
    out = _str(1, name)
    if isinstance(v, float):
        return out + _varint((2 << 3) | 5) + struct.pack("<f", v) + _int(20, 1)
    if isinstance(v, int):
        return out + _int(3, v) + _int(20, 2)
    if isinstance(v, str):
        return out + _bytes(4, v.encode()) + _int(20, 3)
    if isinstance(v, np.ndarray):
        return out + _bytes(5, tensor("", v)) + _int(20, 4)
    if isinstance(v, (list, tuple)):
        return out + b"".join(_int(8, int(x)) for x in v) + _int(20, 7)
    raise TypeError(type(v))


def node(op: str, inputs, outputs, name: str = "", domain: str = "", **attrs) -> bytes:
    out = b"".join(_str(1, i) for i in inputs) + b"".join(_str(2, o) for o in outputs)
    out += _str(3, name or outputs[0]) + _str(4, op)
    out += b"".join(_bytes(5, attribute(k, v)) for k, v in attrs.items())
    if domain:
        out += _str(7, domain)
    return out


def value_info(name: str, elem_type: int, dims) -> bytes:
    shape = b""
    for d in dims:
        shape += _bytes(1, _str(2, d) if isinstance(d, str) else _int(1, int(d)))
    ttype = _int(1, elem_type) + _bytes(2, shape)
    return _str(1, name) + _bytes(2, _bytes(1, ttype))


def model(nodes, initializers, inputs, outputs, opsets=(("", 12),), producer: str = "infur_b200.onnx_write", graph_name: str = "g") -> bytes:
    """nodes / inputs / outputs: lists of encoded messages (``node`` / ``value_info``); initializers: {name: ndarray}."""
    g = b"".join(_bytes(1, n) for n in nodes) + _str(2, graph_name)
    g += b"".join(_bytes(5, tensor(k, v)) for k, v in initializers.items())
    g += b"".join(_bytes(11, i) for i in inputs) + b"".join(_bytes(12, o) for o in outputs)
    m = _int(1, 7) + _str(2, producer) + _bytes(7, g)
    for domain, version in opsets:
        m += _bytes(8, (_str(1, domain) if domain else b"") + _int(2, version))
    return m
