"""Body codec of the serving boundary: V2 JSON bodies <-> float32 rows, through the C library's host-side parser /
printer (include/b200serve.h "body codec").  It replaces, for bodies of the form {"inputs": [[...], ...], ...}, the
json.loads of GraphServer.run (mlrun/serving/server.py:262-277) and the json.dumps of _process_response (:298-308):
same values, same response text, without materialising B x F Python floats.

This is host logic on either side of the device path.  A body the parser does not take (string
or dict inputs, ragged rows) is reported as `NotV2Matrix`; the caller then uses the ordinary per-event JSON path.
"""

import ctypes as C
import json

import numpy as np

from .. import _native as nat


class NotV2Matrix(ValueError):
    """the body is valid for the per-event path but is not a numeric "inputs" matrix"""


def parse_inputs(body, out=None):
    """-> (X float32 (rows, cols), (begin, end) of the member's text).  `out`: optional float32 buffer to fill (e.g. pinned)"""
    if isinstance(body, str):
        body = body.encode()
    lib = nat.load()
    cap = len(body) // 2 + 1  # every number takes at least two bytes ("1,")
    if out is None:
        out = np.empty(cap, dtype=np.float32)
    flat = out.reshape(-1)
    rows, cols, b, e = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    rc = lib.b2s_json_parse_inputs(body, len(body), flat.ctypes.data_as(C.POINTER(C.c_float)), flat.size, C.byref(rows),
                                   C.byref(cols), C.byref(b), C.byref(e))
    if rc == -6:  # B2S_ERR_UNSUPPORTED
        raise NotV2Matrix(lib.b2s_last_error().decode())
    nat.check(rc)
    n = rows.value * cols.value
    return flat[:n].reshape(rows.value, cols.value), (b.value, e.value)


def decode_body(body, out=None):
    """-> (X, rest): rest is the body's other members as json.loads gives them"""
    if isinstance(body, str):
        body = body.encode()
    X, (b, e) = parse_inputs(body, out)
    rest = json.loads(body[:b] + b"null" + body[e:])
    rest.pop("inputs", None)
    return X, rest


def format_outputs(values, flat=None):
    """text of `values.tolist()` as json.dumps prints it; float32 or int32, 1-D (flat list) or 2-D"""
    a = np.ascontiguousarray(values)
    if a.dtype not in (np.float32, np.int32):
        raise TypeError("float32 or int32 results")
    if flat is None:
        flat = a.ndim == 1
    rows, cols = (a.shape[0], 1) if a.ndim == 1 else a.shape
    cap = 64 + rows * (cols * 32 + 4)  # the printer asks for 40 free bytes before each value
    buf = C.create_string_buffer(cap)
    n = C.c_int64()
    nat.check(nat.load().b2s_json_format_outputs(a.ctypes.data, 1 if a.dtype == np.int32 else 0, rows, cols, 1 if flat else 0,
                                                 buf, cap, C.byref(n)))
    return buf.raw[: n.value]


_MARK = " b2s-outputs "


def dumps_with_outputs(response, outputs_text):
    """json.dumps(response) with response["outputs"] replaced by already-formatted text"""
    body = dict(response)
    body["outputs"] = _MARK
    text = json.dumps(body).encode()
    return text.replace(json.dumps(_MARK).encode(), outputs_text, 1)


# ------------------------------------------------------------------------------------------ binary bodies
# Content type "application/x-b2s-f32": the rows themselves instead of their decimal text.  A V2 JSON body of 4096 x 128
# values is 10.8 MB of text for 2 MB of float32; parsing it is the first CPU cost once the kernels are fast.
#   request / response = 16-byte header + row-major 4-byte words, little endian:
#       magic b"B2S1" | uint32 rows | uint32 cols | uint32 flags        (flags bit 0: the words are int32 labels)
BINARY_CONTENT_TYPE = "application/x-b2s-f32"
_MAGIC = b"B2S1"


def encode_rows(values):
    """(rows, cols) float32 (or int32 labels) -> binary body"""
    a = np.ascontiguousarray(values)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2 or a.dtype not in (np.float32, np.int32):
        raise TypeError("a 2-D float32 (or int32) array")
    head = _MAGIC + np.array([a.shape[0], a.shape[1], 1 if a.dtype == np.int32 else 0], dtype="<u4").tobytes()
    return head + a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()


def decode_rows(body):
    """binary body -> (rows, cols) array viewing the body's bytes (float32, or int32 when the flag says so)"""
    mv = memoryview(body)
    if len(mv) < 16 or bytes(mv[:4]) != _MAGIC:
        raise ValueError("not an application/x-b2s-f32 body (bad magic)")
    rows, cols, flags = (int(v) for v in np.frombuffer(mv[4:16], dtype="<u4"))
    if len(mv) != 16 + rows * cols * 4:
        raise ValueError(f"application/x-b2s-f32 body of {rows} x {cols} words must be {16 + rows * cols * 4} bytes, got {len(mv)}")
    return np.frombuffer(mv[16:], dtype="<i4" if flags & 1 else "<f4").reshape(rows, cols)
