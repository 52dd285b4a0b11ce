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
