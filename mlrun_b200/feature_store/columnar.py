"""Columnar sources and results of feature-set ingest: the DataFrame-free boundary.

The reference's ingest walks a DataFrame one dict per row (`storey.DataframeSource`, mlrun/datastore/sources.py:886-895)
and re-assembles a DataFrame (`ReduceToDataFrame`, mlrun/datastore/targets.py:1856-1868).  The device plan is columnar on
both sides, so the cheapest boundary is columnar too: a mapping column -> contiguous array in, a `ColumnBatch` out, both
convertible to / from Arrow without copies.  At ~2.3 KB per row the path is PCIe bound; what decides its speed is whether the
buffers are pinned (`pinned_columns` / `ColumnBatch` are) and whether the frame is long enough to be pipelined.
"""

import numpy as np

from .. import _native as nat
from .ingest import F32, I32, I64, _INT_DTYPES, LoweringError


def is_columnar(source):
    """dict of arrays, pyarrow Table / RecordBatch, or a ColumnBatch (a DataFrame is not: it takes the frame path)"""
    if isinstance(source, (dict, ColumnBatch)):
        return True
    mod = type(source).__module__.split(".")[0]
    return mod == "pyarrow" and hasattr(source, "column_names")


def _as_array(value, name):
    if isinstance(value, np.ndarray):
        return value
    if type(value).__module__.split(".")[0] == "pyarrow":  # ChunkedArray / Array: zero-copy when it has no nulls
        if hasattr(value, "combine_chunks") and getattr(value, "num_chunks", 1) != 1:
            value = value.combine_chunks()
        elif hasattr(value, "chunk"):
            value = value.chunk(0)
        if value.null_count:
            raise ValueError(f"column {name!r} has Arrow nulls: give float columns NaN (the Imputer's input) and fill the others")
        return value.to_numpy(zero_copy_only=True)
    if hasattr(value, "__dlpack__"):
        return np.from_dlpack(value)
    return np.asarray(value)


def as_columns(source):
    """-> {name: 1-D numpy array}, without copying where the producer allows it"""
    if isinstance(source, ColumnBatch):
        return dict(source.columns)
    if isinstance(source, dict):
        return {str(k): _as_array(v, k) for k, v in source.items()}
    if hasattr(source, "column_names"):  # pyarrow.Table / RecordBatch
        return {str(n): _as_array(source.column(n), n) for n in source.column_names}
    raise TypeError(f"{type(source).__name__} is not a columnar source")


def schema_of(columns):
    """[(name, kind)] like ingest.frame_schema, from arrays"""
    schema = []
    for name, a in columns.items():
        s = str(a.dtype)
        if s == "float32":
            kind = F32
        elif s in _INT_DTYPES:
            kind = I32
        elif s.startswith("datetime64") or s == "int64":
            if s == "int64":
                raise LoweringError(f"column {name!r} is int64: only timestamps (datetime64[ns]) are 8-byte columns; cast counters "
                                    "to int32 explicitly (a silent down-cast would change values)")
            kind = I64
        else:
            raise LoweringError(f"column {name!r} has dtype {s}: the device takes float32, (u)int8/16/32, bool and datetime64 columns")
        schema.append((str(name), kind))
    return schema


def pinned_columns(schema, n_rows):
    """{name: pinned (cudaMallocHost) array}: fill these instead of pageable arrays and the H2D copies run at PCIe speed
    (and can overlap the kernels).  `schema`: [(name, dtype)] or a mapping name -> dtype / array."""
    items = list(schema.items() if isinstance(schema, dict) else schema)
    dts = [(str(name), dt.dtype if hasattr(dt, "dtype") else np.dtype(dt)) for name, dt in items]
    # ONE pinned block, columns 64-byte aligned one after the other: neighbours of the same width sit at a constant pitch, which
    # lets b2s_cols_run_host move a row range of all of them with a single 2-D copy (the arrays keep the block alive)
    n_rows = int(n_rows)
    offs, off = [], 0
    for _name, dt in dts:
        offs.append(off)
        off += (n_rows * dt.itemsize + 63) // 64 * 64
    block = nat.pinned_empty((max(off, 64),), np.uint8)
    return {name: block[o: o + n_rows * dt.itemsize].view(dt) for (name, dt), o in zip(dts, offs)}


class ColumnBatch:
    """result of a columnar ingest: ordered {name: array}; the arrays are views of one pinned block that lives as long as
    any of them does.  `index` carries the entity columns of the source, untouched."""

    def __init__(self, columns, n_rows, block=None):
        self.columns = dict(columns)
        self.n_rows = int(n_rows)
        self.index = {}
        self._block = block

    def __getitem__(self, name):
        return self.columns[name]

    def __len__(self):
        return self.n_rows

    @property
    def names(self):
        return list(self.columns)

    def to_arrow(self):
        """pyarrow.Table over the same memory (numeric columns without nulls convert without a copy)"""
        import pyarrow as pa

        cols = {**self.index, **self.columns}
        return pa.table({k: pa.array(v) for k, v in cols.items()})

    def to_pandas(self):
        import pandas as pd

        frame = pd.DataFrame(self.columns, copy=False)
        if self.index:
            frame.index = pd.MultiIndex.from_arrays(list(self.index.values()), names=list(self.index)) if len(self.index) > 1 \
                else pd.Index(next(iter(self.index.values())), name=next(iter(self.index)))
        return frame
