"""Oracle for the feature-set ingest path (test infrastructure; never imported by the product).

`ingest_rows` restates what the reference does to an in-memory frame with the storey engine
(mlrun/feature_store/ingestion.py:38-127): the frame is emitted one dict per row (storey.DataframeSource,
datastore/sources.py:886-895), every row walks the steps' `_do_storey` (oracle/transforms.py, pinned to the real
reference by tests/golden) and the rows are re-assembled into a frame (ReduceToDataFrame, datastore/targets.py:1856-1868).
`ingest_columns` computes the same frame column-wise with numpy/pandas vector ops -- the checker at sizes where
the per-row walk takes minutes, and the "vectorised" CPU baseline; tests/test_ingest_cpu.py ties the two together.
"""

import types

import numpy as np
import pandas as pd


def ingest_rows(steps, df):
    """-> (frame, violations) ; steps are oracle.transforms objects"""
    out = []
    for row in df.to_dict("records"):
        body = row
        for step in steps:
            if type(step).__name__ == "FeaturesetValidator":
                step._do_storey(types.SimpleNamespace(body=body, key=None))
            else:
                body = step._do_storey(body)
        out.append(body)
    violations = sum(getattr(s, "violations", 0) for s in steps)
    return pd.DataFrame(out, index=df.index), violations


def _range_map(x, ranges):
    out = x.astype(np.float64).copy()
    done = np.zeros(len(x), dtype=bool)
    labels_int = True
    for val, (lo, hi) in ranges.items():
        lo = -np.inf if lo == "-inf" else lo
        hi = np.inf if hi == "inf" else hi
        hit = (~done) & (x >= lo) & (x < hi)
        out[hit] = val
        done |= hit
        labels_int &= isinstance(val, (int, np.integer)) and not isinstance(val, bool)
    return out, done, labels_int


def ingest_columns(steps, df):
    """vectorised restatement; -> (frame, {column: violations})"""
    cols = {str(c): df[c].to_numpy() for c in df.columns}
    violations = {}
    for step in steps:
        kind = type(step).__name__
        if kind == "Imputer":  # steps.py:397-406
            for name, a in list(cols.items()):
                fill = step.mapping.get(name, step.default_value)
                if fill is None or a.dtype.kind != "f":
                    continue
                a = a.copy()
                a[np.isnan(a)] = fill
                cols[name] = a
        elif kind == "MapValues":  # steps.py:189-216
            mapped = {}
            for name, a in cols.items():
                if name not in step.mapping:
                    continue
                fmap = step.mapping[name]
                if "ranges" in fmap:
                    out, done, is_int = _range_map(a, fmap["ranges"])
                else:
                    out, done = a.astype(np.float64).copy(), np.zeros(len(a), dtype=bool)
                    is_int = all(isinstance(v, (int, np.integer)) and not isinstance(v, bool) for v in fmap.values())
                    for k, v in fmap.items():
                        hit = (~done) & (a == k)
                        out[hit] = v
                        done |= hit
                if is_int and done.all():
                    out = out.astype(np.int64)
                mapped[step._get_feature_name(name)] = out
            if step.with_original_features:
                mapped.update(cols)
            cols = mapped
        elif kind == "OneHotEncoder":  # steps.py:453-478
            new = {}
            for name, a in cols.items():
                cats = step.mapping.get(name)
                if not cats:
                    new[name] = a
                    continue
                for c in dict.fromkeys(cats):
                    new[f"{name}_{step._sanitized_category(c)}"] = (a == c).astype(np.int64)
            cols = new
        elif kind == "DateExtractor":  # steps.py:593-602
            ts = pd.Series(cols[step.timestamp_col])
            for part in step.parts:
                if part in ("week", "weekofyear"):  # Series.dt lost these accessors; Timestamp.week is the ISO week
                    v = ts.dt.isocalendar().week.to_numpy().astype(np.int64)
                else:
                    v = getattr(ts.dt, part).to_numpy()
                cols[f"{step.timestamp_col}_{part}"] = v
        elif kind == "DropFeatures":  # steps.py:721-729
            for f in step.features:
                cols.pop(f)
        elif kind == "FeaturesetValidator":  # steps.py:117-128 + mlrun/features.py:292-321
            for name, v in step._validators.items():
                if name not in cols:
                    continue
                a = cols[name]
                bad = np.zeros(len(a), dtype=bool)
                if v.min is not None:
                    bad |= a < v.min
                if v.max is not None:
                    bad |= (~bad) & (a > v.max)
                violations[name] = violations.get(name, 0) + int(bad.sum())
        else:
            raise NotImplementedError(kind)
    return pd.DataFrame(cols, index=df.index), violations
