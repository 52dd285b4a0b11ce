"""Lower a run of feature-store steps (+ models + vote) into a DevicePlan.

Semantics lowered are the *storey-engine* (per-event dict) semantics of the reference steps
(mlrun/feature_store/steps.py `_do_storey`), applied to rows whose columns are named by `in_names`
in dict-insertion order:

  Imputer        :397-406   every column: NaN -> mapping.get(name, default_value)
  MapValues      :189-216   only mapped columns survive (with_original_features=False); value maps
                            pass unmapped values through, range maps take the first [lo, hi) hit
  OneHotEncoder  :453-478   a mapped column is replaced in place by one 0/1 column per category
  DropFeatures   :721-729   listed columns removed (missing column -> error)

Anything that cannot be expressed on numeric float32 columns (string values, ops on derived one-hot
columns, Imputer after MapValues on the same column ...) raises `LoweringError`: on the GPU path an
unrecognised step is a hard error, never a silent CPU fallback.
"""

import math
import re

import numpy as np

from . import _native as nat
from .plan import DevicePlan


class LoweringError(ValueError):
    pass


def _sanitized(category):
    # OneHotEncoder._sanitized_category (steps.py:508-513)
    if isinstance(category, str):
        return re.sub("[ -]", "_", category)
    return category


def _ceil_f32(v):
    """smallest float32 >= v.  The device compares float32 values with float32 constants; for a float32 x,
    `x >= lo` and `x < hi` against the reference's float64 bounds are exactly `x >= ceil32(lo)` and `x < ceil32(hi)`."""
    if not math.isfinite(v):
        return v
    f = np.float32(v)
    if float(f) < v:
        f = np.nextafter(f, np.float32(np.inf))
    return float(f)


def _fill_in_same_range(fill, exact_ranges, name):
    """the float32 the device imputes must fall in the range the reference's exact (float64) fill falls in"""
    def label(x):
        return next((i for i, (lo, hi) in enumerate(exact_ranges) if lo <= x < hi), None)

    want = label(fill)
    near = np.float32(fill)
    for cand in (near, np.nextafter(near, np.float32(np.inf)), np.nextafter(near, np.float32(-np.inf))):
        if label(float(cand)) == want:
            return float(cand)
    raise LoweringError(f"Imputer fill {fill!r} of {name!r}: no float32 next to it falls in the same MapValues range")


def _num(v, what):
    if isinstance(v, bool) or not isinstance(v, (int, float, np.integer, np.floating)):
        raise LoweringError(f"{what}: {v!r} is not numeric; the device path handles numeric columns only")
    f = float(v)
    if not math.isfinite(f) and not math.isinf(f):
        raise LoweringError(f"{what}: NaN is not a valid constant")
    if math.isfinite(f) and float(np.float32(f)) != f and isinstance(v, (int, np.integer)):
        raise LoweringError(f"{what}: integer {v} is not exactly representable in float32")
    return f


class ColumnProgram:
    """symbolic execution of the steps over the column list"""

    def __init__(self, in_names):
        self.in_names = list(in_names)
        self.n_in = len(self.in_names)
        if len(set(self.in_names)) != self.n_in:
            raise LoweringError("duplicate input column names")
        # current columns: (name, src, kind, arg)
        self.cols = [(n, i, nat.OUT_COPY, 0.0) for i, n in enumerate(self.in_names)]
        self.fills = {}  # src -> fill
        self.maps = {}  # src -> list of ("value", {k: v}) | ("range", [(lo, hi, val)])
        self.steps = []

    # -- step handlers ---------------------------------------------------------------------------
    def imputer(self, mapping=None, default_value=None, **_):
        mapping = mapping or {}
        for name, src, kind, _arg in self.cols:
            fill = mapping.get(name, default_value)
            if fill is None:
                continue
            if kind != nat.OUT_COPY:
                # a 0/1 one-hot column is never NaN: imputing it is a no-op
                continue
            if src in self.maps:
                raise LoweringError(f"Imputer after MapValues on column {name!r} is not lowered")
            if src in self.fills:
                continue  # already imputed upstream: no NaN can be left
            self.fills[src] = _num(fill, f"Imputer fill for {name!r}")
        self.steps.append("Imputer")

    def map_values(self, mapping, with_original_features=False, suffix="mapped", **_):
        if with_original_features:
            raise LoweringError("MapValues(with_original_features=True) is not lowered")
        new_cols = []
        for name, src, kind, arg in self.cols:
            if name not in mapping:
                continue  # storey mode emits only the mapped features (steps.py:206-211)
            if kind != nat.OUT_COPY:
                raise LoweringError(f"MapValues on derived column {name!r} is not lowered")
            fmap = mapping[name]
            if "ranges" in fmap:
                ranges, exact = [], []
                for val, (lo, hi) in fmap["ranges"].items():
                    lo = -math.inf if lo == "-inf" else _num(lo, f"MapValues range of {name!r}")
                    hi = math.inf if hi == "inf" else _num(hi, f"MapValues range of {name!r}")
                    exact.append((lo, hi))
                    ranges.append((_ceil_f32(lo), _ceil_f32(hi), _num(val, f"MapValues range label of {name!r}")))
                if src in self.fills and len(self.maps.get(src, [])) == 0:
                    self.fills[src] = _fill_in_same_range(self.fills[src], exact, name)
                self.maps.setdefault(src, []).append(("range", ranges))
                others = {k: v for k, v in fmap.items() if k != "ranges"}
                if others:
                    raise LoweringError("MapValues mixing ranges and value replacements is rejected by the reference")
            else:
                vm = {_num(k, f"MapValues key of {name!r}"): _num(v, f"MapValues value of {name!r}") for k, v in fmap.items()}
                # a key that is not a float32 can never equal a float32 event value (the reference compares in float64)
                fill = self.fills.get(src) if len(self.maps.get(src, [])) == 0 else None
                if fill is not None and fill in vm and float(np.float32(fill)) != fill:
                    # the reference maps the exact (float64) fill through this key; the device fill is a float32 and would
                    # miss it: impute the mapped value directly (it must not be a key itself, the map runs once)
                    if vm[fill] in vm:
                        raise LoweringError(f"Imputer fill {fill!r} of {name!r} maps to a value that is itself a key")
                    self.fills[src] = vm[fill]
                vm = {k: v for k, v in vm.items() if not math.isfinite(k) or float(np.float32(k)) == k}
                self.maps.setdefault(src, []).append(("value", vm))
            new_cols.append((name, src, kind, arg))
        self.cols = new_cols
        self.steps.append("MapValues")

    def one_hot(self, mapping, **_):
        new_cols = []
        for name, src, kind, arg in self.cols:
            cats = mapping.get(name)
            if not cats:
                new_cols.append((name, src, kind, arg))
                continue
            if kind != nat.OUT_COPY:
                raise LoweringError(f"OneHotEncoder on derived column {name!r} is not lowered")
            seen = []
            for c in cats:  # de-dup preserving order (steps.py:444-451)
                if c not in seen:
                    seen.append(c)
            for c in seen:
                if isinstance(c, str):
                    raise LoweringError(f"OneHotEncoder category {c!r} of {name!r} is a string; encode it to an integer code first")
                new_cols.append((f"{name}_{_sanitized(c)}", src, nat.OUT_ONEHOT, _num(c, f"category of {name!r}")))
        self.cols = new_cols
        self.steps.append("OneHotEncoder")

    def drop(self, features, **_):
        names = [c[0] for c in self.cols]
        for f in features:
            if f not in names:
                raise LoweringError(f"The ingesting data doesn't contain a feature named '{f}'")
        self.cols = [c for c in self.cols if c[0] not in set(features)]
        self.steps.append("DropFeatures")

    def apply(self, step):
        """dispatch on the reference class name of a step object (duck-typed)"""
        kind = type(step).__name__
        if kind == "Imputer":
            self.imputer(step.mapping, step.default_value)
        elif kind == "OneHotEncoder":
            self.one_hot(step.mapping)
        elif kind == "MapValues":
            self.map_values(step.mapping, step.with_original_features, step.suffix)
        elif kind == "DropFeatures":
            self.drop(step.features)
        else:
            raise LoweringError(f"step class {kind} is not lowerable")
        return self

    # -- results ---------------------------------------------------------------------------------
    @property
    def out_names(self):
        return [c[0] for c in self.cols]

    @property
    def is_identity(self):
        return (not self.fills and not self.maps and len(self.cols) == self.n_in
                and all(c[1] == i and c[2] == nat.OUT_COPY for i, c in enumerate(self.cols)))

    def build_plan(self, models=(), vote=None):
        """models: list of ("linear", dict) | ("trees", PackedTrees); vote: None | (kind, weights)"""
        plan = DevicePlan(self.n_in)
        plan.set_impute(self.fills)
        for src, maps in self.maps.items():
            for kind, m in maps:
                if kind == "value":
                    plan.add_value_map(src, m)
                else:
                    plan.add_range_map(src, m)
        if not self.is_identity or self.fills or self.maps:
            if not self.cols:
                raise LoweringError("the steps leave no output columns")
            plan.set_output_schema([(src, kind, arg) for _n, src, kind, arg in self.cols])
        for kind, packed in models:
            if kind == "linear":
                plan.add_linear(packed["W"], packed["b"], packed["link"], packed["classes"])
            else:
                plan.add_trees(packed)
        if vote is not None:
            plan.set_vote(vote[0], vote[1])
        return plan.finalize()
