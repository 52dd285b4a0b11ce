"""Host logic of the product's online feature service against the golden outputs of the REAL reference
(`online_service_logic`), on CPU: the device table is replaced by a numpy stand-in with the gather kernel's semantics (the
kernel itself is compared with the oracle in tests/test_gpu_enrichment.py).  The served routers (`enrichment_routers`)
score on the device and are GPU tests.

Two representational differences are normalised, both inherent to holding the online rows as a float32 matrix:
a stored None / missing column is a NaN (so None == NaN below), and the label column is not an online feature (it is dropped
from the reference's dict answers)."""

import json
import math
import os

import numpy as np
import pandas as pd
import pytest

from mlrun_b200.feature_store import online as bo
from tests import api_b200, scenarios

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scenarios.json")))


class _HostTable:
    """b2s_table_* semantics in numpy: unknown key -> NaN row; then None / NaN / Inf -> the impute value where one is set"""

    def __init__(self, keys, values, impute=None):
        self.rows = {int(k): i for i, k in enumerate(np.asarray(keys, dtype=np.int64))}
        self.values = np.asarray(values, dtype=np.float32)
        self.n_keys, self.n_feat = self.values.shape
        self.impute = None if impute is None else np.asarray(impute, dtype=np.float32)

    def lookup(self, keys):
        idx = np.array([self.rows.get(int(k), -1) for k in keys], dtype=np.int64)
        out = np.where((idx >= 0)[:, None], self.values[np.maximum(idx, 0)], np.float32(np.nan)).astype(np.float32)
        if self.impute is not None:
            bad = ~np.isfinite(out) & ~np.isnan(self.impute)[None, :]
            out = np.where(bad, self.impute[None, :], out)
        return out, idx >= 0

    def close(self):
        pass


@pytest.fixture(autouse=True)
def _host_table(monkeypatch):
    monkeypatch.setattr(bo, "DeviceTable", _HostTable)


def _frame(features, index_keys, table):
    index = pd.MultiIndex.from_tuples(list(table), names=index_keys) if len(index_keys) > 1 else pd.Index([k[0] for k in table], name=index_keys[0])
    return pd.DataFrame([[row.get(f) for f in features] for row in table.values()], columns=features, index=index, dtype=np.float64)


def online_service(features, index_keys, table, stats, label_column, with_indexes, impute_policy):
    vec = bo.FeatureVector("vec", features, index_keys, _frame(features, index_keys, table), stats, label_column=label_column,
                           with_indexes=with_indexes)
    return vec.get_online_feature_service(impute_policy)


class _Api:
    """tests.api_b200 plus the adapter of this file"""

    def __getattr__(self, name):
        return getattr(api_b200, name)


API = _Api()
API.online_service = online_service


def _norm(v, label):
    if isinstance(v, dict) and "raised" not in v:
        return {k: _norm(x, label) for k, x in v.items() if k != label}
    if isinstance(v, list):
        return [_norm(x, label) for x in v]
    if v is None or v == "nan" or (isinstance(v, float) and math.isnan(v)):
        return "missing"
    return v


def _same(got, want, path=""):
    if isinstance(want, dict):
        assert isinstance(got, dict) and set(got) == set(want), f"{path}: {got!r} vs {want!r}"
        for k in want:
            _same(got[k], want[k], f"{path}/{k}")
    elif isinstance(want, list):
        assert isinstance(got, list) and len(got) == len(want), f"{path}: {got!r} vs {want!r}"
        for i, (g, w) in enumerate(zip(got, want)):
            _same(g, w, f"{path}[{i}]")
    else:
        assert got == want, f"{path}: {got!r} != {want!r}"


def test_online_service_host_logic_matches_the_real_reference():
    got = json.loads(json.dumps(scenarios.online_service_logic(API), default=str))
    want = GOLDEN["online_service_logic"]
    for tag in ("none", "mean", "mixed", "one", "zero"):
        for form in ("lists", "dicts", "one_dict", "extra_column", "impute_values"):
            g, w = _norm(got[tag][form], "y"), _norm(want[tag][form], "y")
            _same(g, w, f"{tag}/{form}")
    _same(_norm(got["with_indexes"], "y"), _norm(want["with_indexes"], "y"), "with_indexes")
    _same(_norm(got["no_label"], None), _norm(want["no_label"], None), "no_label")
    _same(_norm(got["composite"], "y"), _norm(want["composite"], "y"), "composite")
    _same(got["bad_input"], want["bad_input"], "bad_input")
    _same(got["bad_policy"], want["bad_policy"], "bad_policy")
