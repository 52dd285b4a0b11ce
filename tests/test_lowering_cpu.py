"""Lowering + packing validated on CPU against the oracle / scikit-learn through a numpy emulation of
the plan arithmetic (no GPU; the CUDA kernels themselves are checked by the -m gpu tests)."""

import numpy as np
import pytest

from mlrun_b200 import packing
from mlrun_b200.lowering import ColumnProgram, LoweringError
from mlrun_b200.synthetic import flow3_workload, tree_workload
from oracle import batch as obatch
from tests import api_oracle, device_emulator as emu


def _flow3_program(wl):
    prog = ColumnProgram(wl.names)
    prog.apply(api_oracle.Imputer(mapping=dict(wl.impute_mapping), default_value=wl.impute_default))
    prog.apply(api_oracle.OneHotEncoder(mapping={k: list(v) for k, v in wl.onehot_mapping.items()}))
    return prog


@pytest.mark.parametrize("n_models", [1, 4])
def test_flow3_lowering_matches_oracle(n_models):
    wl = flow3_workload(n_rows=512, n_num=20, n_cat=4, seed=11, n_models=n_models)
    ref = obatch.flow3(wl)
    prog = _flow3_program(wl)
    assert prog.out_names == ref["names"] == wl.out_names
    E = emu.transform(prog, wl.X)
    np.testing.assert_array_equal(E.astype(np.float64), ref["expanded"])  # f32 values, exact
    models = [packing.pack_model(m) for m in wl.sklearn_models()]
    per_model = emu.predict(models, E)
    np.testing.assert_allclose(per_model, ref["per_model"], rtol=1e-12, atol=1e-12)


def test_batch_oracle_matches_per_event_oracle():
    """the vectorised restatement agrees with the per-event restatement (the one pinned to the reference)"""
    for n_models in (1, 4):
        wl = flow3_workload(n_rows=64, n_num=12, n_cat=4, seed=5, n_models=n_models)
        server = wl.build_server(api_oracle)
        path = "/" if n_models == 1 else "/v2/models/infer"
        per_event = [server.test(path=path, body=row)["outputs"][0] for row in wl.rows_as_dicts()]
        np.testing.assert_allclose(per_event, obatch.flow3(wl)["out"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("kind", ["regression", "classification"])
def test_gbt_packing_matches_sklearn(kind):
    wl = tree_workload(n_rows=300, n_feat=12, n_models=2, n_trees=7, depth=4, seed=9, kind=kind, n_fit=400)
    ref = obatch.tree_ensemble(wl)
    models = [packing.pack_model(m) for m in wl.models]
    per_model = emu.predict(models, wl.X)
    if kind == "regression":
        np.testing.assert_allclose(per_model, ref["per_model"], rtol=1e-12, atol=1e-12)
    else:
        np.testing.assert_array_equal(per_model, ref["per_model"])


def test_binary_gbt_and_forests_and_single_trees():
    from sklearn.ensemble import GradientBoostingClassifier, RandomForestClassifier, RandomForestRegressor
    from sklearn.linear_model import LogisticRegression, Ridge
    from sklearn.tree import DecisionTreeClassifier, DecisionTreeRegressor

    rng = np.random.default_rng(3)
    X = rng.normal(size=(500, 8)).astype(np.float32)
    y = X[:, 0] * 2 + np.sin(X[:, 1]) + X[:, 2] * X[:, 3]
    yb = (y > 0).astype(int)
    yc = np.digitize(y, [-1, 1])
    Xt = rng.normal(size=(400, 8)).astype(np.float32)
    cases = [
        (GradientBoostingClassifier(n_estimators=5, max_depth=3, random_state=0).fit(X, yb), True),
        (RandomForestRegressor(n_estimators=6, max_depth=5, random_state=0).fit(X, y), False),
        (RandomForestClassifier(n_estimators=6, max_depth=5, random_state=0).fit(X, yc), True),
        (DecisionTreeRegressor(max_depth=6, random_state=0).fit(X, y), False),
        (DecisionTreeClassifier(max_depth=6, random_state=0).fit(X, yc), True),
        (LogisticRegression().fit(X, yb), True),
        (LogisticRegression().fit(X, yc), True),
        (Ridge().fit(X, y), False),
    ]
    for model, exact in cases:
        got = emu.predict([packing.pack_model(model)], Xt)[:, 0]
        want = model.predict(Xt.astype(np.float64))
        if exact:
            np.testing.assert_array_equal(got, want, err_msg=type(model).__name__)
        else:
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, err_msg=type(model).__name__)


def test_threshold_rounding_is_decision_preserving():
    rng = np.random.default_rng(0)
    thr = rng.normal(size=10000)
    t32 = packing.threshold_to_f32(thr)
    assert (t32.astype(np.float64) <= thr).all()
    assert (np.nextafter(t32, np.float32(np.inf)).astype(np.float64) > thr).all()
    x = rng.normal(size=10000).astype(np.float32)
    np.testing.assert_array_equal(x <= t32, x.astype(np.float64) <= thr)


def test_map_values_and_drop_lowering():
    names = ["a", "b", "c", "d"]
    rng = np.random.default_rng(1)
    X = rng.normal(size=(200, 4)).astype(np.float32) * 10
    X[:, 1] = rng.integers(0, 4, size=200)
    X[::17, 0] = np.nan
    mapping = {"a": {"ranges": {0: ["-inf", 0], 1: [0, 5], 2: [5, "inf"]}}, "b": {0: 10, 1: 11}}
    prog = ColumnProgram(names)
    prog.apply(api_oracle.Imputer(default_value=0.5))
    prog.apply(api_oracle.MapValues(mapping=mapping))
    got = emu.transform(prog, X)
    Xi = obatch.impute(X, names, None, 0.5)
    want, out_names = obatch.map_values(Xi, names, mapping)
    assert prog.out_names == out_names == ["a", "b"]
    np.testing.assert_array_equal(got.astype(np.float64), want)
    # per-event restatement agrees too
    imp, mv = api_oracle.Imputer(default_value=0.5), api_oracle.MapValues(mapping=mapping)
    for r in range(0, 200, 13):
        ev = {n: (float(v) if n != "b" else int(v)) for n, v in zip(names, X[r])}
        res = mv.do(imp.do(ev))
        assert list(res.keys()) == ["a", "b"]
        np.testing.assert_array_equal([float(res["a"]), float(res["b"])], want[r])

    prog = ColumnProgram(names)
    prog.apply(api_oracle.DropFeatures(features=["b", "d"]))
    assert prog.out_names == ["a", "c"]
    with pytest.raises(LoweringError):
        ColumnProgram(names).apply(api_oracle.DropFeatures(features=["zz"]))
    with pytest.raises(LoweringError):
        ColumnProgram(names).apply(api_oracle.OneHotEncoder(mapping={"a": ["x", "y"]}))
    with pytest.raises(LoweringError):
        ColumnProgram(names).apply(api_oracle.MapValues(mapping={"a": {1: "one"}}))


def test_range_bounds_that_are_not_float32_keep_the_reference_decisions():
    """MapValues bounds like 0.7 or 0.1 are float64 in the reference; events carry float32 values.  The lowered float32
    constants are rounded so that every float32 value falls in the same range as in the reference's float64 compare"""
    edges = [0.1, 0.7, 0.3, 1.0 / 3.0]
    vals = []
    for e in edges:
        f = np.float32(e)
        vals += [np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))]
    X = np.array(vals + [0.0, 0.5, 2.0], dtype=np.float32).reshape(-1, 1)
    mapping = {"a": {"ranges": {0: ["-inf", 0.1], 1: [0.1, 0.3], 2: [0.3, 1.0 / 3.0], 3: [1.0 / 3.0, 0.7], 4: [0.7, "inf"]}}}
    prog = ColumnProgram(["a"])
    prog.apply(api_oracle.MapValues(mapping=mapping))
    got = emu.transform(prog, X)[:, 0]
    step = api_oracle.MapValues(mapping=mapping)
    want = [step._do_storey({"a": float(v)})["a"] for v in X[:, 0]]
    np.testing.assert_array_equal(got, np.array(want, dtype=np.float32))
    # value maps: a key that is not a float32 never matches
    prog = ColumnProgram(["a"])
    prog.apply(api_oracle.MapValues(mapping={"a": {0.1: 5, 0.5: 6}}))
    got = emu.transform(prog, np.array([[0.1], [0.5]], dtype=np.float32))[:, 0]
    step = api_oracle.MapValues(mapping={"a": {0.1: 5, 0.5: 6}})
    want = [step._do_storey({"a": float(np.float32(v))})["a"] for v in (0.1, 0.5)]
    np.testing.assert_array_equal(got, np.array(want, dtype=np.float32))
