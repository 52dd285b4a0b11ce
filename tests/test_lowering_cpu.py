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


def test_serving_lowering_fuzz_against_per_event_steps():
    """hypothesis: random chains of Imputer / MapValues / OneHotEncoder / DropFeatures over feature-dict events --
    whatever `ColumnProgram` accepts must give, row by row, the values and field order of the per-event steps"""
    from hypothesis import HealthCheck, assume, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=200, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.data())
    def run(data):
        draw = data.draw
        rng = np.random.default_rng(draw(st.integers(0, 10_000)))
        n = draw(st.integers(1, 10))
        num = [f"x{i}" for i in range(draw(st.integers(1, 3)))]
        cat = [f"c{i}" for i in range(draw(st.integers(1, 2)))]
        names = num + cat
        X = np.empty((n, len(names)), dtype=np.float32)
        for j, c in enumerate(names):
            if c in num:
                col = (rng.integers(-6, 7, size=n) / 4).astype(np.float32) + np.float32(draw(st.sampled_from([0.0, 0.1, 0.7])))
                col[rng.random(n) < 0.25] = np.nan
            else:
                col = rng.integers(0, 5, size=n).astype(np.float32)
            X[:, j] = col
        live, steps = list(names), []
        for kind in draw(st.lists(st.sampled_from(["imp", "map", "onehot", "drop"]), min_size=1, max_size=4)):
            if kind == "imp":
                # every numeric column needs a fill: without one the reference turns NaN into None, which its own later
                # steps cannot compare; categorical codes are never missing here
                steps.append(("Imputer", dict(mapping={c: draw(st.sampled_from([0.5, 0.1, -1.0])) for c in num if c in live}, default_value=0)))
            elif kind == "map":
                cands = [c for c in live if c in names]
                if not cands:
                    continue
                col = draw(st.sampled_from(cands))
                if col in num:
                    fmap = {"ranges": {0: ["-inf", 0.1], 1: [0.1, 0.7], 2: [0.7, "inf"]}}
                else:
                    fmap = {0: 3, 1: 4, 9: 1}
                steps.append(("MapValues", dict(mapping={col: fmap})))
                live = [col]
            elif kind == "onehot":
                cands = [c for c in live if c in cat]
                if not cands:
                    continue
                col = draw(st.sampled_from(cands))
                steps.append(("OneHotEncoder", dict(mapping={col: [0, 1, 2]})))
                live = [x for c in live for x in ([f"{col}_{k}" for k in (0, 1, 2)] if c == col else [c])]
            elif kind == "drop" and len(live) > 1:
                col = draw(st.sampled_from(live))
                steps.append(("DropFeatures", dict(features=[col])))
                live = [c for c in live if c != col]
        assume(steps)
        try:
            prog = ColumnProgram(names)
            for cls, kw in steps:
                prog.apply(getattr(api_oracle, cls)(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()}))
        except LoweringError:
            assume(False)
        got = emu.transform(prog, X)
        objs = [getattr(api_oracle, cls)(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()}) for cls, kw in steps]
        for r in range(n):
            ev = {c: (int(X[r, j]) if c in cat else float(X[r, j])) for j, c in enumerate(names)}
            try:
                for o in objs:
                    ev = o._do_storey(ev)
            except Exception:  # noqa: BLE001 -- the reference itself fails on this event
                assume(False)
            assert list(ev.keys()) == prog.out_names
            want = np.array([np.nan if v is None else float(v) for v in ev.values()], dtype=np.float64)
            np.testing.assert_allclose(got[r].astype(np.float64), want, rtol=1e-7, atol=0, equal_nan=True,
                                       err_msg=f"row {r} steps {[s[0] for s in steps]}")

    run()


def test_imputed_values_keep_their_range_and_key():
    """an Imputer fill equal to a non-float32 range bound (0.7) or map key (0.1): the reference maps the exact value"""
    X = np.array([[np.nan, np.nan], [0.25, 0.5]], dtype=np.float32)
    steps = lambda: [api_oracle.Imputer(mapping={"a": 0.7, "b": 0.1}),  # noqa: E731
                     api_oracle.MapValues(mapping={"a": {"ranges": {0: ["-inf", 0.7], 1: [0.7, "inf"]}}, "b": {0.1: 5, 0.5: 6}})]
    prog = ColumnProgram(["a", "b"])
    for s in steps():
        prog.apply(s)
    got = emu.transform(prog, X)
    for r in range(2):
        ev = {"a": float(X[r, 0]), "b": float(X[r, 1])}
        for s in steps():
            ev = s._do_storey(ev)
        np.testing.assert_array_equal(got[r], np.array(list(ev.values()), dtype=np.float32))


def test_event_packing_and_responses_of_the_batched_path():
    """CompiledGraph.pack_events / responses (host code around the one fused launch of run_events)"""
    import types

    from mlrun_b200.serving.compiler import CompiledGraph

    names = [f"f{i}" for i in range(7)]
    cg = CompiledGraph(None, None, names, ("ens", "v1"))
    rng = np.random.default_rng(8)
    bodies = []
    for i in range(50):
        vals = [float(v) for v in rng.normal(size=7) * 10.0 ** rng.integers(-20, 20, size=7)]
        vals[i % 7] = [None, 3, True, np.float32(0.1), np.float64(1e-50), float("nan"), 16777217][i % 7]
        bodies.append(dict(zip(names, vals)))
    want = np.empty((50, 7), dtype=np.float32)
    for i, b in enumerate(bodies):  # the plain formulation: one row at a time, None -> NaN
        want[i] = [np.nan if v is None else v for v in b.values()]
    got = cg.pack_events(bodies)
    assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"]
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    np.testing.assert_array_equal(cg.pack_events(bodies[1:2]).view(np.uint32), want[1:2].view(np.uint32))  # no None: the fast pass
    with pytest.raises(ValueError, match="compiled schema"):
        cg.pack_events([dict(reversed(list(bodies[0].items())))])
    with pytest.raises(ValueError):
        cg.pack_events([{**bodies[1], "f3": "text"}])
    assert cg.pack_events([]).shape == (0, 7)

    ctx = types.SimpleNamespace(Response=lambda **kw: types.SimpleNamespace(**kw))
    out = np.array([[1.5], [2.5], [np.nan]], dtype=np.float32)
    res = cg.responses(out, np.array([0, 0, 1], dtype=np.int32), ctx)
    assert res[:2] == [{"model_name": "ens", "outputs": [1.5], "model_version": "v1"}, {"model_name": "ens", "outputs": [2.5], "model_version": "v1"}]
    assert res[2].status_code == 400 and res[2].body.startswith("ValueError")
    labels = CompiledGraph(None, None, names, ("clf", None)).responses(np.array([[2], [0]], dtype=np.int32), np.zeros(2, dtype=np.int32), ctx)
    assert labels == [{"model_name": "clf", "outputs": [2]}, {"model_name": "clf", "outputs": [0]}] and type(labels[0]["outputs"][0]) is int


def test_only_plain_linear_estimators_are_exported_as_linear():
    """An estimator that merely carries coef_ is not X @ coef_.T + b: GLMs apply exp, SVC(kernel='linear') keeps
    one-vs-one rows, PLS centres.  They must be refused (UnsupportedModel), never exported with the wrong arithmetic."""
    from sklearn.cross_decomposition import PLSRegression
    from sklearn.linear_model import LogisticRegression, PoissonRegressor, Ridge, SGDClassifier
    from sklearn.svm import SVC, LinearSVC

    rng = np.random.default_rng(5)
    X = rng.normal(size=(200, 6))
    y_cnt = rng.poisson(np.exp(0.3 * X[:, 0]) + 0.5)
    y3 = (X[:, 0] > 0.4).astype(int) + (X[:, 1] > 0.2).astype(int)
    for bad in (PoissonRegressor().fit(X, y_cnt), SVC(kernel="linear").fit(X, y3), PLSRegression(n_components=2).fit(X, X[:, 0])):
        with pytest.raises(packing.UnsupportedModel):
            packing.pack_model(bad)
    for good in (Ridge().fit(X, X[:, 0] * 2), LogisticRegression().fit(X, y3), LinearSVC().fit(X, y3),
                 SGDClassifier(random_state=0).fit(X, y3 > 0)):
        kind, packed = packing.pack_model(good)
        assert kind == "linear" and packed["n_features"] == 6
        scores = X @ packed["W"].T + packed["b"]
        if packed["link"] == packing.nat.LINK_IDENTITY:
            got = scores[:, 0]
        elif packed["link"] == packing.nat.LINK_BINARY_GT:
            got = packed["classes"][(scores[:, 0] > 0).astype(int)]
        else:
            got = packed["classes"][scores.argmax(axis=1)]
        np.testing.assert_allclose(got, good.predict(X), rtol=1e-9)


def test_model_server_refuses_a_request_of_the_wrong_width():
    """sklearn's predict raises for a width mismatch (extra columns included); so does the device server, before any
    device call (this runs without a GPU)."""
    from sklearn.linear_model import Ridge
    from sklearn.tree import DecisionTreeRegressor

    from mlrun_b200.serving.device_models import PickleModelServer

    rng = np.random.default_rng(6)
    X = rng.normal(size=(64, 5))
    for model in (Ridge().fit(X, X[:, 0]), DecisionTreeRegressor(max_depth=2).fit(X, X[:, 0])):
        server = PickleModelServer(name="m", model=model)
        with pytest.raises(ValueError, match="expecting 5 features"):
            server.predict({"inputs": rng.normal(size=(3, 7)).tolist()})
        with pytest.raises(ValueError):
            model.predict(rng.normal(size=(3, 7)))
