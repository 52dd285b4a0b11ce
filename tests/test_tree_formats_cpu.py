"""xgboost / LightGBM exporters (mlrun_b200/tree_formats.py) against the CPU restatement of the libraries' walks
(oracle/tree_libs.py), through the numpy emulation of the device arithmetic -- no GPU.  The `-m gpu` twin is
tests/test_gpu_trees.py."""

import json

import numpy as np
import pytest

from mlrun_b200 import _native as nat
from mlrun_b200 import packing, tree_formats
from mlrun_b200.lowering import ColumnProgram
from oracle import tree_libs
from tests import device_emulator as emu
from tests import emulated_plan, tree_fixtures as fx


@pytest.mark.parametrize("objective,num_class", [("reg:squarederror", 0), ("binary:logistic", 0), ("multi:softprob", 3)])
def test_xgboost_json_export_matches_the_published_walk(objective, num_class):
    doc = fx.random_xgb_model(n_trees=10, depth=5, n_feat=7, seed=11, objective=objective, num_class=num_class, base_score=0.3)
    X = fx.grid_inputs(400, 7, seed=12)
    packed = tree_formats.pack_xgboost_json(json.dumps(doc))  # through the text, as a file would come in
    assert packed.cmp_mode == nat.CMP_LT and packed.nan_ok and packed.n_features == 7
    got = emu.trees_predict(packed, X)
    want = tree_libs.xgboost_predict(doc, X)
    if objective == "reg:squarederror":
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    else:
        assert np.array_equal(got, want)


def test_less_than_is_not_less_or_equal():
    """the one place the libraries disagree: x == threshold goes right in xgboost, left in scikit-learn / LightGBM"""
    doc = fx.random_xgb_model(n_trees=1, depth=1, n_feat=1, seed=0, p_leaf=0.0)
    tree = doc["learner"]["gradient_booster"]["model"]["trees"][0]
    tree["split_conditions"] = [0.5, -1.0, 1.0]
    tree["default_left"] = [1, 0, 0]
    X = np.array([[0.5], [np.nextafter(np.float32(0.5), np.float32(0))], [np.nan], [-np.inf]], dtype=np.float32)
    packed = tree_formats.pack_xgboost_json(doc)
    got = emu.trees_predict(packed, X) - 0.5
    np.testing.assert_allclose(got, [1.0, -1.0, -1.0, -1.0])
    np.testing.assert_allclose(tree_libs.xgboost_predict(doc, X) - 0.5, [1.0, -1.0, -1.0, -1.0])
    tree["split_conditions"][0] = float("-inf")  # x < -inf never holds: everything but a default-left NaN goes right
    np.testing.assert_allclose(emu.trees_predict(tree_formats.pack_xgboost_json(doc), X) - 0.5, [1.0, 1.0, -1.0, 1.0])
    np.testing.assert_allclose(tree_libs.xgboost_predict(doc, X) - 0.5, [1.0, 1.0, -1.0, 1.0])


def test_xgboost_dump_format_gives_the_same_model():
    doc = fx.random_xgb_model(n_trees=6, depth=4, n_feat=5, seed=21)
    dump = [json.dumps(t) for t in fx.xgb_doc_to_dump(doc)]
    X = fx.grid_inputs(200, 5, seed=22)
    a = emu.trees_predict(tree_formats.pack_xgboost_json(doc), X)
    b = emu.trees_predict(tree_formats.pack_xgboost_dump(dump, base_score=0.5), X)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(b, tree_libs.xgboost_dump_predict(dump, X, base_score=0.5), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("objective,num_class", [("regression", 1), ("binary", 1), ("multiclass", 3)])
def test_lightgbm_dump_export_matches_the_published_walk(objective, num_class):
    dump = fx.random_lgbm_dump(n_trees=8, depth=5, n_feat=6, seed=31, objective=objective, num_class=num_class)
    X = fx.grid_inputs(400, 6, seed=32)
    packed = tree_formats.pack_lightgbm_dump(json.dumps(dump))
    assert packed.cmp_mode == nat.CMP_LE and packed.nan_ok
    got = emu.trees_predict(packed, X)
    want = tree_libs.lightgbm_predict(dump, X)
    if objective == "regression":
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-9)
    else:
        assert np.array_equal(got, want)


def test_unsupported_documents_are_refused():
    doc = fx.random_xgb_model(n_trees=2, depth=2, n_feat=3, seed=1)
    doc["learner"]["objective"]["name"] = "count:poisson"
    with pytest.raises(packing.UnsupportedModel):
        tree_formats.pack_xgboost_json(doc)
    dump = fx.random_lgbm_dump(n_trees=2, depth=2, n_feat=3, seed=1)
    dump["tree_info"][0]["tree_structure"]["decision_type"] = "=="
    with pytest.raises(packing.UnsupportedModel):
        tree_formats.pack_lightgbm_dump(dump)
    dump = fx.random_lgbm_dump(n_trees=2, depth=2, n_feat=3, seed=1)
    dump["tree_info"][0]["tree_structure"]["missing_type"] = "Zero"
    with pytest.raises(packing.UnsupportedModel):
        tree_formats.pack_lightgbm_dump(dump)
    with pytest.raises(packing.UnsupportedModel):
        packing.pack_model({"x": 1})


def test_model_server_over_an_xgboost_document(monkeypatch, tmp_path):
    """XGBoostModelServer(model_path="model.json"): NaN inputs are data, Inf still fails the request"""
    from mlrun_b200.serving.device_models import XGBoostModelServer

    emulated_plan.install(monkeypatch)
    doc = fx.random_xgb_model(n_trees=8, depth=4, n_feat=6, seed=41)
    path = tmp_path / "model.json"
    path.write_text(json.dumps(doc))
    server = XGBoostModelServer(name="xgb", model_path=str(path))
    server.load()
    X = fx.grid_inputs(64, 6, seed=42)
    got = server.predict({"inputs": X.astype(np.float64).tolist()})
    np.testing.assert_allclose(got, tree_libs.xgboost_predict(doc, X), rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError, match="expecting 6 features"):
        server.predict({"inputs": [[0.0] * 5]})
    bad = X[:3].copy()
    bad[1, 2] = np.inf
    with pytest.raises(ValueError, match="infinity"):
        server.predict({"inputs": bad.astype(np.float64).tolist()})


def test_sklearn_forests_route_nan_like_predict():
    """scikit-learn >= 1.4 forests / >= 1.3 single trees accept NaN at predict time and follow missing_go_to_left"""
    from sklearn.ensemble import GradientBoostingRegressor, RandomForestClassifier, RandomForestRegressor
    from sklearn.tree import DecisionTreeRegressor

    rng = np.random.default_rng(51)
    X = rng.normal(size=(500, 6)).astype(np.float32)
    Xn = X.copy()
    Xn[rng.random(X.shape) < 0.1] = np.nan
    y = X[:, 0] * 2 + X[:, 1] * X[:, 2]
    Xt = fx.grid_inputs(300, 6, seed=52)
    for model in (DecisionTreeRegressor(max_depth=5, random_state=0).fit(Xn, y),
                  RandomForestRegressor(n_estimators=7, max_depth=5, random_state=0).fit(Xn, y),
                  RandomForestClassifier(n_estimators=7, max_depth=4, random_state=0).fit(Xn, (y > 0).astype(int) + (y > 1))):
        kind, packed = packing.pack_model(model)
        assert kind == "trees" and packed.nan_ok and packed.default_left is not None
        got = emu.trees_predict(packed, Xt)
        np.testing.assert_allclose(got, model.predict(Xt.astype(np.float64)), rtol=1e-6, atol=1e-9)
    gbt = GradientBoostingRegressor(n_estimators=5, max_depth=3, random_state=0).fit(X, y)
    assert not packing.pack_model(gbt)[1].nan_ok  # its predict() refuses NaN, so the device flags the row
    plan = emulated_plan.EmulatedPlan(ColumnProgram([f"f{i}" for i in range(6)]), [packing.pack_model(gbt)])
    _out, status = plan.run(Xt, with_status=True)
    assert np.array_equal(status != 0, np.isnan(Xt).any(axis=1))
