"""Export fitted scikit-learn estimators to the device model formats.

The reference serves these estimators through PickleModelServer.predict
(mlrun/frameworks/_ml_common/pkl_model_server.py:52-60 -> `self.model.predict(x)`); the arithmetic is
scikit-learn's.  This module reads the fitted attributes (`coef_/intercept_`, `tree_.*`) and restates
what `predict` computes as (scores, link):

  linear regressors       score = X @ coef_ + intercept_                       link identity
  LogisticRegression      score = X @ coef_.T + intercept_ ; > 0 / argmax      link binary_gt / argmax
  GradientBoosting*       score[k] = init[k] + lr * sum_stages tree[s][k](x)   link identity / binary_ge / argmax
  RandomForestRegressor   score = (1/n) * sum_t tree[t](x)                     link identity
  RandomForestClassifier  score[k] = (1/n) * sum_t proba_t[k](x)               link argmax
  DecisionTree*           one tree

Tree decisions: scikit-learn walks `X[i, feature] <= threshold` with X cast to float32 and float64
thresholds (sklearn/tree/_tree.pyx).  For a float32 x, `x <= t64` is equivalent to `x <= t32` where
t32 is t64 rounded toward -inf to float32 -- that is what is uploaded, so decisions are bit-identical.
"""

import numpy as np

from . import _native as nat
from .plan import PackedTrees


class UnsupportedModel(TypeError):
    pass


def threshold_to_f32(thr64):
    """largest float32 <= thr64 (so that float32 comparisons reproduce the float64 ones)"""
    thr64 = np.asarray(thr64, dtype=np.float64)
    t = thr64.astype(np.float32)
    up = t.astype(np.float64) > thr64
    t[up] = np.nextafter(t[up], np.float32(-np.inf))
    return t


def _int_classes(classes):
    classes = np.asarray(classes)
    if not np.issubdtype(classes.dtype, np.integer):
        as_int = classes.astype(np.int64)
        if not np.array_equal(as_int, classes):
            raise UnsupportedModel("device classifiers need integer class labels")
        classes = as_int
    return classes.astype(np.int32)


# Estimators whose predict() is exactly  X @ coef_.T + intercept_  followed by identity / "> 0" / argmax.  Anything
# else that happens to carry coef_ (GLMs with a log link, SVC(kernel="linear") with its one-vs-one rows, PLS with its
# centring ...) would be exported with the wrong arithmetic, so the export goes by class name, as pack_trees does.
LINEAR_REGRESSORS = frozenset({
    "LinearRegression", "Ridge", "RidgeCV", "Lasso", "LassoCV", "ElasticNet", "ElasticNetCV", "Lars", "LarsCV",
    "LassoLars", "LassoLarsCV", "LassoLarsIC", "OrthogonalMatchingPursuit", "OrthogonalMatchingPursuitCV",
    "BayesianRidge", "ARDRegression", "SGDRegressor", "HuberRegressor", "TheilSenRegressor",
    "PassiveAggressiveRegressor", "LinearSVR", "QuantileRegressor",
})
LINEAR_CLASSIFIERS = frozenset({
    "LogisticRegression", "LogisticRegressionCV", "SGDClassifier", "Perceptron", "PassiveAggressiveClassifier",
    "RidgeClassifier", "RidgeClassifierCV", "LinearSVC", "LinearDiscriminantAnalysis",
})


def pack_linear(model):
    """-> dict(W (K,F) f64, b (K,), link, classes)"""
    name = type(model).__name__
    if name not in LINEAR_REGRESSORS and name not in LINEAR_CLASSIFIERS:
        raise UnsupportedModel(
            f"{name} is not a plain linear estimator (predict != X @ coef_.T + intercept_ with identity / >0 / argmax)")
    if not hasattr(model, "coef_"):
        raise UnsupportedModel(f"{name} has no coef_")
    if name in LINEAR_CLASSIFIERS and not hasattr(model, "classes_"):
        raise UnsupportedModel(f"{name} has no classes_")
    coef = np.asarray(model.coef_, dtype=np.float64)
    intercept = np.atleast_1d(np.asarray(model.intercept_, dtype=np.float64))
    if name in LINEAR_CLASSIFIERS:
        classes = _int_classes(model.classes_)
        W = np.atleast_2d(coef)
        if W.shape[0] == 1:
            return dict(W=W, b=intercept[:1], link=nat.LINK_BINARY_GT, classes=classes)
        return dict(W=W, b=intercept, link=nat.LINK_ARGMAX, classes=classes)
    if coef.ndim == 2 and coef.shape[0] > 1:
        raise UnsupportedModel("multi-target linear regression is not lowered")
    W = coef.reshape(1, -1)
    b = intercept[:1] if intercept.size else np.zeros(1)
    return dict(W=W, b=b, link=nat.LINK_IDENTITY, classes=None)


def _append_tree(tree, value_of, acc):
    """append one sklearn Tree; value_of(node_values) -> scalar leaf value array"""
    t = tree
    n = t.node_count
    feat = np.where(t.children_left == -1, -1, t.feature).astype(np.int32)
    acc["feature"].append(feat)
    acc["threshold"].append(threshold_to_f32(np.where(feat >= 0, t.threshold, 0.0)))
    acc["left"].append(np.where(feat >= 0, t.children_left, 0).astype(np.int32))
    acc["right"].append(np.where(feat >= 0, t.children_right, 0).astype(np.int32))
    acc["leaf"].append(np.asarray(value_of(t.value), dtype=np.float64))
    acc["offset"].append(acc["offset"][-1] + n)
    # where a missing value goes (scikit-learn >= 1.3: Tree.missing_go_to_left; only meaningful for estimators whose
    # predict accepts NaN -- see NAN_ROUTING below)
    mgl = getattr(t, "missing_go_to_left", None)
    acc["default_left"].append(np.zeros(n, dtype=np.uint8) if mgl is None else np.asarray(mgl, dtype=np.uint8))


def _new_acc():
    return {"feature": [], "threshold": [], "left": [], "right": [], "leaf": [], "offset": [0], "slot": [], "scale": [],
            "default_left": []}


# estimators whose predict() lets NaN through and routes it by Tree.missing_go_to_left (scikit-learn >= 1.3 for single
# trees, >= 1.4 for forests; GradientBoosting* refuses NaN): for the others a NaN row stays an error, as in predict()
NAN_ROUTING = frozenset({"DecisionTreeRegressor", "DecisionTreeClassifier", "RandomForestRegressor", "RandomForestClassifier",
                         "ExtraTreesRegressor", "ExtraTreesClassifier"})


def _routes_nan(model):
    if type(model).__name__ not in NAN_ROUTING:
        return False
    tree = model.tree_ if hasattr(model, "tree_") else model.estimators_[0].tree_
    if not hasattr(tree, "missing_go_to_left"):
        return False
    try:  # the installed scikit-learn decides (the estimator tag its predict() consults)
        return bool(model.__sklearn_tags__().input_tags.allow_nan)
    except Exception:
        return False


def _finish(acc, init, link, classes, nan_ok=False):
    return PackedTrees(
        default_left=np.concatenate(acc["default_left"]) if nan_ok else None,
        nan_ok=nan_ok,
        tree_offset=np.asarray(acc["offset"], dtype=np.int32),
        feature=np.concatenate(acc["feature"]),
        threshold=np.concatenate(acc["threshold"]),
        left=np.concatenate(acc["left"]),
        right=np.concatenate(acc["right"]),
        leaf_value=np.concatenate(acc["leaf"]),
        tree_slot=np.asarray(acc["slot"], dtype=np.int32),
        tree_scale=np.asarray(acc["scale"], dtype=np.float64),
        init=np.asarray(init, dtype=np.float64),
        link=link,
        classes=classes,
    )


def pack_trees(model):
    """fitted sklearn tree ensemble -> PackedTrees"""
    name = type(model).__name__
    acc = _new_acc()
    if name in ("GradientBoostingRegressor", "GradientBoostingClassifier"):
        n_feat = model.n_features_in_
        init = np.asarray(model._raw_predict_init(np.zeros((1, n_feat), dtype=np.float64)), dtype=np.float64).ravel()
        stages, K = model.estimators_.shape
        for s in range(stages):
            for k in range(K):
                _append_tree(model.estimators_[s, k].tree_, lambda v: v[:, 0, 0], acc)
                acc["slot"].append(k)
                acc["scale"].append(float(model.learning_rate))
        if name == "GradientBoostingRegressor":
            return _finish(acc, init[:1], nat.LINK_IDENTITY, None)
        classes = _int_classes(model.classes_)
        if K == 1:
            return _finish(acc, init[:1], nat.LINK_BINARY_GE, classes)
        return _finish(acc, init[:K], nat.LINK_ARGMAX, classes)
    if name in ("RandomForestRegressor", "ExtraTreesRegressor"):
        n = len(model.estimators_)
        for est in model.estimators_:
            _append_tree(est.tree_, lambda v: v[:, 0, 0], acc)
            acc["slot"].append(0)
            acc["scale"].append(1.0 / n)
        return _finish(acc, [0.0], nat.LINK_IDENTITY, None, nan_ok=_routes_nan(model))
    if name in ("RandomForestClassifier", "ExtraTreesClassifier"):
        classes = _int_classes(model.classes_)
        n = len(model.estimators_)
        K = len(classes)
        for est in model.estimators_:
            for k in range(K):
                def proba_k(v, k=k):
                    v = v[:, 0, :]
                    return v[:, k] / v.sum(axis=1)
                _append_tree(est.tree_, proba_k, acc)
                acc["slot"].append(k)
                acc["scale"].append(1.0 / n)
        return _finish(acc, np.zeros(K), nat.LINK_ARGMAX, classes, nan_ok=_routes_nan(model))
    if name == "DecisionTreeRegressor":
        _append_tree(model.tree_, lambda v: v[:, 0, 0], acc)
        acc["slot"].append(0)
        acc["scale"].append(1.0)
        return _finish(acc, [0.0], nat.LINK_IDENTITY, None, nan_ok=_routes_nan(model))
    if name == "DecisionTreeClassifier":
        classes = _int_classes(model.classes_)
        K = len(classes)
        for k in range(K):
            def proba_k(v, k=k):
                v = v[:, 0, :]
                return v[:, k] / v.sum(axis=1)
            _append_tree(model.tree_, proba_k, acc)
            acc["slot"].append(k)
            acc["scale"].append(1.0)
        return _finish(acc, np.zeros(K), nat.LINK_ARGMAX, classes, nan_ok=_routes_nan(model))
    raise UnsupportedModel(f"{name} is not a supported tree ensemble")


def pack_model(model):
    """-> ("linear", dict) | ("trees", PackedTrees)"""
    from . import tree_formats  # (imports this module)

    if isinstance(model, PackedTrees):  # already exported, e.g. tree_formats.pack_xgboost_json(open("model.json").read())
        return "trees", model
    if isinstance(model, dict) and ("learner" in model or "tree_info" in model):  # a parsed xgboost / LightGBM document
        return "trees", tree_formats.pack_serialised(model)
    packed = tree_formats.pack_library_model(model)  # live xgboost / LightGBM objects
    if packed is not None:
        return "trees", packed
    name = type(model).__name__
    if name in LINEAR_REGRESSORS or name in LINEAR_CLASSIFIERS:
        packed = pack_linear(model)
        packed["n_features"] = int(packed["W"].shape[1])
        return "linear", packed
    if hasattr(model, "estimators_") or hasattr(model, "tree_"):
        packed = pack_trees(model)
        packed.n_features = int(getattr(model, "n_features_in_", 0)) or None  # the width predict() checks its input against
        return "trees", packed
    raise UnsupportedModel(f"{type(model).__name__}: only linear and tree-ensemble estimators are lowered")


def is_classifier_pack(kind, packed):
    link = packed["link"] if kind == "linear" else packed.link
    return link != nat.LINK_IDENTITY
