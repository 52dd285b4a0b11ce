"""CPU restatement of xgboost / LightGBM tree-ensemble prediction (oracle; test infrastructure -- only tests/, bench.py's
CPU legs and __graft_entry__.smoke() may import this).

The reference serves these libraries through `model.predict(x)`: `XGBoostModelServer = PickleModelServer`
(mlrun/frameworks/xgboost/__init__.py:30, frameworks/_ml_common/pkl_model_server.py:52-60) and `LGBMModelServer.predict`
(frameworks/lgbm/model_server.py:142-159).  The arithmetic is third-party -- xgboost ~=1.1 and lightgbm ~=4.3
(dev-requirements.txt) -- and neither package is under /root/reference nor installed in this image, and the reference's
tests assert no numeric prediction of such a model: **parity unpinned** for this file.  It restates the
libraries' published walks over their own serialised models, one row at a time, in plain Python:

  xgboost  (RegTree::GetNext, src/tree/tree_model.h; predict_fn.h):  missing (NaN) -> the node's default child;
           otherwise the left ("yes") child when fvalue < split_condition (float32 both), else the right one.  The margin
           is base_margin + the float32 sum of the leaves of the trees of the class (tree_info); reg:squarederror returns
           the margin, binary:logistic the class `margin > 0` (probability > 0.5), multi:soft* the argmax margin.
  LightGBM (Tree::NumericalDecision, include/LightGBM/tree.h):  NaN with missing_type != NaN is read as 0.0; a NaN under
           missing_type NaN takes default_left; otherwise left when fvalue <= threshold (float64 both).  Raw score = sum
           of leaf values (shrinkage included); binary -> raw > 0; multiclass -> argmax.
"""

import json
import math

import numpy as np


def _doc(d):
    return json.loads(d) if isinstance(d, (str, bytes, bytearray)) else d


# ------------------------------------------------------------------------------------------ xgboost
def xgboost_margins(model_json, X):
    """(B, n_groups) float64 margins of a `save_model` JSON document"""
    doc = _doc(model_json)
    learner = doc["learner"]
    booster = learner["gradient_booster"]
    model = booster["gbtree"]["model"] if "gbtree" in booster else booster["model"]
    lmp = learner["learner_model_param"]
    num_class = int(lmp.get("num_class", "0") or 0)
    groups = max(num_class, 1)
    base_score = float(lmp.get("base_score", "0.5"))
    objective = learner["objective"]["name"]
    base_margin = math.log(base_score / (1.0 - base_score)) if objective == "binary:logistic" else base_score
    tree_info = model.get("tree_info") or [0] * len(model["trees"])
    X = np.asarray(X, dtype=np.float32)
    out = np.zeros((len(X), groups), dtype=np.float64)
    for r, row in enumerate(X):
        psum = [np.float32(0.0)] * groups
        for ti, tree in enumerate(model["trees"]):
            left, right = tree["left_children"], tree["right_children"]
            cond, feat, dleft = tree["split_conditions"], tree["split_indices"], tree["default_left"]
            nid = 0
            while left[nid] != -1:
                fvalue = row[feat[nid]]
                if np.isnan(fvalue):
                    nid = left[nid] if dleft[nid] else right[nid]
                else:
                    nid = left[nid] if fvalue < np.float32(cond[nid]) else right[nid]
            g = tree_info[ti] if groups > 1 else 0
            psum[g] = np.float32(psum[g] + np.float32(cond[nid]))  # the leaf value lives in split_conditions
        for g in range(groups):
            out[r, g] = float(np.float32(base_margin)) + float(psum[g])
    return out, objective


def xgboost_predict(model_json, X):
    """what XGBRegressor / XGBClassifier .predict returns for the model: values, or class indices"""
    margins, objective = xgboost_margins(model_json, X)
    if objective.startswith("multi:"):
        return np.argmax(margins, axis=1)
    if objective.startswith("binary:"):
        return (margins[:, 0] > 0).astype(int)
    return margins[:, 0]


def xgboost_dump_predict(trees, X, base_score=0.5):
    """the same over `get_dump(dump_format="json")` documents (regression margin)"""
    X = np.asarray(X, dtype=np.float32)
    out = np.zeros(len(X), dtype=np.float64)
    roots = [_doc(t) for t in trees]
    for r, row in enumerate(X):
        psum = np.float32(0.0)
        for root in roots:
            node = root
            while "leaf" not in node:
                fvalue = row[int(str(node["split"])[1:])]
                kids = {c["nodeid"]: c for c in node["children"]}
                if np.isnan(fvalue):
                    node = kids[node["missing"]]
                else:
                    node = kids[node["yes"]] if fvalue < np.float32(node["split_condition"]) else kids[node["no"]]
            psum = np.float32(psum + np.float32(node["leaf"]))
        out[r] = float(np.float32(base_score)) + float(psum)
    return out


# ------------------------------------------------------------------------------------------ LightGBM
def lightgbm_raw(dump, X):
    doc = _doc(dump)
    num_class = int(doc.get("num_class", 1))
    per_iter = int(doc.get("num_tree_per_iteration", num_class))
    X = np.asarray(X, dtype=np.float32)
    out = np.zeros((len(X), max(per_iter, 1)), dtype=np.float64)
    for r, row in enumerate(X):
        for ti, info in enumerate(doc["tree_info"]):
            node = info["tree_structure"]
            while "split_feature" in node:
                fval = float(row[node["split_feature"]])
                missing = node.get("missing_type", "None")
                if math.isnan(fval) and missing != "NaN":
                    fval = 0.0
                if missing == "NaN" and math.isnan(fval):
                    go_left = bool(node.get("default_left", False))
                else:
                    go_left = fval <= float(node["threshold"])
                node = node["left_child"] if go_left else node["right_child"]
            out[r, ti % per_iter if per_iter > 1 else 0] += float(node.get("leaf_value", 0.0))
    return out, str(doc.get("objective", "regression")).split(" ")[0]


def lightgbm_predict(dump, X):
    raw, objective = lightgbm_raw(dump, X)
    if objective.startswith("multiclass"):
        return np.argmax(raw, axis=1)
    if objective in ("binary", "cross_entropy"):
        return (raw[:, 0] > 0).astype(int)
    return raw[:, 0]
