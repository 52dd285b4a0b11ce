"""Hand-built xgboost / LightGBM models in the libraries' own serialisation formats (tests only; neither library is in the
image).  Thresholds and inputs are drawn from the same small grid, so `x == threshold` -- where `<` (xgboost) and `<=`
(scikit-learn, LightGBM) part ways -- and NaN routing are exercised on every tree."""

import numpy as np

GRID = np.array([-np.inf, -2.5, -1.0, -0.5, -0.0, 0.0, 0.25, 0.5, 1.0, 1.0000001, 3.0, 1e-40, -1e-40, np.inf], dtype=np.float32)


def grid_inputs(n_rows, n_feat, seed, nan_frac=0.15, with_inf=False):
    rng = np.random.default_rng(seed)
    grid = GRID if with_inf else GRID[np.isfinite(GRID)]
    X = rng.choice(grid, size=(n_rows, n_feat)).astype(np.float32)
    smooth = rng.random((n_rows, n_feat)) < 0.5
    X[smooth] = rng.normal(size=int(smooth.sum())).astype(np.float32)
    X[rng.random((n_rows, n_feat)) < nan_frac] = np.nan
    return X


def _random_shape(rng, depth, p_leaf):
    """-> list of nodes in BFS order: (left, right) child ids or None for a leaf; the root always splits"""
    nodes, frontier = [None], [(0, 0)]
    while frontier:
        nid, d = frontier.pop(0)
        if d >= depth or (d > 0 and rng.random() < p_leaf):
            continue
        li, ri = len(nodes), len(nodes) + 1
        nodes.extend([None, None])
        nodes[nid] = (li, ri)
        frontier.extend([(li, d + 1), (ri, d + 1)])
    return nodes


def random_xgb_model(n_trees=12, depth=4, n_feat=8, seed=0, objective="reg:squarederror", num_class=0, base_score=0.5,
                     p_leaf=0.2, with_neg_inf=True):
    """a `Booster.save_model("m.json")` document"""
    rng = np.random.default_rng(seed)
    thresholds = GRID[~np.isposinf(GRID)] if with_neg_inf else GRID[np.isfinite(GRID)]
    trees, tree_info = [], []
    groups = max(num_class, 1)
    for t in range(n_trees * groups):
        shape = _random_shape(rng, depth, p_leaf)
        n = len(shape)
        left = [-1] * n
        right = [-1] * n
        cond = [0.0] * n
        feat = [0] * n
        dleft = [0] * n
        for i, kids in enumerate(shape):
            if kids is None:
                cond[i] = float(np.float32(rng.normal() * 0.3))
            else:
                left[i], right[i] = kids
                feat[i] = int(rng.integers(0, n_feat))
                thr = rng.choice(thresholds) if rng.random() < 0.7 else np.float32(rng.normal())
                cond[i] = float(thr)  # json.dumps writes -Infinity for -inf, as xgboost does
                dleft[i] = int(rng.random() < 0.5)
        trees.append({"left_children": left, "right_children": right, "split_conditions": cond, "split_indices": feat,
                      "default_left": dleft, "split_type": [0] * n, "id": t,
                      "tree_param": {"num_nodes": str(n), "num_feature": str(n_feat)}})
        tree_info.append(t % groups)
    return {"learner": {"gradient_booster": {"name": "gbtree", "model": {
        "trees": trees, "tree_info": tree_info,
        "gbtree_model_param": {"num_trees": str(len(trees)), "num_parallel_tree": "1"}}},
        "learner_model_param": {"base_score": repr(float(base_score)), "num_class": str(num_class), "num_feature": str(n_feat)},
        "objective": {"name": objective}}, "version": [1, 7, 0]}


def xgb_doc_to_dump(doc):
    """the same trees as `Booster.get_dump(dump_format="json")` writes them (nested documents)"""
    out = []
    for tree in doc["learner"]["gradient_booster"]["model"]["trees"]:
        left, right = tree["left_children"], tree["right_children"]

        def node(i, depth, tree=tree, left=left, right=right):
            if left[i] == -1:
                return {"nodeid": i, "leaf": tree["split_conditions"][i]}
            return {"nodeid": i, "depth": depth, "split": f"f{tree['split_indices'][i]}",
                    "split_condition": tree["split_conditions"][i], "yes": left[i], "no": right[i],
                    "missing": left[i] if tree["default_left"][i] else right[i],
                    "children": [node(left[i], depth + 1), node(right[i], depth + 1)]}

        out.append(node(0, 0))
    return out


def random_lgbm_dump(n_trees=10, depth=4, n_feat=8, seed=0, objective="regression", num_class=1, p_leaf=0.2):
    """a `Booster.dump_model()` document"""
    rng = np.random.default_rng(seed)
    thresholds = GRID[np.isfinite(GRID)].astype(np.float64)
    infos = []
    for t in range(n_trees * max(num_class, 1)):
        shape = _random_shape(rng, depth, p_leaf)

        def build(i, shape=shape):
            if shape[i] is None:
                return {"leaf_index": i, "leaf_value": float(rng.normal() * 0.3)}
            thr = float(rng.choice(thresholds)) if rng.random() < 0.6 else float(rng.normal())
            if rng.random() < 0.3:
                thr += 1e-9  # a float64 threshold that is not a float32
            return {"split_index": i, "split_feature": int(rng.integers(0, n_feat)), "threshold": thr, "decision_type": "<=",
                    "default_left": bool(rng.random() < 0.5), "missing_type": str(rng.choice(["None", "NaN"])),
                    "left_child": build(shape[i][0]), "right_child": build(shape[i][1])}

        infos.append({"tree_index": t, "num_leaves": sum(1 for s in shape if s is None), "shrinkage": 0.1, "tree_structure": build(0)})
    return {"name": "tree", "version": "v4", "num_class": num_class, "num_tree_per_iteration": max(num_class, 1),
            "max_feature_idx": n_feat - 1, "objective": objective if num_class <= 1 else f"{objective} num_class:{num_class}",
            "average_output": False, "tree_info": infos}
