"""Vectorised numpy restatement of the serving math on (B, F) arrays (oracle; test infrastructure).

The reference runs this math one event at a time (oracle/transforms.py, oracle/ensembles.py restate
that).  These functions compute the same values for a whole batch in float64 -- they are the checker
for the CUDA path at batch sizes and the "vectorised numpy / scikit-learn" CPU baseline of
SURVEY.md §8(d).  tests/test_oracle_batch.py proves they agree with the per-event restatement.
"""

import numpy as np


def impute(X, names, mapping=None, default_value=None):
    """Imputer._impute over every column (feature_store/steps.py:397-406)"""
    X = np.array(X, dtype=np.float64, copy=True)
    mapping = mapping or {}
    for j, n in enumerate(names):
        fill = mapping.get(n, default_value)
        if fill is None:
            continue
        col = X[:, j]
        col[np.isnan(col)] = fill
    return X


def one_hot(X, names, mapping):
    """OneHotEncoder._do_storey (steps.py:453-478): mapped columns are replaced in place"""
    cols, out_names = [], []
    for j, n in enumerate(names):
        cats = mapping.get(n)
        if not cats:
            cols.append(X[:, j])
            out_names.append(n)
            continue
        seen = list(dict.fromkeys(cats))
        for c in seen:
            cols.append((X[:, j] == c).astype(np.float64))
            out_names.append(f"{n}_{c}")
    return np.stack(cols, axis=1), out_names


def map_values(X, names, mapping):
    """MapValues._do_storey with with_original_features=False (steps.py:189-216), numeric maps only"""
    cols, out_names = [], []
    for j, n in enumerate(names):
        if n not in mapping:
            continue
        fmap = mapping[n]
        x = X[:, j]
        out = x.copy()
        if "ranges" in fmap:
            done = np.zeros(len(x), dtype=bool)
            for val, (lo, hi) in fmap["ranges"].items():
                lo = -np.inf if lo == "-inf" else lo
                hi = np.inf if hi == "inf" else hi
                hit = (~done) & (x >= lo) & (x < hi)
                out[hit] = val
                done |= hit
        else:
            done = np.zeros(len(x), dtype=bool)
            for k, v in fmap.items():
                hit = (~done) & (x == k)
                out[hit] = v
                done |= hit
        cols.append(out)
        out_names.append(n)
    return np.stack(cols, axis=1), out_names


def drop(X, names, features):
    keep = [j for j, n in enumerate(names) if n not in set(features)]
    return X[:, keep], [names[j] for j in keep]


def mean_vote(preds, weights):
    """VotingEnsemble._mean_vote (serving/routers.py:732-741): (n,m) @ w(m)"""
    return np.array(preds, dtype=np.float64) @ np.asarray(weights, dtype=np.float64)


def majority_vote(preds, weights):
    """VotingEnsemble._majority_vote (serving/routers.py:708-730)"""
    preds = np.array(preds).astype(int)
    one_hot_rep = np.transpose((np.arange(preds.max() + 1) == preds[..., None]).astype(int), (0, 2, 1))
    return np.argmax(one_hot_rep @ np.asarray(weights, dtype=np.float64), axis=1)


def flow3(wl):
    """Imputer -> OneHotEncoder -> linear model(s) -> (mean vote) for a Flow3Workload.
    returns dict(expanded, per_model (B,M), out (B,))"""
    X = impute(wl.X, wl.names, wl.impute_mapping, wl.impute_default)
    E, out_names = one_hot(X, wl.names, wl.onehot_mapping)
    per_model = np.stack([m.predict(E) for m in wl.sklearn_models()], axis=1)
    if wl.n_models == 1:
        out = per_model[:, 0]
    else:
        out = mean_vote(per_model, np.full(wl.n_models, 1.0 / wl.n_models))
    return {"expanded": E, "names": out_names, "per_model": per_model, "out": out}


def tree_ensemble(wl, weights=None):
    """PickleModelServer.predict per model (np.asarray(inputs) -> model.predict) + VotingEnsemble vote"""
    X = wl.X.astype(np.float64)
    per_model = np.stack([np.asarray(m.predict(X)) for m in wl.models], axis=1)
    M = len(wl.models)
    w = np.full(M, 1.0 / M) if weights is None else np.asarray(weights, dtype=np.float64)
    if wl.kind == "regression":
        out = mean_vote(per_model, w)
    else:
        out = majority_vote(per_model, w)
    return {"per_model": per_model, "out": out}
