"""numpy emulation of a DevicePlan's arithmetic (tests only).

Lets the CPU suite validate `mlrun_b200.lowering` (fills / maps / schema) and `mlrun_b200.packing`
(sklearn -> device formats) against the oracle without a GPU.  It mirrors the kernel's semantics:
float32 inputs, float32 comparisons, float64 accumulation in column / tree order."""

import numpy as np

from mlrun_b200 import _native as nat


def transform(prog, X):
    """stage 1 + schema of a ColumnProgram -> expanded float32 matrix (B, n_out)"""
    X = np.array(X, dtype=np.float32, copy=True)
    for src, fill in prog.fills.items():
        col = X[:, src]
        col[np.isnan(col)] = np.float32(fill)
    for src, maps in prog.maps.items():
        x = X[:, src]
        for kind, m in maps:
            out = x.copy()
            done = np.zeros(len(x), dtype=bool)
            if kind == "value":
                for k, v in m.items():
                    hit = (~done) & (x == np.float32(k))
                    out[hit] = np.float32(v)
                    done |= hit
            else:
                for lo, hi, v in m:
                    hit = (~done) & (x >= np.float32(lo)) & (x < np.float32(hi))
                    out[hit] = np.float32(v)
                    done |= hit
            x = out
        X[:, src] = x
    cols = []
    for _name, src, kind, arg in prog.cols:
        if kind == nat.OUT_ONEHOT:
            cols.append((X[:, src] == np.float32(arg)).astype(np.float32))
        else:
            cols.append(X[:, src])
    return np.stack(cols, axis=1)


def link(scores, link_kind, classes):
    if link_kind == nat.LINK_IDENTITY:
        return scores[:, 0]
    if link_kind == nat.LINK_BINARY_GT:
        idx = (scores[:, 0] > 0).astype(int)
    elif link_kind == nat.LINK_BINARY_GE:
        idx = (scores[:, 0] >= 0).astype(int)
    else:
        idx = np.argmax(scores, axis=1)
    return idx if classes is None else np.asarray(classes)[idx]


def linear_predict(packed, E):
    scores = E.astype(np.float64) @ packed["W"].T + packed["b"]
    return link(scores, packed["link"], packed["classes"])


def trees_predict(t, E):
    B = E.shape[0]
    scores = np.tile(t.init, (B, 1)).astype(np.float64)
    for ti in range(t.n_trees):
        base = t.tree_offset[ti]
        node = np.zeros(B, dtype=np.int64)
        active = t.feature[base + node] >= 0
        while active.any():
            f = t.feature[base + node]
            thr = t.threshold[base + node]
            x = E[np.arange(B), np.where(f >= 0, f, 0)]
            go_left = x <= thr
            nxt = np.where(go_left, t.left[base + node], t.right[base + node])
            node = np.where(active, nxt, node)
            active = t.feature[base + node] >= 0
        scores[:, t.tree_slot[ti]] += t.tree_scale[ti] * t.leaf_value[base + node]
    return link(scores, t.link, t.classes)


def predict(models, E):
    out = []
    for kind, packed in models:
        out.append(linear_predict(packed, E) if kind == "linear" else trees_predict(packed, E))
    return np.stack(out, axis=1)
