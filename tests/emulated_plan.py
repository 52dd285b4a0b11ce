"""A DevicePlan stand-in for the CPU suite (tests only): the numpy emulation of tests/device_emulator.py behind the
`run(X, with_status)` surface, so that the product's host layer -- graph building, lowering, packing, routers, the V2
protocol around device model servers, run_batch / run_events -- can be driven end to end without a GPU.  The CUDA kernels
themselves are compared with the oracle in the `-m gpu` tests; nothing in mlrun_b200 imports this."""

import numpy as np

from mlrun_b200 import _native as nat
from tests import device_emulator as emu


class EmulatedPlan:
    kernel = "numpy emulation (tests/emulated_plan.py)"
    finalized = True

    def __init__(self, prog, models=(), vote=None):
        self.prog, self.models, self.vote = prog, list(models), vote
        self.n_in = prog.n_in
        self.n_models = len(self.models)
        classes = [(m["classes"] if kind == "linear" else m.classes) is not None for kind, m in self.models]
        kind = vote[0] if vote is not None else nat.VOTE_NONE
        # b2s_plan_finalize: integer outputs for classifier labels / majority votes, never for a mean vote
        self.out_is_int = bool(self.models) and (any(classes) or kind == nat.VOTE_MAJORITY) and kind != nat.VOTE_MEAN
        self.out_cols = len(prog.cols) if not self.models else (self.n_models if kind == nat.VOTE_NONE else 1)

    @property
    def out_dtype(self):
        return np.int32 if self.out_is_int else np.float32

    def run(self, X, with_status=False, with_stats=False):
        if X.dtype != np.float32 or X.ndim != 2 or X.shape[1] != self.n_in:
            raise ValueError(f"rows must be a float32 (B, {self.n_in}) array with unit inner stride")
        E = emu.transform(self.prog, X) if len(X) else np.zeros((0, len(self.prog.cols)), dtype=np.float32)
        status = np.zeros(len(X), dtype=np.int32)
        if not self.models:
            out = E
        else:
            # NaN routing (b2s_trees3.cuh): when every model is a tree ensemble that routes missing values, the plan has
            # no feature steps but an Imputer, and no linear scorer, a NaN is data and only Inf flags the row
            nan_ok = (all(kind == "trees" and getattr(m, "nan_ok", False) for kind, m in self.models)
                      and not self.prog.maps and all(k == nat.OUT_COPY for _n, _s, k, _a in self.prog.cols)
                      and [s for _n, s, _k, _a in self.prog.cols] == list(range(self.n_in)))
            bad = np.isinf(E) if nan_ok else ~np.isfinite(E)
            status |= bad.any(axis=1).astype(np.int32) * nat.ROW_NONFINITE_INPUT
            with np.errstate(all="ignore"):
                per = emu.predict(self.models, np.where(bad, 0.0, E).astype(np.float32)) if len(X) else np.zeros((0, self.n_models))
            kind = self.vote[0] if self.vote is not None else nat.VOTE_NONE
            if kind == nat.VOTE_NONE:
                out = per
            elif kind == nat.VOTE_MEAN:  # VotingEnsemble._mean_vote: (B, M) @ w
                out = (per.astype(np.float64) @ np.asarray(self.vote[1], dtype=np.float64))[:, None]
            else:  # VotingEnsemble._majority_vote: weighted one-hot tally, first maximum
                labels = per.astype(np.int64)
                status |= (labels < 0).any(axis=1).astype(np.int32) * nat.ROW_BAD_LABEL
                safe = np.maximum(labels, 0)
                tally = np.zeros((len(X), int(safe.max(initial=0)) + 1), dtype=np.float64)
                for m, w in enumerate(self.vote[1]):
                    np.add.at(tally, (np.arange(len(X)), safe[:, m]), w)
                out = np.argmax(tally, axis=1)[:, None]
        res = (np.ascontiguousarray(out).astype(self.out_dtype),)
        if with_status:
            res += (status,)
        if with_stats:
            res += ({"rows": len(X), "kernels": 0},)
        return res if len(res) > 1 else res[0]

    def close(self):
        pass


def install(monkeypatch):
    """every plan the lowering builds from here on is emulated"""
    from mlrun_b200.lowering import ColumnProgram

    monkeypatch.setattr(ColumnProgram, "build_plan", lambda self, models=(), vote=None: EmulatedPlan(self, models, vote))
