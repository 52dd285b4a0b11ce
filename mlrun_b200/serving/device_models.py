"""Device-backed model servers: the scikit-learn / xgboost-style scorers of the reference, on CUDA.

The reference serves fitted estimators through PickleModelServer.predict
(mlrun/frameworks/_ml_common/pkl_model_server.py:52-60: `np.asarray(inputs)` -> `model.predict(x)` ->
`.tolist()`); it *is* SKLearnModelServer and XGBoostModelServer (frameworks/sklearn/__init__.py:29,
frameworks/xgboost/__init__.py:30).  Here the fitted estimator is exported once (mlrun_b200.packing) and
`predict` runs the packed model as a DevicePlan.  There is no CPU fallback: without the CUDA library /
a GPU `predict` raises NativeError; an estimator that cannot be exported raises UnsupportedModel at
load time.

`feature_rows=True` lets a server (or a VotingEnsemble) sit directly behind feature-store steps: a
body that is a flat {feature: value} dict is treated as one input row -- the idiom of the reference's
EnrichmentVotingEnsemble.preprocess (serving/routers.py:1335-1342).
"""

import numpy as np

from .. import packing
from ..lowering import ColumnProgram
from .model_server import V2ModelServer
from .routing import VotingEnsemble


def _rows_of(inputs):
    """V2 `inputs` (list of rows, or one flat row) -> (B, F) float32 matrix"""
    if isinstance(inputs, np.ndarray):
        X = inputs
    else:
        X = np.asarray(inputs, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(1, -1)
    if X.ndim != 2:
        raise ValueError(f"inputs must be a list of rows, got shape {X.shape}")
    return np.ascontiguousarray(X, dtype=np.float32)


class PickleModelServer(V2ModelServer):
    """model= a fitted estimator, or model_path= a local .pkl (cloudpickle) file"""

    def __init__(self, context=None, name=None, model_path=None, model=None, protocol=None, input_path=None,
                 result_path=None, feature_rows=False, **kwargs):
        super().__init__(context=context, name=name, model_path=model_path, model=model, protocol=protocol,
                         input_path=input_path, result_path=result_path, **kwargs)
        self.feature_rows = feature_rows or None  # None keeps it out of to_dict() when unset
        self._packed = None
        self._plan = None

    def load(self):
        if self.model is None:
            if str(self.model_path or "").endswith(".json"):
                # an xgboost `save_model("m.json")` / LightGBM `dump_model()` document: the libraries' own portable format
                import json

                model_file, _ = self.get_model(".json")
                with open(model_file) as fp:
                    self.model = json.load(fp)
                return
            from cloudpickle import load

            model_file, _ = self.get_model(".pkl")
            with open(model_file, "rb") as fp:
                self.model = load(fp)

    @property
    def packed(self):
        """("linear", dict) | ("trees", PackedTrees) -- exported on first use, cached"""
        if self._packed is None:
            if self.model is None:
                self.load()
            self._packed = packing.pack_model(self.model)
        return self._packed

    def _own_plan(self, n_features):
        kind, packed = self.packed
        want = packed.get("n_features") if kind == "linear" else getattr(packed, "n_features", None)
        if want and int(want) != int(n_features):
            # scikit-learn's predict -> validate_data raises this for a width mismatch (extra columns included)
            raise ValueError(f"X has {n_features} features, but {self._model_kind()} is expecting {want} "
                             "features as input.")
        if self._plan is None or self._plan.n_in != n_features:
            self._plan = ColumnProgram([f"f{i}" for i in range(n_features)]).build_plan([self.packed])
        return self._plan

    def _model_kind(self):
        return "the tree ensemble" if isinstance(self.model, dict) else type(self.model).__name__

    def preprocess(self, request, operation):
        if self.feature_rows and isinstance(request, dict) and "inputs" not in request:
            request = {"inputs": [list(request.values())]}
        return request

    def predict(self, request):
        inputs = request["inputs"]
        if inputs and isinstance(inputs[0], dict):
            inputs = np.column_stack([np.asarray(v) for v in inputs[0].values()])
        X = _rows_of(inputs)
        out, status = self._own_plan(X.shape[1]).run(X, with_status=True)
        if status.any():
            # scikit-learn's check_array raises for NaN/Inf inputs; the whole request fails like there
            raise ValueError("Input X contains NaN or infinity.")
        return out[:, 0].tolist()

    def explain(self, request):
        return f"A model server named '{self.name}'"


SKLearnModelServer = PickleModelServer
XGBoostModelServer = PickleModelServer  # mlrun/frameworks/xgboost/__init__.py:30; models: live objects or save_model JSON


class LGBMModelServer(PickleModelServer):
    """mlrun/frameworks/lgbm/model_server.py:142-159 (`self.model.predict(x)`): a live LightGBM object, or the document
    `Booster.dump_model()` writes (model_path="model.json")"""


class FeatureRowModelServer(PickleModelServer):
    def __init__(self, context=None, name=None, model_path=None, model=None, protocol=None, input_path=None,
                 result_path=None, feature_rows=True, **kwargs):
        super().__init__(context=context, name=name, model_path=model_path, model=model, protocol=protocol,
                         input_path=input_path, result_path=result_path, feature_rows=True, **kwargs)


class FeatureRowVotingEnsemble(VotingEnsemble):
    """VotingEnsemble whose router-level preprocess turns a feature dict into a V2 `inputs` row"""

    feature_rows = True

    def preprocess(self, event):
        body = event.body
        if isinstance(body, dict) and "inputs" not in body:
            event.body = {"inputs": [list(body.values())]}
        return event
