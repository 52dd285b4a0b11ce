"""Pickle-based model servers (oracle restatement; test infrastructure).

Follows mlrun/frameworks/_ml_common/pkl_model_server.py:24-70; upstream aliases it as
SKLearnModelServer (frameworks/sklearn/__init__.py:29) and XGBoostModelServer
(frameworks/xgboost/__init__.py:30).  The arithmetic itself is third-party (scikit-learn): the
oracle calls scikit-learn's own `predict`, which is what the reference does.
"""

import numpy as np
import pandas as pd

from .model_protocol import V2ModelServer


def _as_model_input(inputs):
    """a list whose first item is a dict becomes a frame OF THAT FIRST DICT; anything else an ndarray"""
    first_is_dict = bool(inputs) and isinstance(inputs[0], dict)
    return pd.DataFrame(inputs[0]) if first_is_dict else np.asarray(inputs)


class PickleModelServer(V2ModelServer):
    def explain(self, request):
        return f"A model server named '{self.name}'"

    def predict(self, request):
        """pkl_model_server.py:52-60: the third-party model's own predict, then tolist"""
        return self.model.predict(_as_model_input(request["inputs"])).tolist()

    def load(self):
        import cloudpickle

        path, _extra = self.get_model(".pkl")
        with open(path, "rb") as handle:
            self.model = cloudpickle.load(handle)


SKLearnModelServer = PickleModelServer
XGBoostModelServer = PickleModelServer
