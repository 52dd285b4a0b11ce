"""Pickle-based model servers (oracle restatement; test infrastructure).

Follows mlrun/frameworks/_ml_common/pkl_model_server.py:24-70; upstream aliases it as
SKLearnModelServer (frameworks/sklearn/__init__.py:29) and XGBoostModelServer
(frameworks/xgboost/__init__.py:30).  The arithmetic itself is third-party (scikit-learn): the
oracle calls scikit-learn's own `predict`, which is what the reference does.
"""

import numpy as np
import pandas as pd

from .model_protocol import V2ModelServer


class PickleModelServer(V2ModelServer):
    def load(self):
        from cloudpickle import load

        model_file, _ = self.get_model(".pkl")
        with open(model_file, "rb") as fp:
            self.model = load(fp)

    def predict(self, request):
        """np.asarray(inputs) -> model.predict -> tolist (pkl_model_server.py:52-60)"""
        inputs = request["inputs"]
        if inputs and isinstance(inputs[0], dict):
            x = pd.DataFrame(inputs[0])
        else:
            x = np.asarray(inputs)
        return self.model.predict(x).tolist()

    def explain(self, request):
        return f"A model server named '{self.name}'"


SKLearnModelServer = PickleModelServer
XGBoostModelServer = PickleModelServer
