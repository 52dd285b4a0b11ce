"""`api` namespace over the REAL reference (only usable where /root/reference exists).

Generation-time only: imported by tests/golden/gen_golden.py, never by the test-suite.
"""

import json
import os

from tests.golden import _refshim

_refshim.install()

import mlrun  # noqa: E402
import mlrun.serving  # noqa: E402
import mlrun.serving.routers  # noqa: E402
import storey  # noqa: E402  (the shim's stub)
from mlrun.feature_store import steps as _steps  # noqa: E402
from mlrun.frameworks._ml_common.pkl_model_server import PickleModelServer  # noqa: E402
from mlrun.runtimes import nuclio_init_hook  # noqa: E402
from mlrun.serving import server as _server  # noqa: E402
from mlrun.serving import states as _states  # noqa: E402

new_function = mlrun.new_function
V2ModelServer = mlrun.serving.V2ModelServer
VotingEnsemble = mlrun.serving.routers.VotingEnsemble
ParallelRun = mlrun.serving.routers.ParallelRun
ModelRouter = mlrun.serving.routers.ModelRouter
MockEvent = _server.MockEvent
MockTrigger = _server.MockTrigger
GraphContext = _server.GraphContext
create_graph_server = _server.create_graph_server
RouterStep = _states.RouterStep
TaskStep = _states.TaskStep
MapClass = storey.MapClass
Imputer = _steps.Imputer
OneHotEncoder = _steps.OneHotEncoder
MapValues = _steps.MapValues
DropFeatures = _steps.DropFeatures
DateExtractor = _steps.DateExtractor
SetEventMetadata = _steps.SetEventMetadata
SKLearnModelServer = PickleModelServer
from mlrun.serving.merger import Merge  # noqa: E402

NAME = "reference"


class FeatureRowVotingEnsemble(VotingEnsemble):
    def preprocess(self, event):
        body = event.body
        if isinstance(body, dict) and "inputs" not in body:
            event.body = {"inputs": [list(body.values())]}
        return event


class FeatureRowModelServer(SKLearnModelServer):
    def preprocess(self, request, operation):
        if isinstance(request, dict) and "inputs" not in request:
            request = {"inputs": [list(request.values())]}
        return request


def init_from_spec(spec, namespace):
    os.environ["SERVING_SPEC_ENV"] = json.dumps(spec)
    context = GraphContext()
    context.is_mock = True
    nuclio_init_hook(context, namespace, "serving_v2")
    return context
