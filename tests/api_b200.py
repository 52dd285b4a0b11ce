"""`api` namespace over the product (mlrun_b200), see tests/scenarios.py"""

import json
import os

from mlrun_b200 import serving
from mlrun_b200.feature_store import steps as _steps
from mlrun_b200.feature_store.transforms import _Step
from mlrun_b200.serving import host

new_function = serving.new_function
V2ModelServer = serving.V2ModelServer
VotingEnsemble = serving.VotingEnsemble
ParallelRun = serving.ParallelRun
ModelRouter = serving.ModelRouter
MockEvent = serving.MockEvent
MockTrigger = serving.MockTrigger
GraphContext = serving.GraphContext
create_graph_server = serving.create_graph_server
RouterStep = serving.RouterStep
TaskStep = serving.TaskStep
MapClass = _Step
Imputer = _steps.Imputer
OneHotEncoder = _steps.OneHotEncoder
MapValues = _steps.MapValues
DropFeatures = _steps.DropFeatures
DateExtractor = _steps.DateExtractor
SetEventMetadata = _steps.SetEventMetadata
SKLearnModelServer = serving.SKLearnModelServer
FeatureRowVotingEnsemble = serving.FeatureRowVotingEnsemble
FeatureRowModelServer = serving.FeatureRowModelServer
NAME = "mlrun_b200"


def init_from_spec(spec, namespace):
    os.environ[host.SERVING_SPEC_ENV] = json.dumps(spec)
    context = GraphContext()
    context.is_mock = True
    host.nuclio_init_hook(context, namespace, "serving_v2")
    return context
