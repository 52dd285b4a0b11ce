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


def online_service(features, index_keys, table, stats, label_column, with_indexes, impute_policy):
    """the REAL OnlineVectorService (feature_store/feature_vector.py:903-1067) over a stub of the storey graph that reads
    the online store: emit(row) answers with the row joined with the table's values for its key"""
    import types

    from mlrun.feature_store.feature_vector import OnlineVectorService

    class _Controller:
        def emit(self, row, return_awaitable_result=True):
            data = dict(row)
            data.update(table.get(tuple(row[k] for k in index_keys), {}))
            return types.SimpleNamespace(await_result=lambda: types.SimpleNamespace(body=data))

        def terminate(self):
            pass

    vector = types.SimpleNamespace(
        status=types.SimpleNamespace(features={f: None for f in features}, label_column=label_column, index_keys=list(index_keys)),
        spec=types.SimpleNamespace(with_indexes=with_indexes), get_stats_table=lambda: stats)
    svc = OnlineVectorService(vector, types.SimpleNamespace(controller=_Controller()), list(index_keys), impute_policy,
                              requested_columns=list(features))
    svc.initialize()
    return svc


EnrichmentModelRouter = mlrun.serving.routers.EnrichmentModelRouter
EnrichmentVotingEnsemble = mlrun.serving.routers.EnrichmentVotingEnsemble
_ONLINE_VECTORS = {}


def register_online_vector(uri, features, index_keys, table, stats, label_column, with_indexes):
    """what `get_feature_vector(uri).get_online_feature_service(impute_policy=...)` answers for the REAL enrichment
    routers (serving/routers.py:1180-1187): the REAL OnlineVectorService over the stub store read"""
    import types

    import mlrun.feature_store

    _ONLINE_VECTORS[uri] = types.SimpleNamespace(get_online_feature_service=lambda impute_policy=None: online_service(
        features, index_keys, table, stats, label_column, with_indexes, impute_policy))
    mlrun.feature_store.get_feature_vector = lambda u, *a, **k: _ONLINE_VECTORS[u]


def validator_step(rules, columns):
    """the REAL FeaturesetValidator (feature_store/steps.py:94-149) over the REAL MinMaxValidator / Feature (mlrun/features.py);
    the feature set it reads its validators from is a stub of the store resource"""
    import types

    from mlrun.features import Feature, MinMaxValidator

    feats = {col: Feature(validator=MinMaxValidator(**kw)) for col, kw in rules.items()}
    ctx = types.SimpleNamespace(get_store_resource=lambda uri: types.SimpleNamespace(spec=types.SimpleNamespace(features=feats)),
                                logger=None)
    return _steps.FeaturesetValidator(featureset=".", columns=columns, context=ctx)
