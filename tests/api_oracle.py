"""`api` namespace over the CPU oracle (see tests/scenarios.py)."""

import os

import oracle
from oracle import ensembles, host, model_protocol, model_servers, topology, transforms

new_function = oracle.new_function
V2ModelServer = model_protocol.V2ModelServer
VotingEnsemble = ensembles.VotingEnsemble
ParallelRun = ensembles.ParallelRun
ModelRouter = ensembles.ModelRouter
MockEvent = host.MockEvent
MockTrigger = host.MockTrigger
GraphContext = host.GraphContext
create_graph_server = host.create_graph_server
RouterStep = topology.RouterStep
TaskStep = topology.TaskStep
MapClass = topology.MapClass
Imputer = transforms.Imputer
OneHotEncoder = transforms.OneHotEncoder
MapValues = transforms.MapValues
DropFeatures = transforms.DropFeatures
DateExtractor = transforms.DateExtractor
SetEventMetadata = transforms.SetEventMetadata
SKLearnModelServer = model_servers.SKLearnModelServer
NAME = "oracle"
from oracle.merger import Merge  # noqa: E402,F401
from oracle import enrichment as _enrichment  # noqa: E402

EnrichmentModelRouter = _enrichment.EnrichmentModelRouter
EnrichmentVotingEnsemble = _enrichment.EnrichmentVotingEnsemble
register_feature_vector = _enrichment.register_feature_vector
get_feature_vector = _enrichment.get_feature_vector


class FeatureRowVotingEnsemble(VotingEnsemble):
    """router-level preprocess turning a feature dict into a V2 `inputs` row (the idiom of
    EnrichmentVotingEnsemble.preprocess, serving/routers.py:1335-1342)"""

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
    """tests/serving/test_serving.py:218-228 (init_ctx): spec via env + nuclio init hook"""
    import json

    os.environ[host.SERVING_SPEC_ENV] = json.dumps(spec)
    context = GraphContext()
    context.is_mock = True
    host.nuclio_init_hook(context, namespace, "serving_v2")
    return context


def online_service(features, index_keys, table, stats, label_column, with_indexes, impute_policy):
    """tests/scenarios.py online_service_logic: the oracle's OnlineVectorService over the same dict table"""
    vec = _enrichment.FeatureVector("vec", features, index_keys, table, stats, label_column=label_column, with_indexes=with_indexes)
    return vec.get_online_feature_service(impute_policy)


def register_online_vector(uri, features, index_keys, table, stats, label_column, with_indexes):
    register_feature_vector(uri, _enrichment.FeatureVector("vec", features, index_keys, table, stats, label_column=label_column,
                                                         with_indexes=with_indexes))


def validator_step(rules, columns):
    from oracle import transforms as _t

    return _t.FeaturesetValidator(columns=columns, validators={c: _t.MinMaxValidator(**kw) for c, kw in rules.items()
                                                                 if not columns or c in columns})
