"""Oracle for real-time feature enrichment (test infrastructure; never imported by the product).

Restates OnlineVectorService (mlrun/feature_store/feature_vector.py:903-1067) and the enrichment routers'
`preprocess` (mlrun/serving/routers.py:1118-1196, 1199-1342).  The reference resolves each entity row by emitting it
into a storey graph that reads the online (NoSQL) store; that read is storage (out of scope) and is restated here as a
dict lookup -- **parity unpinned** for the store access itself (the reference only tests it against a live v3io /
Redis), while everything around it (missing columns -> None, the impute policy incl. "$mean"-style statistics, index
removal, the all-falsy-row -> None quirk, `as_list`) follows the source line by line.
"""

from copy import copy

import numpy as np

from .ensembles import ModelRouter, VotingEnsemble
from .helpers import MLRunInvalidArgumentError

_REGISTRY = {}


def register_feature_vector(uri, vector):
    _REGISTRY[uri] = vector


def get_feature_vector(uri):
    return _REGISTRY[uri]


class FeatureVector:
    """what the online service reads from mlrun.feature_store.FeatureVector: feature names, index keys, label
    column, the stats table, and (instead of the NoSQL target) the online rows {key tuple: {feature: value}}"""

    def __init__(self, name, features, index_keys, table, stats=None, label_column=None, with_indexes=False):
        self.name = name
        self.features = list(features)
        self.index_keys = list(index_keys)
        self.table = table
        self.stats = stats
        self.label_column = label_column
        self.with_indexes = with_indexes

    def get_stats_table(self):
        return self.stats

    def get_online_feature_service(self, impute_policy=None):
        svc = OnlineVectorService(self, impute_policy)
        svc.initialize()
        return svc


class OnlineVectorService:
    def __init__(self, vector, impute_policy=None):
        self.vector = vector
        self.impute_policy = impute_policy or {}
        self._index_columns = vector.index_keys
        self._requested_columns = vector.features
        self._impute_values = {}

    def initialize(self):
        """feature_vector.py:935-968"""
        if not self.impute_policy:
            return
        impute_policy = copy(self.impute_policy)
        feature_stats = self.vector.get_stats_table()
        self._impute_values = {}
        feature_keys = list(self.vector.features)
        if self.vector.label_column in feature_keys:
            feature_keys.remove(self.vector.label_column)
        if "*" in impute_policy:
            value = impute_policy.pop("*")
            for name in feature_keys:
                if name not in impute_policy:
                    if isinstance(value, str) and value.startswith("$"):
                        self._impute_values[name] = feature_stats.loc[name, value[1:]]
                    else:
                        self._impute_values[name] = value
        for name, value in impute_policy.items():
            if name not in feature_keys:
                raise MLRunInvalidArgumentError(f"feature {name} in impute_policy but not in feature vector")
            if isinstance(value, str) and value.startswith("$"):
                self._impute_values[name] = feature_stats.loc[name, value[1:]]
            else:
                self._impute_values[name] = value

    def _read(self, row):
        """the storey QueryByKey graph: the entity row joined with what the online table holds for its key"""
        key = tuple(row[k] for k in self._index_columns)
        data = dict(row)
        data.update(self.vector.table.get(key, {}))
        return data

    def get(self, entity_rows, as_list=False):
        """feature_vector.py:975-1067"""
        results = []
        if isinstance(entity_rows, dict):
            entity_rows = [entity_rows]
        if not entity_rows or not isinstance(entity_rows, list) or not isinstance(entity_rows[0], (list, dict)):
            raise MLRunInvalidArgumentError(
                f"input data is of type {type(entity_rows)}. must be a list of lists or list of dicts")
        if isinstance(entity_rows[0], list):
            if not self._index_columns or len(entity_rows[0]) != len(self._index_columns):
                raise MLRunInvalidArgumentError("input list must be in the same size of the index_keys list")
            entity_rows = [{self._index_columns[i]: item[i] for i in range(len(self._index_columns))} for item in entity_rows]
        for row in entity_rows:
            data = self._read(row)
            if data:
                actual_columns = data.keys()
                if all(col in self._index_columns for col in actual_columns):
                    results.append(None)  # didn't get any data from the graph
                    continue
                for column in self._requested_columns:
                    if column not in actual_columns and column != self.vector.label_column:
                        data[column] = None
                if self._impute_values:
                    for name in data.keys():
                        v = data[name]
                        if v is None or (isinstance(v, float) and (np.isinf(v) or np.isnan(v))):
                            data[name] = self._impute_values.get(name, v)
                if not self.vector.with_indexes:
                    for name in self.vector.index_keys:
                        data.pop(name, None)
                if not any(data.values()):
                    data = None
            if as_list and data:
                data = [data.get(key, None) for key in self._requested_columns if key != self.vector.label_column]
            results.append(data)
        return results

    def close(self):
        pass


class _EnrichmentMixin:
    def _init_enrichment(self, feature_vector_uri, impute_policy):
        self.feature_vector_uri = feature_vector_uri
        self.impute_policy = impute_policy or {}
        self._feature_service = None

    def post_init(self, mode="sync"):
        super().post_init(mode)
        self._feature_service = get_feature_vector(self.feature_vector_uri).get_online_feature_service(
            impute_policy=self.impute_policy)

    def preprocess(self, event):
        """routers.py:1189-1196 / 1335-1342"""
        import json

        if isinstance(event.body, (str, bytes)):
            event.body = json.loads(event.body)
        event.body["inputs"] = self._feature_service.get(event.body["inputs"], as_list=True)
        return event


class EnrichmentModelRouter(_EnrichmentMixin, ModelRouter):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None, health_prefix=None,
                 feature_vector_uri="", impute_policy=None, **kwargs):
        super().__init__(context, name, routes, protocol, url_prefix, health_prefix, **kwargs)
        self._init_enrichment(feature_vector_uri, impute_policy)


class EnrichmentVotingEnsemble(_EnrichmentMixin, VotingEnsemble):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None, health_prefix=None,
                 vote_type=None, executor_type=None, prediction_col_name=None, feature_vector_uri="", impute_policy=None,
                 **kwargs):
        kw = dict(context=context, name=name, routes=routes, protocol=protocol, url_prefix=url_prefix,
                  health_prefix=health_prefix, vote_type=vote_type, prediction_col_name=prediction_col_name, **kwargs)
        if executor_type is not None:
            kw["executor_type"] = executor_type
        super().__init__(**kw)
        self._init_enrichment(feature_vector_uri, impute_policy)
