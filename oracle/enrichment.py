"""Oracle for real-time feature enrichment (test infrastructure; never imported by the product).

Restates OnlineVectorService (mlrun/feature_store/feature_vector.py:903-1067) and the enrichment routers'
`preprocess` (mlrun/serving/routers.py:1118-1196, 1199-1342).  The reference resolves each entity row by emitting it
into a storey graph that reads the online (NoSQL) store; that read is storage (out of scope) and is restated here as a
dict lookup -- **parity unpinned** for the store access itself (the reference only tests it against a live v3io /
Redis).  Everything around it is PINNED: the golden scenario `online_service_logic` drives the REAL OnlineVectorService
(initialize + get) over a stub controller that does the same dict lookup, and this class must give the same answers
(missing columns -> None, the impute policy incl. "$mean"-style statistics, index removal, the all-falsy-row -> None
quirk, `as_list`, the argument errors).
"""

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
        """resolve the impute policy into {feature: value} (feature_vector.py:935-968): "*" is the default for every feature
        (the label excluded) that has no entry of its own; "$mean"-style values are looked up in the stats table"""
        policy = dict(self.impute_policy)
        if not policy:
            return
        stats = self.vector.get_stats_table()
        imputable = [f for f in self.vector.features if f != self.vector.label_column]

        def resolved(feature, value):
            return stats.loc[feature, value[1:]] if isinstance(value, str) and value.startswith("$") else value

        values = {}
        if "*" in policy:
            default = policy.pop("*")
            values = {f: resolved(f, default) for f in imputable if f not in policy}
        for feature, value in policy.items():
            if feature not in imputable:
                raise MLRunInvalidArgumentError(f"feature {feature} in impute_policy but not in feature vector")
            values[feature] = resolved(feature, value)
        self._impute_values = values

    def _read(self, row):
        """the storey QueryByKey graph: the entity row joined with what the online table holds for its key"""
        found = self.vector.table.get(tuple(row[k] for k in self._index_columns), {})
        return {**row, **found}

    def _entity_dicts(self, entity_rows):
        """a dict is one row; rows given as lists are zipped with the index columns (feature_vector.py:997-1022)"""
        rows = [entity_rows] if isinstance(entity_rows, dict) else entity_rows
        if not (rows and isinstance(rows, list) and isinstance(rows[0], (list, dict))):
            raise MLRunInvalidArgumentError(f"input data is of type {type(rows)}. must be a list of lists or list of dicts")
        if isinstance(rows[0], dict):
            return rows
        names = self._index_columns
        if not names or len(rows[0]) != len(names):
            raise MLRunInvalidArgumentError("input list must be in the same size of the index_keys list")
        return [{names[i]: item[i] for i in range(len(names))} for item in rows]

    def _finish(self, data):
        """one answer of the graph -> the row handed back, or None (feature_vector.py:1027-1056)"""
        if not data:
            return data
        if all(column in self._index_columns for column in data):
            return None  # only the entity columns came back: nothing is stored for this key
        label = self.vector.label_column
        for column in self._requested_columns:
            if column != label:
                data.setdefault(column, None)
        if self._impute_values:
            for column, v in data.items():
                if v is None or (isinstance(v, float) and not np.isfinite(v)):
                    data[column] = self._impute_values.get(column, v)
        if not self.vector.with_indexes:
            for column in self.vector.index_keys:
                data.pop(column, None)
        return data if any(data.values()) else None  # QUIRK: an all-falsy row (zeros) is reported as missing

    def get(self, entity_rows, as_list=False):
        """feature_vector.py:975-1067"""
        label = self.vector.label_column
        results = []
        for row in self._entity_dicts(entity_rows):
            data = self._finish(self._read(row))
            if as_list and data:
                data = [data.get(column) for column in self._requested_columns if column != label]
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
