"""feature_store.steps transforms (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/feature_store/steps.py:
  get_engine :30-37, MLRunStep.do :53-67 (rebinds itself on the first event),
  FeaturesetValidator :94-149 (+ MinMaxValidator.check, mlrun/features.py:292-321),
  MapValues :152-246, Imputer :377-413, OneHotEncoder :427-513, DateExtractor :516-612,
  SetEventMetadata :635-696, DropFeatures :699-753.  `_do_spark` is out of scope.
"""

import re
import uuid
from collections import OrderedDict

import numpy as np
import pandas as pd

from .helpers import MLRunInvalidArgumentError, get_in
from .step_io import StepToDict
from .topology import MapClass


def get_engine(first_event):
    if hasattr(first_event, "body"):
        first_event = first_event.body
    if isinstance(first_event, pd.DataFrame):
        return "pandas"
    if hasattr(first_event, "rdd"):
        return "spark"
    return "storey"


class MLRunStep(MapClass):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._engine_to_do_method = {"pandas": self._do_pandas, "storey": self._do_storey}

    def do(self, event):
        engine = get_engine(event)
        self.do = self._engine_to_do_method.get(engine, None)
        if self.do is None:
            raise MLRunInvalidArgumentError(
                f"Unrecognized engine: {engine}. Available engines are: pandas, spark and storey"
            )
        return self.do(event)

    def _do_pandas(self, event):
        raise NotImplementedError

    def _do_storey(self, event):
        raise NotImplementedError


class MinMaxValidator:
    """mlrun/features.py:228-321 (check only)"""

    def __init__(self, check_type=None, severity=None, min=None, max=None):
        self.check_type = check_type
        self.severity = severity or "info"
        self.min = min
        self.max = max

    def check(self, value):
        if self.min is not None and value is not None and value < self.min:
            return False, {"message": "value is smaller than min", "min": self.min, "value": value}
        if self.max is not None and value is not None and value > self.max:
            return False, {"message": "value is greater than max", "max": self.max, "value": value}
        return True, {}


class FeaturesetValidator(StepToDict, MLRunStep):
    """validation only prints; events pass through unchanged (steps.py:94-149).  The oracle takes the
    validators directly (`validators={col: MinMaxValidator}`) instead of a feature-set URI."""

    def __init__(self, featureset=None, columns=None, name=None, validators=None, **kwargs):
        kwargs["full_event"] = True
        super().__init__(**kwargs)
        self.featureset = featureset or "."
        self.columns = columns
        self.name = name
        self._validators = dict(validators or {})
        self.violations = 0

    def _do_storey(self, event):
        body = event.body
        for name, validator in self._validators.items():
            if name in body:
                ok, args = validator.check(body[name])
                if not ok:
                    self.violations += 1
                    message = args.pop("message")
                    key_text = f" key={event.key}" if event.key else ""
                    print(f"{validator.severity}! {name} {message},{key_text} args={args}")
        return event

    def _do_pandas(self, event):
        body = event.body
        for column in body:
            validator = self._validators.get(column, None)
            if validator:
                violations, all_args, message = 0, [], ""
                for i in body.index:
                    ok, args = validator.check(body.at[i, column])
                    if not ok:
                        violations += 1
                        all_args.append(args)
                        message = args.pop("message")
                if violations:
                    self.violations += violations
                    print(f"{validator.severity}! {column} {message}, column={column}, has {violations} violations args={all_args}")
        return event


class MapValues(StepToDict, MLRunStep):
    def __init__(self, mapping, with_original_features=False, suffix="mapped", **kwargs):
        super().__init__(**kwargs)
        self.mapping = mapping
        self.with_original_features = with_original_features
        self.suffix = suffix

    def _map_value(self, feature, value):
        """ranges: first [lo, hi) hit in dict order; else dict.get(value, value) (steps.py:189-201)"""
        feature_map = self.mapping.get(feature, {})
        if "ranges" in feature_map:
            for val, val_range in feature_map.get("ranges", {}).items():
                lo = val_range[0] if val_range[0] != "-inf" else -np.inf
                hi = val_range[1] if val_range[1] != "inf" else np.inf
                if value >= lo and value < hi:
                    return val
        return feature_map.get(value, value)

    def _get_feature_name(self, feature):
        return f"{feature}_{self.suffix}" if self.with_original_features else feature

    def _do_storey(self, event):
        mapped = {
            self._get_feature_name(f): self._map_value(f, v) for f, v in event.items() if f in self.mapping
        }
        if self.with_original_features:
            mapped.update(event)
        return mapped

    def _do_pandas(self, event):
        """closed="both" ranges; unmapped -> None (steps.py:218-246)"""
        df = pd.DataFrame(index=event.index)
        for feature in event.columns:
            feature_map = self.mapping.get(feature, {})
            if "ranges" in feature_map:
                for val, val_range in feature_map.get("ranges", {}).items():
                    lo = val_range[0] if val_range[0] != "-inf" else -np.inf
                    hi = val_range[1] if val_range[1] != "inf" else np.inf
                    feature_map["ranges"][val] = [lo, hi]
                matchdf = pd.DataFrame.from_dict(feature_map["ranges"], "index").reset_index()
                matchdf.index = pd.IntervalIndex.from_arrays(left=matchdf[0], right=matchdf[1], closed="both")
                df[self._get_feature_name(feature)] = matchdf.loc[event[feature]]["index"].values
            elif feature_map:
                df[self._get_feature_name(feature)] = event[feature].map(lambda x: feature_map.get(x, None))
        if self.with_original_features:
            df = pd.concat([event, df], axis=1)
        return df


class Imputer(StepToDict, MLRunStep):
    def __init__(self, method="avg", default_value=None, mapping=None, **kwargs):
        super().__init__(**kwargs)
        self.mapping = mapping or {}
        self.method = method
        self.default_value = default_value

    def _impute(self, feature, value):
        if pd.isna(value):
            return self.mapping.get(feature, self.default_value)
        return value

    def _do_storey(self, event):
        """every feature of the dict is imputed (steps.py:397-406)"""
        return {feature: self._impute(feature, val) for feature, val in event.items()}

    def _do_pandas(self, event):
        """columns whose fill is None are skipped (steps.py:408-413); the reference's in-place
        chained fillna is a no-op under pandas copy-on-write, so assign the filled column back"""
        for feature in event.columns:
            val = self.mapping.get(feature, self.default_value)
            if val is not None:
                event[feature] = event[feature].fillna(val)
        return event


class OneHotEncoder(StepToDict, MLRunStep):
    def __init__(self, mapping, **kwargs):
        super().__init__(**kwargs)
        self.mapping = mapping
        for key, values in mapping.items():
            for val in values:
                if not (isinstance(val, str) or isinstance(val, (int, np.integer))):
                    raise MLRunInvalidArgumentError(
                        "For OneHotEncoder you must provide int or string mapping list"
                    )
            mapping[key] = list(OrderedDict.fromkeys(values).keys())

    def _encode(self, feature, value):
        """steps.py:453-471"""
        encoding = self.mapping.get(feature, [])
        if encoding:
            one_hot = {f"{feature}_{OneHotEncoder._sanitized_category(c)}": 0 for c in encoding}
            if value in encoding:
                one_hot[f"{feature}_{OneHotEncoder._sanitized_category(value)}"] = 1
            elif self.logger:
                self.logger.warn(
                    f"OneHotEncoder does not have an encoding for value '{value}' of feature '{feature}'"
                )
            return one_hot
        return {feature: value}

    def _do_storey(self, event):
        encoded = {}
        for feature, val in event.items():
            encoded.update(self._encode(feature, val))
        return encoded

    def _do_pandas(self, event):
        """steps.py:480-491"""
        for key, values in self.mapping.items():
            event[key] = pd.Categorical(event[key], categories=list(values))
            encoded = pd.get_dummies(event[key], prefix=key, dtype=np.int64)
            encoded.rename(columns={n: OneHotEncoder._sanitized_category(n) for n in encoded.columns}, inplace=True)
            event = pd.concat([event.loc[:, :key], encoded, event.loc[:, key:]], axis=1)
        event.drop(columns=list(self.mapping.keys()), inplace=True)
        return event

    @staticmethod
    def _sanitized_category(category):
        if isinstance(category, str):
            return re.sub("[ -]", "_", category)
        return category


class DateExtractor(StepToDict, MLRunStep):
    def __init__(self, parts, timestamp_col=None, **kwargs):
        super().__init__(**kwargs)
        self.timestamp_col = timestamp_col if timestamp_col else "timestamp"
        self.parts = parts

    def _get_key_name(self, part):
        return f"{self.timestamp_col}_{part}"

    def _extract_timestamp(self, event):
        try:
            return event[self.timestamp_col]
        except KeyError:
            raise MLRunInvalidArgumentError(f"{self.timestamp_col} does not exist in the event")

    def _do_storey(self, event):
        timestamp = pd.Timestamp(self._extract_timestamp(event))
        for part in self.parts:
            event[self._get_key_name(part)] = getattr(timestamp, part)
        return event

    def _do_pandas(self, event):
        timestamp = self._extract_timestamp(event)
        for part in self.parts:
            event[self._get_key_name(part)] = timestamp.map(lambda x: getattr(pd.Timestamp(x), part))
        return event


class SetEventMetadata(MapClass):
    def __init__(self, id_path=None, key_path=None, random_id=None, **kwargs):
        kwargs["full_event"] = True
        super().__init__(**kwargs)
        self.id_path = id_path
        self.key_path = key_path
        self.random_id = random_id
        self._tagging_funcs = []

    def to_dict(self, *a, **k):
        return {
            "class_name": f"{self.__class__.__module__}.{self.__class__.__qualname__}",
            "name": self.name or self.__class__.__name__,
            "class_args": {k: v for k, v in (("id_path", self.id_path), ("key_path", self.key_path),
                                               ("random_id", self.random_id)) if v is not None},
            "full_event": True,
        }

    def post_init(self, mode="sync"):
        def add_metadata(name, path, operator=str):
            def _add_meta(event):
                setattr(event, name, operator(get_in(event.body, path)))

            return _add_meta

        def set_random_id(event):
            event.id = uuid.uuid4().hex

        self._tagging_funcs = []
        if self.id_path:
            self._tagging_funcs.append(add_metadata("id", self.id_path))
        if self.key_path:
            self._tagging_funcs.append(add_metadata("key", self.key_path))
        if self.random_id:
            self._tagging_funcs.append(set_random_id)

    def do(self, event):
        for func in self._tagging_funcs:
            func(event)
        return event


class DropFeatures(StepToDict, MLRunStep):
    def __init__(self, features, **kwargs):
        super().__init__(**kwargs)
        self.features = features

    def _do_storey(self, event):
        for feature in self.features:
            try:
                del event[feature]
            except KeyError:
                raise MLRunInvalidArgumentError(
                    f"The ingesting data doesn't contain a feature named '{feature}'"
                )
        return event

    def _do_pandas(self, event):
        return event.drop(columns=self.features)
