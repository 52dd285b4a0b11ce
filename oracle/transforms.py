"""feature_store.steps transforms (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/feature_store/steps.py:
  get_engine :30-37, MLRunStep.do :53-67 (rebinds itself on the first event),
  FeaturesetValidator :94-149 (+ MinMaxValidator.check, mlrun/features.py:292-321),
  MapValues :152-246, Imputer :377-413, OneHotEncoder :427-513, DateExtractor :516-612,
  SetEventMetadata :635-696, DropFeatures :699-753.  `_do_spark` is out of scope.

Every step has the reference's two row engines: `_do_storey(event_dict)` (one event at a time) and
`_do_pandas(frame)`; `do` picks one from the first payload it sees and keeps it.  Quirks that are part of the observable
behaviour are kept and marked QUIRK.
"""

import re
import uuid

import numpy as np
import pandas as pd

from .helpers import MLRunInvalidArgumentError, get_in
from .step_io import StepToDict
from .topology import MapClass


def get_engine(first_event):
    payload = getattr(first_event, "body", first_event)
    if isinstance(payload, pd.DataFrame):
        return "pandas"
    return "spark" if hasattr(payload, "rdd") else "storey"


class MLRunStep(MapClass):
    """engine dispatch shared by the steps"""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._engine_to_do_method = {"pandas": self._do_pandas, "storey": self._do_storey}

    def do(self, event):
        engine = get_engine(event)
        handler = self._engine_to_do_method.get(engine)
        if handler is None:
            raise MLRunInvalidArgumentError(
                f"Unrecognized engine: {engine}. Available engines are: pandas, spark and storey")
        self.do = handler  # QUIRK: the instance method is replaced; later payloads never re-select the engine
        return handler(event)

    def _do_pandas(self, event):
        raise NotImplementedError

    def _do_storey(self, event):
        raise NotImplementedError


# ------------------------------------------------------------------------------------------ validation
def _shown(value, limit=40):
    """how a violating value is reported: its str, cut at 40 characters (mlrun/features.py:24-35)"""
    text = str(value)
    return text if len(text) <= limit else text[:limit] + "..."


class MinMaxValidator:
    """mlrun/features.py:265-321 (check only; `check_type` needs the feature's value type and is not restated)"""

    kind = "minmax"

    def __init__(self, check_type=None, severity=None, min=None, max=None):
        self.check_type, self.severity = check_type, severity  # QUIRK: no default severity -- reports start with "None!"
        self.min, self.max = min, max

    def check(self, value):
        """(ok, details).  A comparison that raises (None, a string against a number) is a violation whose message is the
        exception's text"""
        try:
            if self.min is not None and value < self.min:
                return False, {"message": "value is smaller than min", "min": self.min, "value": _shown(value)}
            if self.max is not None and value > self.max:
                return False, {"message": "value is greater than max", "max": self.max, "value": _shown(value)}
        except Exception as err:  # noqa: BLE001
            return False, {"message": str(err), "type": self.kind}
        return True, {}


class FeaturesetValidator(StepToDict, MLRunStep):
    """validation only prints; events pass through unchanged (steps.py:94-149).  The oracle takes the
    validators directly (`validators={col: MinMaxValidator}`) instead of a feature-set URI."""

    def __init__(self, featureset=None, columns=None, name=None, validators=None, **kwargs):
        super().__init__(**dict(kwargs, full_event=True))
        self.featureset = featureset or "."
        self.columns, self.name = columns, name
        self._validators = dict(validators or {})
        self.violations = 0

    def _do_storey(self, event):
        row = event.body
        where = f" key={event.key}" if event.key else ""
        for column, rule in self._validators.items():
            if column not in row:
                continue
            passed, details = rule.check(row[column])
            if passed:
                continue
            self.violations += 1
            what = details.pop("message")
            print(f"{rule.severity}! {column} {what},{where} args={details}")
        return event

    def _do_pandas(self, event):
        frame = event.body
        for column in frame:
            rule = self._validators.get(column)
            if not rule:
                continue
            failures = [d for ok, d in (rule.check(frame.at[i, column]) for i in frame.index) if not ok]
            if not failures:
                continue
            what = ""
            for d in failures:  # the report carries the LAST failure's message, and the details without it
                what = d.pop("message")
            self.violations += len(failures)
            print(f"{rule.severity}! {column} {what}, column={column}, has {len(failures)} violations args={failures}")
        return event


# ------------------------------------------------------------------------------------------ value maps
def _bounds(pair):
    """only these two spellings are understood: "-inf" as a lower and "inf" as an upper bound"""
    lo, hi = pair[0], pair[1]
    return (-np.inf if isinstance(lo, str) and lo == "-inf" else lo), (np.inf if isinstance(hi, str) and hi == "inf" else hi)


class MapValues(StepToDict, MLRunStep):
    def __init__(self, mapping, with_original_features=False, suffix="mapped", **kwargs):
        super().__init__(**kwargs)
        self.mapping, self.with_original_features, self.suffix = mapping, with_original_features, suffix

    def _get_feature_name(self, feature):
        return f"{feature}_{self.suffix}" if self.with_original_features else feature

    @classmethod
    def validate_args(cls, feature_set, **kwargs):
        """steps.py:331-370, run by FeatureSet.validate_steps at ingest: a column's replacement values share one type
        (NaN does not count; neither do the "-inf" / "inf" spellings of range bounds) and a column has ranges or single
        replacements, not both"""
        def counted(v):
            return not (isinstance(v, (float, np.float64, np.float32, np.float16)) and np.isnan(v))

        for column, rules in kwargs.get("mapping", []).items():
            if "ranges" not in rules:
                seen = {type(v) for v in rules.values() if counted(v)}
            elif len(rules) > 1:
                raise MLRunInvalidArgumentError(
                    f"MapValues - mapping values of the same column can not combine ranges and "
                    f"single replacement, which is the case for column '{column}'")
            else:
                seen = {type(v) for bounds in rules["ranges"].values() for v in bounds
                        if v != "-inf" and v != "inf" and counted(v)}
            if len(seen) > 1:
                raise MLRunInvalidArgumentError(
                    f"MapValues - mapping values of the same column must be in the"
                    f" same type, which was not the case for Column '{column}'")

    def _map_value(self, feature, value):
        """ranges: first [lo, hi) hit in dict order; else dict.get(value, value) (steps.py:189-201)"""
        rules = self.mapping.get(feature, {})
        for label, pair in (rules.get("ranges", {}) if "ranges" in rules else {}).items():
            lo, hi = _bounds(pair)
            if value >= lo and value < hi:  # operand order as upstream: a non-number raises with the same text
                return label
        return rules.get(value, value)

    def _do_storey(self, event):
        out = {self._get_feature_name(k): self._map_value(k, v) for k, v in event.items() if k in self.mapping}
        if self.with_original_features:
            out.update(event)
        return out

    def _do_pandas(self, event):
        """closed="both" ranges; unmapped -> None (steps.py:218-246)"""
        mapped = pd.DataFrame(index=event.index)
        for column in event.columns:
            rules = self.mapping.get(column, {})
            target = self._get_feature_name(column)
            if "ranges" in rules:
                for label in list(rules["ranges"]):  # QUIRK: the user's mapping is rewritten with numeric bounds
                    rules["ranges"][label] = list(_bounds(rules["ranges"][label]))
                table = pd.DataFrame.from_dict(rules["ranges"], "index").reset_index()
                table.index = pd.IntervalIndex.from_arrays(left=table[0], right=table[1], closed="both")
                mapped[target] = table.loc[event[column]]["index"].values
            elif rules:
                mapped[target] = event[column].map(lambda v, _r=rules: _r.get(v, None))
        return pd.concat([event, mapped], axis=1) if self.with_original_features else mapped


# ------------------------------------------------------------------------------------------ imputing
class Imputer(StepToDict, MLRunStep):
    def __init__(self, method="avg", default_value=None, mapping=None, **kwargs):
        super().__init__(**kwargs)
        self.mapping, self.method, self.default_value = mapping or {}, method, default_value

    def _impute(self, feature, value):
        return self.mapping.get(feature, self.default_value) if pd.isna(value) else value

    def _do_storey(self, event):
        """every feature of the dict is imputed (steps.py:397-406)"""
        return {k: self._impute(k, v) for k, v in event.items()}

    def _do_pandas(self, event):
        """columns whose fill is None are skipped (steps.py:408-413); the reference's in-place
        chained fillna is a no-op under pandas copy-on-write, so assign the filled column back"""
        fills = {c: self.mapping.get(c, self.default_value) for c in event.columns}
        for column, fill in fills.items():
            if fill is not None:
                event[column] = event[column].fillna(fill)
        return event


# ------------------------------------------------------------------------------------------ one-hot
class OneHotEncoder(StepToDict, MLRunStep):
    def __init__(self, mapping, **kwargs):
        super().__init__(**kwargs)
        self.mapping = mapping
        for feature in list(mapping):
            categories = mapping[feature]
            if not all(isinstance(c, (str, int, np.integer)) for c in categories):
                raise MLRunInvalidArgumentError("For OneHotEncoder you must provide int or string mapping list")
            mapping[feature] = list(dict.fromkeys(categories))  # de-duplicated in place, first occurrence wins

    @staticmethod
    def _sanitized_category(category):
        return re.sub("[ -]", "_", category) if isinstance(category, str) else category

    def _encode(self, feature, value):
        """steps.py:453-471"""
        categories = self.mapping.get(feature, [])
        if not categories:
            return {feature: value}
        key = lambda c: f"{feature}_{self._sanitized_category(c)}"  # noqa: E731
        fields = dict.fromkeys(map(key, categories), 0)
        if value in categories:
            fields[key(value)] = 1  # QUIRK: keyed by the VALUE's spelling -- 2.0 matching category 2 adds "<f>_2.0"
        elif self.logger:
            self.logger.warn(f"OneHotEncoder does not have an encoding for value '{value}' of feature '{feature}'")
        return fields

    def _do_storey(self, event):
        out = {}
        for feature, value in event.items():
            out.update(self._encode(feature, value))
        return out

    def _do_pandas(self, event):
        """steps.py:480-491"""
        for feature, categories in self.mapping.items():
            event[feature] = pd.Categorical(event[feature], categories=list(categories))
            dummies = pd.get_dummies(event[feature], prefix=feature, dtype=np.int64)
            dummies = dummies.rename(columns={c: self._sanitized_category(c) for c in dummies.columns})
            event = pd.concat([event.loc[:, :feature], dummies, event.loc[:, feature:]], axis=1)
        return event.drop(columns=list(self.mapping))


# ------------------------------------------------------------------------------------------ dates, metadata, drop
class DateExtractor(StepToDict, MLRunStep):
    def __init__(self, parts, timestamp_col=None, **kwargs):
        super().__init__(**kwargs)
        self.timestamp_col, self.parts = timestamp_col or "timestamp", parts

    def _get_key_name(self, part):
        return f"{self.timestamp_col}_{part}"

    def _extract_timestamp(self, event):
        if self.timestamp_col not in event:
            raise MLRunInvalidArgumentError(f"{self.timestamp_col} does not exist in the event")
        return event[self.timestamp_col]

    def _do_storey(self, event):
        moment = pd.Timestamp(self._extract_timestamp(event))
        event.update({self._get_key_name(p): getattr(moment, p) for p in self.parts})
        return event

    def _do_pandas(self, event):
        column = self._extract_timestamp(event)
        for part in self.parts:
            event[self._get_key_name(part)] = column.map(lambda v, _p=part: getattr(pd.Timestamp(v), _p))
        return event


class SetEventMetadata(MapClass):
    """copies id / key from body paths onto the event, or draws a random id (steps.py:635-696)"""

    _FIELDS = ("id_path", "key_path", "random_id")

    def __init__(self, id_path=None, key_path=None, random_id=None, **kwargs):
        super().__init__(**dict(kwargs, full_event=True))
        self.id_path, self.key_path, self.random_id = id_path, key_path, random_id
        self._tagging_funcs = []

    def to_dict(self, *a, **k):
        cls = type(self)
        args = {f: getattr(self, f) for f in self._FIELDS if getattr(self, f) is not None}
        return {"class_name": f"{cls.__module__}.{cls.__qualname__}", "name": self.name or cls.__name__,
                "class_args": args, "full_event": True}

    def post_init(self, mode="sync"):
        def from_body(attr, path):
            return lambda event: setattr(event, attr, str(get_in(event.body, path)))

        steps = []
        if self.id_path:
            steps.append(from_body("id", self.id_path))
        if self.key_path:
            steps.append(from_body("key", self.key_path))
        if self.random_id:
            steps.append(lambda event: setattr(event, "id", uuid.uuid4().hex))
        self._tagging_funcs = steps

    def do(self, event):
        for tag in self._tagging_funcs:
            tag(event)
        return event


class DropFeatures(StepToDict, MLRunStep):
    def __init__(self, features, **kwargs):
        super().__init__(**kwargs)
        self.features = features

    @classmethod
    def validate_args(cls, feature_set, **kwargs):
        """only features can be dropped: not an entity, not the label column, not the timestamp key (steps.py:737-753)"""
        doomed = kwargs.get("features", [])
        spec = feature_set.spec
        entities = set(doomed).intersection(spec.entities.keys())
        if entities:
            raise MLRunInvalidArgumentError(f"DropFeatures can only drop features, not entities: {entities}")
        protected = {"label_column": spec.label_column, "timestamp_key": spec.timestamp_key}
        for role, column in protected.items():
            if column in doomed:
                raise MLRunInvalidArgumentError(f"DropFeatures can not drop {role}: {column}")

    def _do_storey(self, event):
        for feature in self.features:
            if feature not in event:
                raise MLRunInvalidArgumentError(f"The ingesting data doesn't contain a feature named '{feature}'")
            del event[feature]
        return event

    def _do_pandas(self, event):
        return event.drop(columns=self.features)
