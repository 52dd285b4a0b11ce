"""Step (de)serialisation + input/result path plumbing (oracle restatement; test infrastructure).

Follows mlrun/serving/utils.py:22-109.
"""

import inspect

from .helpers import get_in, update_in

event_id_key = "MLRUN-EVENT-ID"  # serving/utils.py:22
event_path_key = "MLRUN-EVENT-PATH"  # serving/utils.py:23


def _needs_mapping(body, which):
    if not hasattr(body, "__getitem__"):
        raise TypeError(f"{which} parameter supports only dict-like event bodies")


def _extract_input_data(input_path, body):
    """the part of the body a step sees (serving/utils.py:26-31)"""
    if input_path:
        _needs_mapping(body, "input_path")
        return get_in(body, input_path)
    return body


def _update_result_body(result_path, event_body, result):
    """the result is merged into the body only when there is a result_path AND a truthy body; else it replaces the body
    (serving/utils.py:34-43)"""
    if not (result_path and event_body):
        return result
    _needs_mapping(event_body, "result_path")
    update_in(event_body, result_path, result)
    return event_body


_META = ("context", "name", "input_path", "result_path", "full_event", "kwargs")


def _class_path(cls):
    return cls.__qualname__ if cls.__module__ in ("__main__", "builtins") else f"{cls.__module__}.{cls.__qualname__}"


class StepToDict:
    """a step object describes itself from its constructor signature (serving/utils.py:46-97):
    class_args = the non-None attributes named like constructor parameters (+ the saved **kwargs), minus the meta keys"""

    meta_keys = list(_META)

    def _constructor_fields(self, fields, exclude):
        names = fields or getattr(self, "_dict_fields", None) or list(inspect.signature(self.__init__).parameters)
        return [n for n in names if not (exclude and n in exclude)]

    def to_dict(self, fields=None, exclude=None, strip=False):
        names = self._constructor_fields(fields, exclude)
        args = {n: getattr(self, n) for n in names if n not in _META and getattr(self, n, None) is not None}
        if "kwargs" in names and (hasattr(self, "kwargs") or hasattr(self, "_kwargs")):
            saved = getattr(self, "kwargs", {}) or getattr(self, "_kwargs", {})
            args.update({k: v for k, v in saved.items() if k not in _META})
        out = {"class_name": _class_path(type(self)), "name": getattr(self, "name", None) or type(self).__name__,
               "class_args": args}
        if hasattr(self, "_STEP_KIND"):
            out["kind"] = self._STEP_KIND
        for key, attr in (("input_path", "_input_path"), ("result_path", "_result_path")):
            if getattr(self, attr, None) is not None:
                out[key] = getattr(self, attr)
        if getattr(self, "_full_event", None):
            out["full_event"] = self._full_event
        return out


class RouterToDict(StepToDict):
    """routers never serialise their routes (serving/utils.py:105-109)"""

    _STEP_KIND = "router"

    def to_dict(self, fields=None, exclude=None, strip=False):
        return StepToDict.to_dict(self, exclude=["routes"], strip=strip)
