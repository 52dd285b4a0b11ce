"""Step (de)serialisation + input/result path plumbing (oracle restatement; test infrastructure).

Follows mlrun/serving/utils.py:22-109.
"""

import inspect

from .helpers import get_in, update_in

event_id_key = "MLRUN-EVENT-ID"  # serving/utils.py:22
event_path_key = "MLRUN-EVENT-PATH"  # serving/utils.py:23


def _extract_input_data(input_path, body):
    """serving/utils.py:26-31"""
    if not input_path:
        return body
    if not hasattr(body, "__getitem__"):
        raise TypeError("input_path parameter supports only dict-like event bodies")
    return get_in(body, input_path)


def _update_result_body(result_path, event_body, result):
    """merge only when result_path AND a truthy body; otherwise replace (serving/utils.py:34-43)"""
    if result_path and event_body:
        if not hasattr(event_body, "__getitem__"):
            raise TypeError("result_path parameter supports only dict-like event bodies")
        update_in(event_body, result_path, result)
        return event_body
    return result


class StepToDict:
    """auto-serialise a step object from its __init__ signature (serving/utils.py:46-97)"""

    meta_keys = ["context", "name", "input_path", "result_path", "full_event", "kwargs"]

    def to_dict(self, fields=None, exclude=None, strip=False):
        fields = fields or getattr(self, "_dict_fields", None)
        if not fields:
            fields = list(inspect.signature(self.__init__).parameters.keys())
        if exclude:
            fields = [f for f in fields if f not in exclude]

        args = {}
        for key in fields:
            if key in self.meta_keys:
                continue
            val = getattr(self, key, None)
            if val is not None:
                args[key] = val
        if "kwargs" in fields and (hasattr(self, "kwargs") or hasattr(self, "_kwargs")):
            extra = getattr(self, "kwargs", {}) or getattr(self, "_kwargs", {})
            for key, val in extra.items():
                if key not in self.meta_keys:
                    args[key] = val

        module = self.__class__.__module__
        path = self.__class__.__qualname__
        if module not in ("__main__", "builtins"):
            path = f"{module}.{path}"
        struct = {
            "class_name": path,
            "name": self.name if getattr(self, "name", None) else self.__class__.__name__,
            "class_args": args,
        }
        if hasattr(self, "_STEP_KIND"):
            struct["kind"] = self._STEP_KIND
        if getattr(self, "_input_path", None) is not None:
            struct["input_path"] = self._input_path
        if getattr(self, "_result_path", None) is not None:
            struct["result_path"] = self._result_path
        if getattr(self, "_full_event", None):
            struct["full_event"] = self._full_event
        return struct


class RouterToDict(StepToDict):
    """serving/utils.py:105-109"""

    _STEP_KIND = "router"

    def to_dict(self, fields=None, exclude=None, strip=False):
        return super().to_dict(exclude=["routes"], strip=strip)
