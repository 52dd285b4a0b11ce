"""to_dict() for step / model / router objects (wire format of mlrun/serving/utils.py:46-109)."""

import inspect

_META = ("context", "name", "input_path", "result_path", "full_event", "kwargs")


class StepMeta:
    _STEP_KIND = None
    _dict_exclude = ()

    def to_dict(self, fields=None, exclude=None, strip=False):
        fields = fields or getattr(self, "_dict_fields", None) or list(inspect.signature(self.__init__).parameters)
        skip = set(exclude or ()) | set(self._dict_exclude)
        args = {}
        for key in fields:
            if key in _META or key in skip:
                continue
            val = getattr(self, key, None)
            if val is not None:
                args[key] = val
        if "kwargs" in fields:
            extra = getattr(self, "kwargs", None) or getattr(self, "_kwargs", None) or {}
            args.update({k: v for k, v in extra.items() if k not in _META})
        cls = type(self)
        path = cls.__qualname__ if cls.__module__ in ("__main__", "builtins") else f"{cls.__module__}.{cls.__qualname__}"
        out = {"class_name": path, "name": getattr(self, "name", None) or cls.__name__, "class_args": args}
        if self._STEP_KIND:
            out["kind"] = self._STEP_KIND
        for attr, key in (("_input_path", "input_path"), ("_result_path", "result_path")):
            if getattr(self, attr, None) is not None:
                out[key] = getattr(self, attr)
        if getattr(self, "_full_event", None):
            out["full_event"] = self._full_event
        return out
