"""dict <-> object plumbing for graph objects (wire format of mlrun/model.py ModelObj / ObjectDict)."""

import copy
import inspect
import json


def _blank(v):
    return v is None or (isinstance(v, (dict, list)) and len(v) == 0)


class Serde:
    """objects whose wire form is {field: value} over `_dict_fields` (or the __init__ signature);
    None / empty containers are omitted, nested objects use their own to_dict"""

    _dict_fields = []

    def _wire_fields(self, fields=None):
        return fields or self._dict_fields or list(inspect.signature(self.__init__).parameters)

    def to_dict(self, fields=None, exclude=None, strip=False):
        out = {}
        for name in self._wire_fields(fields):
            if exclude and name in exclude:
                continue
            val = getattr(self, name, None)
            if hasattr(val, "to_dict"):
                val = val.to_dict()
            if not _blank(val):
                out[name] = val
        return out

    @classmethod
    def from_dict(cls, struct=None, fields=None, deprecated_fields=None):
        obj = cls()
        for name in obj._wire_fields(fields):
            if struct and name in struct:
                setattr(obj, name, struct[name])
        return obj

    @staticmethod
    def _verify_dict(param, name, new_type=None):
        if param is not None and not isinstance(param, dict) and not hasattr(param, "to_dict"):
            raise ValueError(f"Parameter {name} must be a dict or object")
        if new_type and (isinstance(param, dict) or param is None):
            return new_type.from_dict(param)
        return param

    def to_json(self, exclude=None, strip=False):
        return json.dumps(self.to_dict(exclude=exclude), default=str)

    def to_yaml(self, exclude=None, strip=False):
        import yaml

        return yaml.safe_dump(json.loads(self.to_json(exclude)), default_flow_style=False, sort_keys=False)

    def copy(self):
        return type(self).from_dict(copy.deepcopy(self.to_dict()))


class StepDict:
    """insertion-ordered {name: step}; dict children are built through the `kind` -> class map"""

    def __init__(self, classes, default_kind=""):
        self._items = {}
        self._classes = classes
        self._default = default_kind

    @classmethod
    def from_dict(cls, classes, children=None, default_kind=""):
        new = cls(classes, default_kind)
        if children is None:
            return new
        if not isinstance(children, dict):
            raise ValueError("children must be a dict")
        for name, child in children.items():
            new._items[name] = new._adopt(child, name)
        return new

    def _adopt(self, child, name):
        if getattr(child, "kind", None) in self._classes:
            child.name = name
            return child
        if isinstance(child, dict):
            kind = child.get("kind", self._default)
            if kind not in self._classes:
                raise ValueError(f"illegal object kind {kind}")
            obj = self._classes[kind].from_dict(child)
            obj.name = name
            return obj
        raise ValueError(f"illegal child (should be dict or child kind), {child}")

    def update(self, key, item):
        self._items[key] = self._adopt(item, key)
        return self._items[key]

    def to_dict(self, strip=False):
        return {k: v.to_dict() for k, v in self._items.items()}

    def values(self):
        return self._items.values()

    def keys(self):
        return self._items.keys()

    def items(self):
        return self._items.items()

    def __len__(self):
        return len(self._items)

    def __iter__(self):
        return iter(self._items)

    def __contains__(self, key):
        return key in self._items

    def __getitem__(self, key):
        return self._items[key]

    def __setitem__(self, key, item):
        self.update(key, item)

    def __delitem__(self, key):
        del self._items[key]
