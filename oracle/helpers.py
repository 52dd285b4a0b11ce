"""Small helpers the serving path leans on (oracle restatement; test infrastructure).

Follows (behaviour, not text):
  * mlrun/utils/helpers.py:446-459   get_in
  * mlrun/utils/helpers.py:476-519   update_in (with backslash-escaped dots)
  * mlrun/utils/helpers.py:1095-1174 get_class / get_function (incl. "(expr)" handlers)
  * mlrun/model.py:46-359            ModelObj.to_dict/from_dict, ObjectDict
"""

import importlib
import inspect
import logging
import sys
from types import ModuleType


class MLRunInvalidArgumentError(ValueError):
    """mirror of mlrun.errors.MLRunInvalidArgumentError (a ValueError subclass upstream too)"""


def err_to_str(err):
    """text of an exception and of its chain of causes (mlrun/errors.py:126-149): messages joined by ", caused by: ", an
    exception without a message shown by its repr, a chain that loops back cut where it repeats, more than 32 000 characters
    reduced to the first and last 16 000"""
    if not err:
        return ""
    if isinstance(err, str):
        return err
    chain, texts = [], []
    while err and err not in chain:
        chain.append(err)
        texts.append(str(err) or repr(err))
        err = err.__cause__
    text = ", caused by: ".join(texts)
    if len(text) > 32_000:
        text = text[:16_000] + "...truncated..." + text[-16_000:]
    return text


class _Logger:
    """tiny stand-in for mlrun.utils.logger (info/debug/warn/error + *_with)"""

    def __init__(self):
        self._log = logging.getLogger("oracle")

    def _emit(self, level, msg, kw):
        if kw:
            msg = f"{msg} {kw}"
        self._log.log(level, msg)

    def debug(self, msg, *a, **kw):
        self._emit(logging.DEBUG, msg, kw)

    def info(self, msg, *a, **kw):
        self._emit(logging.INFO, msg, kw)

    def warn(self, msg, *a, **kw):
        self._emit(logging.WARNING, msg, kw)

    warning = warn

    def error(self, msg, *a, **kw):
        self._emit(logging.ERROR, msg, kw)

    info_with = info
    debug_with = debug
    warn_with = warn
    error_with = error


logger = _Logger()


# --------------------------------------------------------------------------- paths
def get_in(obj, keys, default=None):
    """dotted-path read; a falsy intermediate or a missing key yields `default`
    (helpers.py:446-459)."""
    if isinstance(keys, str):
        keys = keys.split(".")
    cur = obj
    for key in keys:
        if not cur or key not in cur:
            return default
        cur = cur[key]
    return cur


def _split_escaped(key):
    out, buf, esc = [], "", False
    for ch in key:
        if ch == "." and not esc:
            out.append(buf)
            buf = ""
        elif ch == "\\":
            esc = not esc
        else:
            buf += ch
    out.append(buf)
    return out


_missing = object()


def update_in(obj, key, value, append=False, replace=True):
    """dotted-path write creating intermediate dicts (helpers.py:476-519)."""
    parts = _split_escaped(key) if isinstance(key, str) else key
    cur = obj
    for part in parts[:-1]:
        nxt = cur.get(part, _missing)
        if nxt is _missing:
            nxt = cur[part] = {}
        cur = nxt
    last = parts[-1]
    if last not in cur:
        cur[last] = [] if append else {}
    if append:
        if isinstance(value, list):
            cur[last] += value
        else:
            cur[last].append(value)
    elif replace or not cur.get(last):
        cur[last] = value


# --------------------------------------------------------------------------- class / function lookup
def _ns_as_dict(ns):
    if isinstance(ns, ModuleType):
        return {
            k: v
            for k, v in inspect.getmembers(
                ns, lambda o: inspect.isfunction(o) or isinstance(o, type)
            )
        }
    return ns


def _search(name, namespaces):
    if not namespaces:
        return None
    if not isinstance(namespaces, list):
        namespaces = [namespaces]
    for ns in namespaces:
        ns = _ns_as_dict(ns)
        if ns and name in ns:
            return ns[name]
    return None


# the reference resolves "mlrun.serving.X" style names by importing mlrun; the oracle maps the
# well-known upstream paths onto its own modules so specs written for the reference load here.
_ALIASES = {
    "mlrun.serving.ModelRouter": "oracle.ensembles.ModelRouter",
    "mlrun.serving.routers.ModelRouter": "oracle.ensembles.ModelRouter",
    "mlrun.serving.VotingEnsemble": "oracle.ensembles.VotingEnsemble",
    "mlrun.serving.routers.VotingEnsemble": "oracle.ensembles.VotingEnsemble",
    "mlrun.serving.routers.ParallelRun": "oracle.ensembles.ParallelRun",
    "mlrun.serving.V2ModelServer": "oracle.model_protocol.V2ModelServer",
    "mlrun.serving.v2_serving.V2ModelServer": "oracle.model_protocol.V2ModelServer",
    "mlrun.feature_store.steps.Imputer": "oracle.transforms.Imputer",
    "mlrun.feature_store.steps.OneHotEncoder": "oracle.transforms.OneHotEncoder",
    "mlrun.feature_store.steps.MapValues": "oracle.transforms.MapValues",
    "mlrun.feature_store.steps.DropFeatures": "oracle.transforms.DropFeatures",
    "mlrun.feature_store.steps.DateExtractor": "oracle.transforms.DateExtractor",
    "mlrun.feature_store.steps.SetEventMetadata": "oracle.transforms.SetEventMetadata",
    "mlrun.feature_store.steps.FeaturesetValidator": "oracle.transforms.FeaturesetValidator",
    "mlrun.frameworks.sklearn.SKLearnModelServer": "oracle.model_servers.SKLearnModelServer",
    "mlrun.frameworks.xgboost.XGBoostModelServer": "oracle.model_servers.XGBoostModelServer",
}


def _import_dotted(path):
    path = _ALIASES.get(path, path)
    if "." not in path:
        raise ImportError(f"cannot resolve {path}")
    mod_name, attr = path.rsplit(".", 1)
    module = importlib.import_module(mod_name)
    try:
        return getattr(module, attr)
    except AttributeError as exc:
        raise ImportError(f"{attr} not found in {mod_name}") from exc


def get_class(class_name, namespace=None):
    """class object from a name: namespaces first, then dotted import (helpers.py:1133-1145)."""
    if isinstance(class_name, type):
        return class_name
    found = _search(class_name, namespace)
    if found is not None:
        return found
    try:
        return _import_dotted(class_name)
    except (ImportError, ValueError) as exc:
        raise ImportError(f"Failed to import {class_name}") from exc


def get_function(function, namespaces, reload_modules=False):
    """callable from a name; "(expr)" becomes `lambda event: expr` (helpers.py:1148-1174)."""
    if callable(function):
        return function
    function = function.strip()
    if function.startswith("("):
        if not function.endswith(")"):
            raise ValueError('function expression must start with "(" and end with ")"')
        return eval("lambda event: " + function[1:-1], {}, {})
    found = _search(function, namespaces)
    if found is not None:
        return found
    try:
        return _import_dotted(function)
    except (ImportError, ValueError) as exc:
        raise ImportError(
            f"state/function init failed, handler '{function}' not found"
        ) from exc


def get_caller_globals(skip_prefixes=("oracle.", "oracle")):
    """globals of the first caller outside this package (helpers.py:1095-1108)."""
    try:
        frame = sys._getframe(2)
        while frame is not None:
            name = frame.f_globals.get("__name__", "")
            if not (name == "oracle" or name.startswith("oracle.")):
                return frame.f_globals
            frame = frame.f_back
    except Exception:
        return None
    return None


# --------------------------------------------------------------------------- model objects
def _empty(v):
    return v is None or (isinstance(v, (dict, list)) and not v)


class ModelObj:
    """to_dict/from_dict driven by `_dict_fields` or the __init__ signature (model.py:46-260).
    None and empty dict/list values are not serialised (model.py:160-181)."""

    _dict_fields = []

    @staticmethod
    def _verify_dict(param, name, new_type=None):
        if param is not None and not isinstance(param, dict) and not hasattr(param, "to_dict"):
            raise ValueError(f"Parameter {name} must be a dict or object")
        if new_type and (isinstance(param, dict) or param is None):
            return new_type.from_dict(param)
        return param

    def _fields(self, fields=None):
        return (
            fields
            or self._dict_fields
            or list(inspect.signature(self.__init__).parameters.keys())
        )

    def to_dict(self, fields=None, exclude=None, strip=False):
        out = {}
        for name in self._fields(fields):
            if exclude and name in exclude:
                continue
            val = getattr(self, name, None)
            if _empty(val):
                continue
            if hasattr(val, "to_dict"):
                val = val.to_dict()
                if _empty(val):
                    continue
            out[name] = val
        return out

    @classmethod
    def from_dict(cls, struct=None, fields=None, deprecated_fields=None):
        struct = {} if struct is None else struct
        obj = cls()
        for name in obj._fields(fields):
            if name in struct:
                setattr(obj, name, struct[name])
        return obj

    def to_yaml(self, exclude=None, strip=False):
        import yaml

        return yaml.safe_dump(self.to_dict(exclude=exclude), default_flow_style=False, sort_keys=False)

    def to_json(self, exclude=None, strip=False):
        import json

        return json.dumps(self.to_dict(exclude=exclude))

    def copy(self):
        import copy

        return self.from_dict(copy.deepcopy(self.to_dict()))


class ObjectDict:
    """ordered name -> step container that builds children from dicts by `kind` (model.py:117-205)."""

    def __init__(self, classes_map, default_kind=""):
        self._children = {}
        self._classes_map = classes_map
        self._default_kind = default_kind

    def values(self):
        return self._children.values()

    def keys(self):
        return self._children.keys()

    def items(self):
        return self._children.items()

    def __len__(self):
        return len(self._children)

    def __iter__(self):
        yield from self._children.keys()

    def __contains__(self, name):
        return name in self._children

    def __getitem__(self, name):
        return self._children[name]

    def __setitem__(self, key, item):
        self._children[key] = self._to_child(item, key)

    def __delitem__(self, key):
        del self._children[key]

    def update(self, key, item):
        child = self._to_child(item, key)
        self._children[key] = child
        return child

    def to_dict(self, strip=False):
        return {k: v.to_dict() for k, v in self._children.items()}

    @classmethod
    def from_dict(cls, classes_map, children=None, default_kind=""):
        if children is None:
            return cls(classes_map, default_kind)
        if not isinstance(children, dict):
            raise ValueError("children must be a dict")
        new = cls(classes_map, default_kind)
        for name, child in children.items():
            new._children[name] = new._to_child(child, name)
        return new

    def _to_child(self, child, name):
        if hasattr(child, "kind") and child.kind in self._classes_map:
            child.name = name
            return child
        if isinstance(child, dict):
            kind = child.get("kind", self._default_kind)
            if kind not in self._classes_map:
                raise ValueError(f"illegal object kind {kind}")
            obj = self._classes_map[kind].from_dict(child)
            obj.name = name
            return obj
        raise ValueError(f"illegal child (should be dict or child kind), {child}")

    def copy(self):
        import copy

        return ObjectDict.from_dict(self._classes_map, copy.deepcopy(self.to_dict()), self._default_kind)
