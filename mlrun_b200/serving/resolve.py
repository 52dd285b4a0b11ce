"""Name resolution, errors and logging for the serving host.

Behaviour mirrored: mlrun/utils/helpers.py:1095-1174 (get_caller_globals / get_class / get_function with
"(expr)" lambda handlers), mlrun/errors.py (MLRunInvalidArgumentError is a ValueError), err_to_str.
Reference import paths ("mlrun.serving.routers.VotingEnsemble", ...) resolve to this package so that
specs and topologies written for the reference load unchanged.
"""

import importlib
import inspect
import logging
import sys
import types


class MLRunInvalidArgumentError(ValueError):
    pass


class GraphError(Exception):
    """error in graph topology or configuration"""


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


class _Log:
    def __init__(self, name="mlrun_b200"):
        self._l = logging.getLogger(name)

    def _w(self, lvl, msg, kw):
        self._l.log(lvl, f"{msg} {kw}" if kw else msg)

    def debug(self, msg, *a, **kw):
        self._w(logging.DEBUG, msg, kw)

    def info(self, msg, *a, **kw):
        self._w(logging.INFO, msg, kw)

    def warn(self, msg, *a, **kw):
        self._w(logging.WARNING, msg, kw)

    def error(self, msg, *a, **kw):
        self._w(logging.ERROR, msg, kw)

    warning = warn
    info_with = info
    debug_with = debug
    warn_with = warn
    error_with = error


logger = _Log()

# reference dotted paths -> this package
_REF_PATHS = {
    "mlrun.serving.ModelRouter": "mlrun_b200.serving.routing.ModelRouter",
    "mlrun.serving.routers.ModelRouter": "mlrun_b200.serving.routing.ModelRouter",
    "mlrun.serving.VotingEnsemble": "mlrun_b200.serving.routing.VotingEnsemble",
    "mlrun.serving.routers.VotingEnsemble": "mlrun_b200.serving.routing.VotingEnsemble",
    "mlrun.serving.routers.ParallelRun": "mlrun_b200.serving.routing.ParallelRun",
    "mlrun.serving.V2ModelServer": "mlrun_b200.serving.model_server.V2ModelServer",
    "mlrun.serving.v2_serving.V2ModelServer": "mlrun_b200.serving.model_server.V2ModelServer",
    "mlrun.frameworks.sklearn.SKLearnModelServer": "mlrun_b200.serving.device_models.SKLearnModelServer",
    "mlrun.frameworks.xgboost.XGBoostModelServer": "mlrun_b200.serving.device_models.XGBoostModelServer",
    "mlrun.frameworks.lgbm.LGBMModelServer": "mlrun_b200.serving.device_models.LGBMModelServer",
}
for _n in ("Imputer", "OneHotEncoder", "MapValues", "DropFeatures", "DateExtractor", "SetEventMetadata", "FeaturesetValidator"):
    _REF_PATHS[f"mlrun.feature_store.steps.{_n}"] = f"mlrun_b200.feature_store.transforms.{_n}"
    _REF_PATHS[f"mlrun.feature_store.{_n}"] = f"mlrun_b200.feature_store.transforms.{_n}"


def _lookup(name, namespaces):
    if not namespaces:
        return None
    for ns in namespaces if isinstance(namespaces, list) else [namespaces]:
        if ns is None:
            continue
        if isinstance(ns, types.ModuleType):
            obj = getattr(ns, name, None)
            if inspect.isfunction(obj) or isinstance(obj, type):
                return obj
        elif name in ns:
            return ns[name]
    return None


def _import(path):
    path = _REF_PATHS.get(path, path)
    if path.startswith("mlrun."):  # any other class of the reference's serving / feature-store packages, by its own path
        try:
            return _import("mlrun_b200." + path[len("mlrun."):])
        except ImportError:
            pass
    module, _, attr = path.rpartition(".")
    if not module:
        raise ImportError(f"cannot import {path!r}")
    mod = importlib.import_module(module)
    if not hasattr(mod, attr):
        raise ImportError(f"{attr} not found in {module}")
    return getattr(mod, attr)


def get_class(class_name, namespace=None):
    if isinstance(class_name, type):
        return class_name
    found = _lookup(class_name, namespace)
    if found is not None:
        return found
    try:
        return _import(class_name)
    except (ImportError, ValueError) as exc:
        raise ImportError(f"Failed to import {class_name}") from exc


def get_function(function, namespaces, reload_modules=False):
    if callable(function):
        return function
    function = function.strip()
    if function.startswith("("):
        if not function.endswith(")"):
            raise ValueError('function expression must start with "(" and end with ")"')
        return eval("lambda event: " + function[1:-1], {}, {})  # same contract as the reference's handler strings
    found = _lookup(function, namespaces)
    if found is not None:
        return found
    try:
        return _import(function)
    except (ImportError, ValueError) as exc:
        raise ImportError(f"state/function init failed, handler '{function}' not found") from exc


def caller_globals():
    """globals of the first frame outside this package"""
    frame = sys._getframe(1)
    while frame is not None:
        name = frame.f_globals.get("__name__", "")
        if not name.startswith("mlrun_b200"):
            return frame.f_globals
        frame = frame.f_back
    return None
