"""Model routers and the voting ensemble (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/serving/routers.py:
  BaseModelRouter :43-164, ModelRouter :167-211, ParallelRun :245-477,
  VotingEnsemble :480-991 (_resolve_route :623-706, _majority_vote :708-730, _mean_vote :732-741,
  logic :746-787, _apply_logic :789-810, do_event :812-914, _normalize_weights :962-991).
Process-pool execution (routers.py:378-396) is out of scope for the oracle; "process" falls back to
the thread pool (same results, same key-order caveat).

Layout of this restatement: every router's `do_event` is `_enter` (input path, preprocess hook, body parsing, health /
prefix checks) + a class-specific middle + `_leave` (result path); URL targets are split once by `_target_segments`.
Behavioural quirks of the reference that callers can observe are kept and marked QUIRK.
"""

import concurrent.futures
import copy
import json
import traceback
from enum import Enum
from io import BytesIO

import numpy as np

from .helpers import logger
from .model_protocol import _ModelLogPusher, now_date
from .step_io import RouterToDict, _extract_input_data, _update_result_body


class BaseModelRouter(RouterToDict):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None,
                 health_prefix=None, input_path=None, result_path=None, **kwargs):
        self.name, self.context, self.routes = name, context, routes
        self.protocol = protocol or "v2"
        root = f"/{self.protocol}"
        self.url_prefix = url_prefix or f"{root}/models"
        self.health_prefix = health_prefix or f"{root}/health"
        self.inputs_key = "inputs" if self.protocol != "v1" else "instances"
        self._input_path, self._result_path = input_path, result_path
        self.kwargs = kwargs

    # ---- hooks ------------------------------------------------------------------------------------
    def post_init(self, mode="sync"):
        self.context.logger.info(f"Loaded {list(self.routes.keys())}")

    def get_metadata(self):
        return {"name": type(self).__name__, "version": "v2", "extensions": []}

    def preprocess(self, event):
        return event

    def postprocess(self, event):
        return event

    def _handle_event(self, event):
        return event

    # ---- shared prologue / epilogue -------------------------------------------------------------------
    def parse_event(self, event):
        """routers.py:86-112 (data_url download is I/O: out of scope)"""
        raw = event.body
        try:
            return raw if isinstance(raw, dict) else json.loads(raw)
        except Exception as exc:
            if (getattr(event, "content_type", "") or "").startswith("image/"):
                return {self.inputs_key: [BytesIO(raw)]}
            raise ValueError("Unrecognized request format") from exc

    def _pre_handle_event(self, event):
        """routers.py:122-141 -- body parsing, health / metadata answers, prefix check"""
        is_get = (event.method or "POST") == "GET"
        if event.body and not is_get:
            event.body = self.parse_event(event)
        path = getattr(event, "path", "")
        if is_get and (path == "/" or path.startswith(self.health_prefix)):
            event.terminated = True
            event.body = self.get_metadata()
        elif path and path != "/" and not path.startswith(self.url_prefix):
            raise ValueError(f"illegal path prefix {path}, must start with {self.url_prefix}")
        return event

    def _enter(self, event):
        whole = event.body
        event.body = _extract_input_data(self._input_path, whole)
        return whole, self._pre_handle_event(self.preprocess(event))

    def _leave(self, event, whole, payload):
        event.body = _update_result_body(self._result_path, whole, payload)
        return event

    def _target_segments(self, urlpath):
        """path below the url prefix, split on "/" ([] when nothing follows the prefix)"""
        rest = urlpath[len(self.url_prefix):].strip("/")
        return rest.split("/") if rest else []

    def do_event(self, event, *args, **kwargs):
        whole, event = self._enter(event)
        if not getattr(event, "terminated", None):
            event = self.postprocess(self._handle_event(event))
        return self._leave(event, whole, event.body)


class ModelRouter(BaseModelRouter):
    def _resolve_route(self, body, urlpath):
        """routers.py:168-197 -> (model key, route step | None, operation)"""
        model, op = "", None
        if urlpath and urlpath != "/":
            parts = self._target_segments(urlpath)
            if not parts:
                return "", None, ""
            model = parts[0]
            if len(parts) > 2 and parts[1] == "versions":
                model, parts = f"{model}:{parts[2]}", parts[2:]
            op = "/".join(parts[1:])
        if isinstance(body, dict):
            model = model or body.get("model", list(self.routes.keys())[0])
            op = body.get("operation", op)
        if op is None:
            op = "infer"
        if model not in self.routes:
            raise ValueError(f"model {model} doesnt exist, available models: {' | '.join(self.routes.keys())}")
        return model, self.routes[model], op

    def _handle_event(self, event):
        _name, route, op = self._resolve_route(event.body, event.path)
        if not route:
            event.terminated = True
            event.body = {"models": list(self.routes.keys())}
            return event
        event.path = op
        answer = route.run(event)
        event.body = answer.body if answer else None
        return event


class ParallelRunnerModes(str, Enum):
    array = "array"
    process = "process"
    thread = "thread"

    @staticmethod
    def all():
        return [ParallelRunnerModes.thread, ParallelRunnerModes.process, ParallelRunnerModes.array]


class VotingTypes(str, Enum):
    classification = "classification"
    regression = "regression"


class OperationTypes(str, Enum):
    infer = "infer"
    predict = "predict"
    explain = "explain"


class ParallelRun(BaseModelRouter):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None,
                 health_prefix=None, extend_event=None, executor_type=ParallelRunnerModes.thread, **kwargs):
        super().__init__(context=context, name=name, routes=routes, protocol=protocol,
                         url_prefix=url_prefix, health_prefix=health_prefix, **kwargs)
        self.name = name or "ParallelRun"
        self.extend_event = extend_event
        self.executor_type = ParallelRunnerModes(executor_type)
        self._pool = None

    def merger(self, body, results):
        for part in results.values():
            body.update(part)
        return body

    def _apply_logic(self, results, event=None):
        if not self.extend_event:
            event.body = {}
        return self.merger(event.body, results)

    # ---- fan-out (routers.py:365-455) -------------------------------------------------------------------
    def _init_pool(self):
        if self._pool is None and self.executor_type != ParallelRunnerModes.array:
            self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, len(self.routes)))
        return self._pool

    def _shutdown_pool(self):
        pool, self._pool = self._pool, None
        if pool is not None:
            pool.shutdown()

    @staticmethod
    def _wrap_method(route, handler, event):
        return route, handler(event)

    def _parallel_run(self, event):
        """array: route order; pools: completion order, and a route that raised is simply missing"""
        if self.executor_type == ParallelRunnerModes.array:
            return {key: step.run(copy.copy(event)).body for key, step in self.routes.items()}
        pool = self._init_pool()
        pending = [pool.submit(ParallelRun._wrap_method, key, self.routes[key].run, copy.copy(event)) for key in self.routes.keys()]
        gathered = {}
        for done in concurrent.futures.as_completed(pending):
            try:
                key, answer = done.result()
            except Exception as exc:
                logger.error(traceback.format_exc())
                print(f"child route generated an exception: {exc}")
                continue
            gathered[key] = answer.body
        return gathered

    def do_event(self, event, *args, **kwargs):
        """routers.py:340-363"""
        whole, event = self._enter(event)
        if getattr(event, "terminated", None):
            self._leave(event, whole, event.body)
            self._shutdown_pool()
            return event
        answer = copy.copy(event)
        self._apply_logic(self._parallel_run(event), answer)
        answer = self.postprocess(answer)
        return self._leave(event, whole, answer.body if answer else None)


class VotingEnsemble(ParallelRun):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None,
                 health_prefix=None, vote_type=None, weights=None, executor_type=ParallelRunnerModes.thread,
                 format_response_with_col_name_flag=False, prediction_col_name="prediction", **kwargs):
        super().__init__(context=context, name=name, routes=routes, protocol=protocol, url_prefix=url_prefix,
                         health_prefix=health_prefix, executor_type=executor_type, **kwargs)
        self.name = name or "VotingEnsemble"
        self.vote_type, self.vote_flag = vote_type, vote_type is not None
        self.weights = weights
        tracked = bool(context) and context.stream.enabled
        self._model_logger = _ModelLogPusher(self, context) if tracked else None
        self.version = kwargs.get("version", "v1")
        self.log_router = True
        self.prediction_col_name = prediction_col_name or "prediction"
        self.format_response_with_col_name_flag = format_response_with_col_name_flag
        self.model_endpoint_uid = None

    def post_init(self, mode="sync"):
        ctx = self.context
        if not (getattr(ctx, "_server", None) or getattr(ctx, "server", None)):
            logger.warn("GraphServer not initialized for VotingEnsemble instance")
            return
        self._update_weights(self.weights)

    # ---- weights (routers.py:962-991) ---------------------------------------------------------------------
    def _normalize_weights(self, weights_dict):
        """QUIRK: sums >= ~1 are used as given ([1, 1, 1, 1] turns the mean into a sum); the 'normalise' branch divides a
        0-d object array (`np.array(dict_values)`) and raises TypeError, so weights summing to < 1 never worked"""
        if weights_dict is None:
            share = 1 / len(self.routes)
            return {key: share for key in self.routes.keys()}
        total = np.sum([*weights_dict.values()])
        if 1.0 - total <= 1e-5:
            return weights_dict
        scaled = (np.array(weights_dict.values()) / total).tolist()
        return dict(zip(weights_dict.keys(), scaled))

    def _update_weights(self, weights_dict):
        self._weights = self._normalize_weights(weights_dict)
        for key in self.routes.keys():
            self._weights.setdefault(key, 0)

    # ---- routing (routers.py:623-706) -----------------------------------------------------------------------
    def _resolve_route(self, body, urlpath):
        """-> (model key | ensemble name | "", route step | None, operation)"""
        model, op = "", None
        if urlpath and urlpath != "/":
            parts = self._target_segments(urlpath)
            if not parts:
                return "", None, ""
            if len(parts) == 1 and parts[0] in OperationTypes._value2member_map_:
                self.log_router = True  # a bare operation addresses the ensemble itself
                return self.name, None, OperationTypes(parts[0])
            model = parts[0]
            if len(parts) > 2 and parts[1] == "versions":
                model, parts = f"{parts[0]}:{parts[2]}", parts[2:]
            op = "/".join(parts[1:])
        if isinstance(body, dict):
            model = model or self.name
            op = body.get("operation", op)
        if op is None:
            op = "infer"
        if model in self.routes:
            self.log_router = False  # QUIRK: a direct model call switches router-level tracking off until a bare operation
            return model, self.routes[model], op
        if model != self.name:
            known = " | ".join(self.routes.keys())
            raise ValueError(f"model {model} doesnt exist, available models: "
                             f"{known} | {self.name} or an operation alone for ensemble operation")
        return model, None, op

    # ---- the vote (routers.py:708-810) ---------------------------------------------------------------------
    def _majority_vote(self, all_predictions, weights):
        """one-hot (n, c, m) @ w (m) -> argmax over classes, first max wins"""
        labels = np.array(all_predictions)
        classes = np.arange(labels.max() + 1)
        tally = np.transpose((classes == labels[..., None]).astype(int), (0, 2, 1)) @ weights
        return np.argmax(tally, axis=1).tolist()

    def _mean_vote(self, all_predictions, weights):
        """(n, m) float64 @ w (m)"""
        return (np.array(all_predictions) @ weights).tolist()

    def _is_int(self, value):
        return float(value).is_integer()

    def logic(self, predictions, weights):
        if not self.vote_flag:  # QUIRK: inferred from the FIRST request's values and never revisited
            integral = all(self._is_int(v) for row in predictions for v in row)
            self.vote_type = VotingTypes.classification if integral else VotingTypes.regression
            self.vote_flag = True
        if self.vote_type == VotingTypes.classification:
            return self._majority_vote([[int(v) for v in row] for row in predictions], weights)
        return self._mean_vote(predictions, weights)

    def _apply_logic(self, results, event=None):
        """(m, n) outputs -> (n, m); weights in result-key order"""
        column = self.prediction_col_name if self.format_response_with_col_name_flag else None
        per_model = [r["outputs"][column] if column is not None else r["outputs"] for r in results.values()]
        weights = np.array([self._weights[key] for key in results.keys()])
        return self.logic(np.array(per_model).T, weights)

    # ---- the event handler (routers.py:812-914) ---------------------------------------------------------------
    def validate(self, request, method):
        if self.protocol == "v2" and method != "GET":
            if "inputs" not in request:
                raise Exception('Expected key "inputs" in request body')
            if not isinstance(request["inputs"], list):
                raise Exception('Expected "inputs" to be a list')
        return request

    def _ensemble_metadata(self, event):
        meta = {"name": self.name, "version": self.version or "", "inputs": [], "outputs": []}
        for child in self.routes.values():
            described = child.run(copy.copy(event)).body
            meta["inputs"] = meta["inputs"] or described["inputs"]
            meta["outputs"] = meta["outputs"] or described["outputs"]
            if meta["inputs"] and meta["outputs"]:
                break
        return meta

    def do_event(self, event, *args, **kwargs):
        started = now_date()
        whole, event = self._enter(event)
        if getattr(event, "terminated", None):
            self._leave(event, whole, event.body)
            self._shutdown_pool()
            return event

        target, route, op = self._resolve_route(event.body, event.path)
        event.path = op
        if not target and route is None:
            event.terminated = True
            listing = {"models": list(self.routes.keys()) + [self.name], "weights": self.weights}
            return self._leave(event, whole, listing)

        request = self.validate(event.body, event.method)
        to_ensemble = target == self.name
        if to_ensemble and event.method != "GET":
            votes = self._apply_logic(self._parallel_run(event))
            if self.format_response_with_col_name_flag:
                votes = {self.prediction_col_name: votes}
            answer = copy.copy(event)
            answer.body = {"id": event.id, "model_name": self.name, "outputs": votes}
            if self.version:
                answer.body["model_version"] = self.version
        elif to_ensemble and event.method == "GET" and not op:
            answer = copy.copy(event)
            answer.body = self._ensemble_metadata(event)
        else:
            answer = route.run(event)

        answer = self.postprocess(answer)
        if self._model_logger and self.log_router:
            if "id" not in request:
                request["id"] = answer.body["id"]
            self._model_logger.push(started, request, answer.body)
        return self._leave(event, whole, answer.body if answer else None)
