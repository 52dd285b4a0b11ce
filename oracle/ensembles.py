"""Model routers and the voting ensemble (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/serving/routers.py:
  BaseModelRouter :43-164, ModelRouter :167-211, ParallelRun :245-477,
  VotingEnsemble :480-991 (_resolve_route :623-706, _majority_vote :708-730, _mean_vote :732-741,
  logic :746-787, _apply_logic :789-810, do_event :812-914, _normalize_weights :962-991).
Process-pool execution (routers.py:378-396) is out of scope for the oracle; "process" falls back to
the thread pool (same results, same key-order caveat).
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
        self.name = name
        self.context = context
        self.routes = routes
        self.protocol = protocol or "v2"
        self.url_prefix = url_prefix or f"/{self.protocol}/models"
        self.health_prefix = health_prefix or f"/{self.protocol}/health"
        self.inputs_key = "instances" if self.protocol == "v1" else "inputs"
        self._input_path = input_path
        self._result_path = result_path
        self.kwargs = kwargs

    def parse_event(self, event):
        """routers.py:86-112 (data_url download is I/O: out of scope)"""
        parsed = {}
        try:
            body = event.body if isinstance(event.body, dict) else json.loads(event.body)
            parsed = body
        except Exception as exc:
            content_type = getattr(event, "content_type", "") or ""
            if content_type.startswith("image/"):
                parsed[self.inputs_key] = [BytesIO(event.body)]
            else:
                raise ValueError("Unrecognized request format") from exc
        return parsed

    def post_init(self, mode="sync"):
        self.context.logger.info(f"Loaded {list(self.routes.keys())}")

    def get_metadata(self):
        return {"name": self.__class__.__name__, "version": "v2", "extensions": []}

    def _pre_handle_event(self, event):
        """routers.py:122-141"""
        method = event.method or "POST"
        if event.body and method != "GET":
            event.body = self.parse_event(event)
        urlpath = getattr(event, "path", "")
        if method == "GET" and (urlpath == "/" or urlpath.startswith(self.health_prefix)):
            setattr(event, "terminated", True)
            event.body = self.get_metadata()
            return event
        if urlpath and not urlpath.startswith(self.url_prefix) and not urlpath == "/":
            raise ValueError(f"illegal path prefix {urlpath}, must start with {self.url_prefix}")
        return event

    def do_event(self, event, *args, **kwargs):
        original_body = event.body
        event.body = _extract_input_data(self._input_path, event.body)
        event = self.preprocess(event)
        event = self._pre_handle_event(event)
        if not getattr(event, "terminated", None):
            event = self.postprocess(self._handle_event(event))
        event.body = _update_result_body(self._result_path, original_body, event.body)
        return event

    def _handle_event(self, event):
        return event

    def preprocess(self, event):
        return event

    def postprocess(self, event):
        return event


class ModelRouter(BaseModelRouter):
    def _resolve_route(self, body, urlpath):
        """routers.py:168-197"""
        subpath = None
        model = ""
        if urlpath and not urlpath == "/":
            subpath = ""
            urlpath = urlpath[len(self.url_prefix):].strip("/")
            if not urlpath:
                return "", None, ""
            segments = urlpath.split("/")
            model = segments[0]
            if len(segments) > 2 and segments[1] == "versions":
                model = model + ":" + segments[2]
                segments = segments[2:]
            if len(segments) > 1:
                subpath = "/".join(segments[1:])
        if isinstance(body, dict):
            model = model or body.get("model", list(self.routes.keys())[0])
            subpath = body.get("operation", subpath)
        if subpath is None:
            subpath = "infer"
        if model not in self.routes:
            models = " | ".join(self.routes.keys())
            raise ValueError(f"model {model} doesnt exist, available models: {models}")
        return model, self.routes[model], subpath

    def _handle_event(self, event):
        name, route, subpath = self._resolve_route(event.body, event.path)
        if not route:
            setattr(event, "terminated", True)
            event.body = {"models": list(self.routes.keys())}
            return event
        event.path = subpath
        response = route.run(event)
        event.body = response.body if response else None
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

    def _apply_logic(self, results, event=None):
        if not self.extend_event:
            event.body = {}
        return self.merger(event.body, results)

    def merger(self, body, results):
        for result in results.values():
            body.update(result)
        return body

    def do_event(self, event, *args, **kwargs):
        """routers.py:340-363"""
        original_body = event.body
        event.body = _extract_input_data(self._input_path, event.body)
        event = self.preprocess(event)
        event = self._pre_handle_event(event)
        if getattr(event, "terminated", None):
            event.body = _update_result_body(self._result_path, original_body, event.body)
            self._shutdown_pool()
            return event
        response = copy.copy(event)
        results = self._parallel_run(event)
        self._apply_logic(results, response)
        response = self.postprocess(response)
        event.body = _update_result_body(self._result_path, original_body, response.body if response else None)
        return event

    def _init_pool(self):
        if self._pool is None and self.executor_type != ParallelRunnerModes.array:
            self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, len(self.routes)))
        return self._pool

    def _shutdown_pool(self):
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None

    def _parallel_run(self, event):
        """routers.py:414-455: array = in route order; pools = completion order, raising routes dropped"""
        if self.executor_type == ParallelRunnerModes.array:
            return {name: step.run(copy.copy(event)).body for name, step in self.routes.items()}
        futures = []
        executor = self._init_pool()
        for route in self.routes.keys():
            step = self.routes[route]
            futures.append(executor.submit(ParallelRun._wrap_method, route, step.run, copy.copy(event)))
        results = {}
        for future in concurrent.futures.as_completed(futures):
            try:
                key, result = future.result()
                results[key] = result.body
            except Exception as exc:
                logger.error(traceback.format_exc())
                print(f"child route generated an exception: {exc}")
        return results

    @staticmethod
    def _wrap_method(route, handler, event):
        return route, handler(event)


class VotingEnsemble(ParallelRun):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None,
                 health_prefix=None, vote_type=None, weights=None, executor_type=ParallelRunnerModes.thread,
                 format_response_with_col_name_flag=False, prediction_col_name="prediction", **kwargs):
        super().__init__(context=context, name=name, routes=routes, protocol=protocol, url_prefix=url_prefix,
                         health_prefix=health_prefix, executor_type=executor_type, **kwargs)
        self.name = name or "VotingEnsemble"
        self.vote_type = vote_type
        self.vote_flag = self.vote_type is not None
        self.weights = weights
        self._model_logger = _ModelLogPusher(self, context) if context and context.stream.enabled else None
        self.version = kwargs.get("version", "v1")
        self.log_router = True
        self.prediction_col_name = prediction_col_name or "prediction"
        self.format_response_with_col_name_flag = format_response_with_col_name_flag
        self.model_endpoint_uid = None

    def post_init(self, mode="sync"):
        server = getattr(self.context, "_server", None) or getattr(self.context, "server", None)
        if not server:
            logger.warn("GraphServer not initialized for VotingEnsemble instance")
            return
        self._update_weights(self.weights)

    def _resolve_route(self, body, urlpath):
        """routers.py:623-706"""
        subpath = None
        model = ""
        if urlpath and not urlpath == "/":
            subpath = ""
            urlpath = urlpath[len(self.url_prefix):].strip("/")
            if not urlpath:
                return "", None, ""
            segments = urlpath.split("/")
            if len(segments) == 1:
                try:
                    operation = OperationTypes(segments[0])
                except ValueError:
                    model = segments[0]
                else:
                    self.log_router = True
                    return self.name, None, operation
            if len(segments) > 2 and segments[1] == "versions":
                model = f"{segments[0]}:{segments[2]}"
                segments = segments[2:]
            else:
                model = segments[0]
            if len(segments) > 1:
                subpath = "/".join(segments[1:])
        if isinstance(body, dict):
            model = model or self.name
            subpath = body.get("operation", subpath)
        if subpath is None:
            subpath = "infer"
        if model in self.routes:
            self.log_router = False
            return model, self.routes[model], subpath
        elif model != self.name:
            models = " | ".join(self.routes.keys())
            raise ValueError(
                f"model {model} doesnt exist, available models: "
                f"{models} | {self.name} or an operation alone for ensemble operation"
            )
        return model, None, subpath

    def _majority_vote(self, all_predictions, weights):
        """one-hot (n,c,m) @ w(m) -> argmax over classes, first max wins (routers.py:708-730)"""
        preds = np.array(all_predictions)
        one_hot = np.transpose(
            (np.arange(preds.max() + 1) == preds[..., None]).astype(int), (0, 2, 1)
        )
        weighted = one_hot @ weights
        return np.argmax(weighted, axis=1).tolist()

    def _mean_vote(self, all_predictions, weights):
        """(n,m) float64 @ w(m) (routers.py:732-741)"""
        return (np.array(all_predictions) @ weights).tolist()

    def _is_int(self, value):
        return float(value).is_integer()

    def logic(self, predictions, weights):
        """vote-type inference happens once and sticks (routers.py:746-787)"""
        if not self.vote_flag:
            if all(all(map(self._is_int, row)) for row in predictions):
                self.vote_type = VotingTypes.classification
            else:
                self.vote_type = VotingTypes.regression
            self.vote_flag = True
        if self.vote_type == VotingTypes.classification:
            int_predictions = [list(map(int, row)) for row in predictions]
            return self._majority_vote(int_predictions, weights)
        return self._mean_vote(predictions, weights)

    def _apply_logic(self, results, event=None):
        """(m,n) outputs -> (n,m); weights in result-key order (routers.py:789-810)"""
        flat = np.array(
            [
                (r["outputs"][self.prediction_col_name] if self.format_response_with_col_name_flag else r["outputs"])
                for r in results.values()
            ]
        ).T
        weights = [self._weights[name] for name in results.keys()]
        return self.logic(flat, np.array(weights))

    def do_event(self, event, *args, **kwargs):
        """routers.py:812-914"""
        start = now_date()
        original_body = event.body
        event.body = _extract_input_data(self._input_path, event.body)
        event = self.preprocess(event)
        event = self._pre_handle_event(event)
        if getattr(event, "terminated", None):
            event.body = _update_result_body(self._result_path, original_body, event.body)
            self._shutdown_pool()
            return event

        name, route, subpath = self._resolve_route(event.body, event.path)
        event.path = subpath

        if not name and route is None:
            setattr(event, "terminated", True)
            event.body = {"models": list(self.routes.keys()) + [self.name], "weights": self.weights}
            event.body = _update_result_body(self._result_path, original_body, event.body)
            return event

        request = self.validate(event.body, event.method)
        if name == self.name and event.method != "GET":
            predictions = self._parallel_run(event)
            votes = self._apply_logic(predictions)
            if self.format_response_with_col_name_flag:
                votes = {self.prediction_col_name: votes}
            response = copy.copy(event)
            body = {"id": event.id, "model_name": self.name, "outputs": votes}
            if self.version:
                body["model_version"] = self.version
            response.body = body
        elif name == self.name and event.method == "GET" and not subpath:
            response = copy.copy(event)
            body = {"name": self.name, "version": self.version or "", "inputs": [], "outputs": []}
            for child in self.routes.values():
                child_resp = child.run(copy.copy(event))
                body["inputs"] = body["inputs"] or child_resp.body["inputs"]
                body["outputs"] = body["outputs"] or child_resp.body["outputs"]
                if body["inputs"] and body["outputs"]:
                    break
            response.body = body
        else:
            response = route.run(event)

        response = self.postprocess(response)
        if self._model_logger and self.log_router:
            if "id" not in request:
                request["id"] = response.body["id"]
            self._model_logger.push(start, request, response.body)
        event.body = _update_result_body(self._result_path, original_body, response.body if response else None)
        return event

    def validate(self, request, method):
        if self.protocol == "v2" and method != "GET":
            if "inputs" not in request:
                raise Exception('Expected key "inputs" in request body')
            if not isinstance(request["inputs"], list):
                raise Exception('Expected "inputs" to be a list')
        return request

    def _normalize_weights(self, weights_dict):
        """routers.py:962-980 -- including its quirk: sums >= ~1 are returned as given and the
        'normalise' branch divides a 0-d object array (dict_values) and raises TypeError"""
        if weights_dict is None:
            n = len(self.routes)
            return dict(zip(self.routes.keys(), [1 / n] * n))
        values = [*weights_dict.values()]
        total = np.sum(values)
        if 1.0 - total <= 1e-5:
            return weights_dict
        new_values = (np.array(weights_dict.values()) / total).tolist()
        return dict(zip(weights_dict.keys(), new_values))

    def _update_weights(self, weights_dict):
        self._weights = self._normalize_weights(weights_dict)
        for model in self.routes.keys():
            if model not in self._weights.keys():
                self._weights[model] = 0
