"""Model routers and the voting ensemble (plugin-API mirror of mlrun/serving/routers.py:43-991).

URL / body routing, the V2 router protocol and the vote arithmetic keep the reference's observable
behaviour (including its quirks: the one-time vote-type inference, weights that are never actually
normalised, routes that raised being dropped from the vote).  What differs is where the arithmetic
runs: when every route of a VotingEnsemble is a device model server the graph compiler fuses
fan-out + per-model predict + vote into one CUDA plan (see compiler.py); the per-event code below is
the plugin-compatibility path for arbitrary Python model classes.
"""

import concurrent.futures
import copy
import json
import traceback
from enum import Enum
from io import BytesIO

import numpy as np

from .model_server import _ModelLogPusher, now_date
from .paths import merge_result, select_input
from .resolve import logger
from .step_meta import StepMeta


class _RouterMeta(StepMeta):
    _STEP_KIND = "router"
    _dict_exclude = ("routes",)


class BaseModelRouter(_RouterMeta):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None, health_prefix=None,
                 input_path=None, result_path=None, **kwargs):
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

    # hooks
    def preprocess(self, event):
        return event

    def postprocess(self, event):
        return event

    def post_init(self, mode="sync"):
        self.context.logger.info(f"Loaded {list(self.routes.keys())}")

    def get_metadata(self):
        return {"name": type(self).__name__, "version": "v2", "extensions": []}

    def parse_event(self, event):
        try:
            return event.body if isinstance(event.body, dict) else json.loads(event.body)
        except Exception as exc:
            if (getattr(event, "content_type", "") or "").startswith("image/"):
                return {self.inputs_key: [BytesIO(event.body)]}
            raise ValueError("Unrecognized request format") from exc

    def _pre_handle_event(self, event):
        method = event.method or "POST"
        if event.body and method != "GET":
            event.body = self.parse_event(event)
        path = getattr(event, "path", "")
        if method == "GET" and (path == "/" or path.startswith(self.health_prefix)):
            event.terminated = True
            event.body = self.get_metadata()
        elif path and path != "/" and not path.startswith(self.url_prefix):
            raise ValueError(f"illegal path prefix {path}, must start with {self.url_prefix}")
        return event

    def _enter(self, event):
        """shared prologue: input_path, preprocess hook, health / prefix handling"""
        original = event.body
        event.body = select_input(self._input_path, event.body)
        event = self._pre_handle_event(self.preprocess(event))
        return original, event

    def do_event(self, event, *args, **kwargs):
        original, event = self._enter(event)
        if not getattr(event, "terminated", None):
            event = self.postprocess(self._handle_event(event))
        event.body = merge_result(self._result_path, original, event.body)
        return event

    def _handle_event(self, event):
        return event

    def _parse_url(self, urlpath):
        """`<prefix>/<model>[/versions/<ver>][/<operation...>]` -> (model, operation) as the routers read a URL
        (serving/routers.py:143-181): None when the event has no routing URL ("" or "/"), ("", "") for the bare prefix;
        a version becomes part of the model key ("name:ver"); the operation is "" when the URL stops at the model."""
        if not urlpath or urlpath == "/":
            return None
        pieces = [piece for piece in urlpath[len(self.url_prefix):].strip("/").split("/")]
        if pieces == [""]:
            return "", ""
        model, operation = pieces[0], pieces[1:]
        if len(pieces) > 2 and pieces[1] == "versions":
            model, operation = f"{pieces[0]}:{pieces[2]}", pieces[3:]
        return model, "/".join(operation)

    @staticmethod
    def _operation_of(body, from_url):
        """the body of a stream event may name the operation (routers.py:183-190); "infer" when nobody does"""
        op = body.get("operation", from_url) if isinstance(body, dict) else from_url
        return "infer" if op is None else op


class ModelRouter(BaseModelRouter):
    def _resolve_route(self, body, urlpath):
        """-> (model key, route, operation); ("", None, "") for the bare prefix (the caller lists the models)"""
        target = self._parse_url(urlpath)
        if target == ("", ""):
            return "", None, ""
        model, url_op = target if target else ("", None)
        if not model and isinstance(body, dict):  # stream events carry the route in the body; the first model is the default
            model = body.get("model", next(iter(self.routes)))
        operation = self._operation_of(body, url_op)
        if model not in self.routes:
            raise ValueError(f"model {model} doesnt exist, available models: {' | '.join(self.routes.keys())}")
        return model, self.routes[model], operation

    def _handle_event(self, event):
        name, route, subpath = self._resolve_route(event.body, event.path)
        if not route:
            event.terminated = True
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
    """fan the event out to every route and merge the route results (routers.py:245-477).  `process`
    pools are served by threads here: the results are the same and nothing on this path benefits from
    forking once the arithmetic lives on the GPU"""

    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None, health_prefix=None,
                 extend_event=None, executor_type=ParallelRunnerModes.thread, **kwargs):
        super().__init__(context=context, name=name, routes=routes, protocol=protocol, url_prefix=url_prefix,
                         health_prefix=health_prefix, **kwargs)
        self.name = name or "ParallelRun"
        self.extend_event = extend_event
        self.executor_type = ParallelRunnerModes(executor_type)
        self._pool = None

    def merger(self, body, results):
        for result in results.values():
            body.update(result)
        return body

    def _apply_logic(self, results, event=None):
        if not self.extend_event:
            event.body = {}
        return self.merger(event.body, results)

    def _shutdown_pool(self):
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None

    def _parallel_run(self, event):
        if self.executor_type == ParallelRunnerModes.array:
            return {name: step.run(copy.copy(event)).body for name, step in self.routes.items()}
        if self._pool is None:
            self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, len(self.routes)))
        jobs = {self._pool.submit(step.run, copy.copy(event)): name for name, step in self.routes.items()}
        results = {}
        for job in concurrent.futures.as_completed(jobs):
            try:
                results[jobs[job]] = job.result().body
            except Exception as exc:  # a failing route is dropped from the result set
                logger.error(traceback.format_exc())
                print(f"child route generated an exception: {exc}")
        return results

    def do_event(self, event, *args, **kwargs):
        original, event = self._enter(event)
        if getattr(event, "terminated", None):
            event.body = merge_result(self._result_path, original, event.body)
            self._shutdown_pool()
            return event
        response = copy.copy(event)
        self._apply_logic(self._parallel_run(event), response)
        response = self.postprocess(response)
        event.body = merge_result(self._result_path, original, response.body if response else None)
        return event


class VotingEnsemble(ParallelRun):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None, health_prefix=None,
                 vote_type=None, weights=None, executor_type=ParallelRunnerModes.thread,
                 format_response_with_col_name_flag=False, prediction_col_name="prediction", **kwargs):
        super().__init__(context=context, name=name, routes=routes, protocol=protocol, url_prefix=url_prefix,
                         health_prefix=health_prefix, executor_type=executor_type, **kwargs)
        self.name = name or "VotingEnsemble"
        self.vote_type = vote_type
        self.vote_flag = vote_type is not None
        self.weights = weights
        self._model_logger = _ModelLogPusher(self, context) if context is not None and context.stream.enabled else None
        self.version = kwargs.get("version", "v1")
        self.log_router = True
        self.prediction_col_name = prediction_col_name or "prediction"
        self.format_response_with_col_name_flag = format_response_with_col_name_flag
        self.model_endpoint_uid = None

    def post_init(self, mode="sync"):
        if not (getattr(self.context, "_server", None) or getattr(self.context, "server", None)):
            logger.warn("GraphServer not initialized for VotingEnsemble instance")
            return
        self._update_weights(self.weights)

    # ---- weights (routers.py:962-991, quirks included) -----------------------------------------------
    def _normalize_weights(self, weights_dict):
        if weights_dict is None:
            n = len(self.routes)
            return dict(zip(self.routes.keys(), [1 / n] * n))
        if 1.0 - np.sum([*weights_dict.values()]) <= 1e-5:
            return weights_dict  # sums >= ~1 are used as given ([1,1,1,1] turns the mean into a sum)
        # the reference divides `np.array(dict_values)` (a 0-d object array) here and raises TypeError:
        # weights summing to < 1 never worked; keep failing the same way instead of inventing semantics
        return dict(zip(weights_dict.keys(), (np.array(weights_dict.values()) / np.sum([*weights_dict.values()])).tolist()))

    def _update_weights(self, weights_dict):
        self._weights = self._normalize_weights(weights_dict)
        for model in self.routes.keys():
            self._weights.setdefault(model, 0)

    # ---- routing (routers.py:623-706) -----------------------------------------------------------------
    def _resolve_route(self, body, urlpath):
        """-> (name, route, operation).  route None + name == self.name: the ensemble itself answers (vote over all routes);
        a URL that is `<prefix>/<operation>` alone addresses the ensemble too (routers.py:640-706)"""
        target = self._parse_url(urlpath)
        if target == ("", ""):
            return "", None, ""
        model, url_op = target if target else ("", None)
        if target and "/" not in urlpath[len(self.url_prefix):].strip("/"):  # one segment: an operation of the router, or a model
            if model in OperationTypes._value2member_map_:
                self.log_router = True
                return self.name, None, OperationTypes(model)
        if not model and isinstance(body, dict):
            model = self.name
        operation = self._operation_of(body, url_op)
        if model in self.routes:
            self.log_router = False  # plain pass-through to one model: its own server logs the event
            return model, self.routes[model], operation
        if model == self.name:
            return model, None, operation
        raise ValueError(
            f"model {model} doesnt exist, available models: "
            f"{' | '.join(self.routes.keys())} | {self.name} or an operation alone for ensemble operation")

    # ---- vote arithmetic (routers.py:708-810) -- the CPU form of what vote_and_store does on the device ---
    def _majority_vote(self, all_predictions, weights):
        preds = np.array(all_predictions)
        one_hot = np.transpose((np.arange(preds.max() + 1) == preds[..., None]).astype(int), (0, 2, 1))
        return np.argmax(one_hot @ weights, axis=1).tolist()

    def _mean_vote(self, all_predictions, weights):
        return (np.array(all_predictions) @ weights).tolist()

    def _is_int(self, value):
        return float(value).is_integer()

    def logic(self, predictions, weights):
        if not self.vote_flag:  # inferred once, from the first request, then it sticks
            all_int = all(all(self._is_int(v) for v in row) for row in predictions)
            self.vote_type = VotingTypes.classification if all_int else VotingTypes.regression
            self.vote_flag = True
        if self.vote_type == VotingTypes.classification:
            return self._majority_vote([[int(v) for v in row] for row in predictions], weights)
        return self._mean_vote(predictions, weights)

    def _apply_logic(self, results, event=None):
        col = self.prediction_col_name
        per_model = [(r["outputs"][col] if self.format_response_with_col_name_flag else r["outputs"]) for r in results.values()]
        weights = np.array([self._weights[name] for name in results.keys()])
        return self.logic(np.array(per_model).T, weights)

    def validate(self, request, method):
        if self.protocol == "v2" and method != "GET":
            if "inputs" not in request:
                raise Exception('Expected key "inputs" in request body')
            if not isinstance(request["inputs"], list):
                raise Exception('Expected "inputs" to be a list')
        return request

    def do_event(self, event, *args, **kwargs):
        start = now_date()
        original, event = self._enter(event)
        if getattr(event, "terminated", None):
            event.body = merge_result(self._result_path, original, event.body)
            self._shutdown_pool()
            return event
        name, route, subpath = self._resolve_route(event.body, event.path)
        event.path = subpath
        if not name and route is None:
            event.terminated = True
            listing = {"models": list(self.routes.keys()) + [self.name], "weights": self.weights}
            event.body = merge_result(self._result_path, original, listing)
            return event

        request = self.validate(event.body, event.method)
        if name == self.name and event.method != "GET":
            votes = self._apply_logic(self._parallel_run(event))
            if self.format_response_with_col_name_flag:
                votes = {self.prediction_col_name: votes}
            response = copy.copy(event)
            response.body = {"id": event.id, "model_name": self.name, "outputs": votes}
            if self.version:
                response.body["model_version"] = self.version
        elif name == self.name and event.method == "GET" and not subpath:
            response = copy.copy(event)
            meta = {"name": self.name, "version": self.version or "", "inputs": [], "outputs": []}
            for child in self.routes.values():
                got = child.run(copy.copy(event)).body
                meta["inputs"] = meta["inputs"] or got["inputs"]
                meta["outputs"] = meta["outputs"] or got["outputs"]
                if meta["inputs"] and meta["outputs"]:
                    break
            response.body = meta
        else:
            response = route.run(event)

        response = self.postprocess(response)
        if self._model_logger and self.log_router:
            request.setdefault("id", response.body["id"])
            self._model_logger.push(start, request, response.body)
        event.body = merge_result(self._result_path, original, response.body if response else None)
        return event


class _EnrichmentMixin:
    """EnrichmentModelRouter / EnrichmentVotingEnsemble (routers.py:1118-1196, 1199-1342): entity keys in, feature
    vectors to the child models; the online read is one device launch (mlrun_b200.feature_store.online)"""

    def _init_enrichment(self, feature_vector_uri, impute_policy):
        self.feature_vector_uri = feature_vector_uri
        self.impute_policy = impute_policy or {}
        self._feature_service = None

    def post_init(self, mode="sync"):
        from ..feature_store.online import get_feature_vector

        super().post_init(mode)
        self._feature_service = get_feature_vector(self.feature_vector_uri).get_online_feature_service(
            impute_policy=self.impute_policy)

    def preprocess(self, event):
        if isinstance(event.body, (str, bytes)):
            event.body = json.loads(event.body)
        event.body["inputs"] = self._feature_service.get(event.body["inputs"], as_list=True)
        return event


class EnrichmentModelRouter(_EnrichmentMixin, ModelRouter):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None, health_prefix=None,
                 feature_vector_uri="", impute_policy=None, **kwargs):
        super().__init__(context, name, routes, protocol, url_prefix, health_prefix, **kwargs)
        self._init_enrichment(feature_vector_uri, impute_policy)


class EnrichmentVotingEnsemble(_EnrichmentMixin, VotingEnsemble):
    def __init__(self, context=None, name=None, routes=None, protocol=None, url_prefix=None, health_prefix=None,
                 vote_type=None, executor_type=ParallelRunnerModes.thread, prediction_col_name=None, feature_vector_uri="",
                 impute_policy=None, **kwargs):
        super().__init__(context=context, name=name, routes=routes, protocol=protocol, url_prefix=url_prefix,
                         health_prefix=health_prefix, vote_type=vote_type, executor_type=executor_type,
                         prediction_col_name=prediction_col_name, **kwargs)
        self._init_enrichment(feature_vector_uri, impute_policy)
