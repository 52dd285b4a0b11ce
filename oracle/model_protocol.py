"""V2 model-server protocol (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/serving/v2_serving.py:
  V2ModelServer :32-426 (do_event :228-342, validate :362-371, _inputs_to_list :391-426),
  _ModelLogPusher :429-504.
Model-store access (`get_model`) is out of scope: models are in-memory objects or local pickles.

Layout of this restatement: `do_event` resolves (operation, event id) from the path / body, then dispatches to one
handler per operation family; the scoring families (`predict`-like, `explain`) share `_score`, everything that answers
without scoring terminates the event.  The tracking-stream writer keeps the reference's two record layouts
(single event, micro-batch of `values`).
"""

import threading
import time
import traceback
from datetime import datetime, timezone

from .helpers import MLRunInvalidArgumentError, logger  # noqa: F401  (logger re-exported for subclasses)
from .step_io import StepToDict, _extract_input_data, _update_result_body

_SCORING_OPS = frozenset(("predict", "infer", "infer_dict", "predict_dict"))
_LOAD_POLLS, _LOAD_POLL_SECONDS = 50, 5


def now_date():
    return datetime.now(timezone.utc)


def _iso(moment):
    return moment.isoformat(sep=" ", timespec="microseconds")


class V2ModelServer(StepToDict):
    def __init__(self, context=None, name=None, model_path=None, model=None, protocol=None,
                 input_path=None, result_path=None, **kwargs):
        base, _, version = (name or "").partition(":")  # "<model>:<version>" keys come from /versions/<ver>/ URLs
        self.name = base if name else name
        self.version = version
        self.context = context
        self.protocol = protocol or "v2"
        self.model_path, self.model_spec = model_path, None
        self._input_path, self._result_path = input_path, result_path
        self._kwargs = self._params = kwargs
        tracked = bool(context) and context.stream.enabled
        self._model_logger = _ModelLogPusher(self, context) if tracked else None
        self.metrics, self.labels = {}, {}
        self.error = ""
        self.model = model if model else None
        self.ready = bool(model)
        self.model_endpoint_uid = None

    # ---- loading (v2_serving.py:124-154; endpoint records are control plane: skipped) -----------------
    def load(self):
        if not (self.ready or self.model):
            raise ValueError("please specify a load method or a model object")

    def _load_now(self):
        try:
            self.load()
        except Exception as exc:
            self.error = exc
            self.context.logger.error(traceback.format_exc())
            raise RuntimeError(f"failed to load model {self.name}") from exc
        self.ready = True
        self.context.logger.info(f"model {self.name} was loaded")

    def post_init(self, mode="sync"):
        if self.ready:
            return
        if mode != "async":
            self._load_now()
            return
        threading.Thread(target=self._load_now, daemon=True).start()
        self.context.logger.info(f"started async model loading for {self.name}")

    def _await_ready(self, event):
        """:209-219 -- HTTP callers are refused while loading, stream triggers wait"""
        if self.ready:
            return
        if not event.trigger or event.trigger.kind in ("http", ""):
            raise RuntimeError(f"model {self.name} is not ready yet")
        self.context.logger.info(f"waiting for model {self.name} to load")
        for _ in range(_LOAD_POLLS):
            time.sleep(_LOAD_POLL_SECONDS)
            if self.ready:
                return
        raise RuntimeError(f"model {self.name} is not ready {self.error}")

    # ---- user surface ------------------------------------------------------------------------------
    def get_param(self, key, default=None):
        return self._params[key] if key in self._params else self.context.get_param(key, default=default)

    def set_metric(self, name, value):
        self.metrics[name] = value

    def get_model(self, suffix=""):
        """v2_serving.py:166-202 over the local slice of mlrun.artifacts.get_model (artifacts/model.py:434-475): a path that
        ends with the suffix (default ".pkl") is the model file; any other path is listed as a directory and the first
        entry with the suffix is taken; nothing found is a ValueError.  (Store URIs / model-spec yaml: the artifact store.)"""
        from pathlib import Path

        wanted = suffix or ".pkl"
        location = str(self.model_path)
        if location.endswith(wanted):
            return location, {}
        folder = Path(location)
        for entry in (folder.iterdir() if folder.is_dir() else []):
            if entry.name.endswith(wanted):
                return str(entry), {}
        raise ValueError(f"cant resolve model file for {location} suffix{wanted}")

    def preprocess(self, request, operation):
        return request

    def postprocess(self, request):
        return request

    def predict(self, request):
        raise NotImplementedError()

    def explain(self, request):
        raise NotImplementedError()

    def logged_results(self, request, response, op):
        return None, None

    def validate(self, request, operation):
        """:362-371"""
        if self.protocol != "v2":
            return request
        if "inputs" not in request:
            raise Exception('Expected key "inputs" in request body')
        if not isinstance(request["inputs"], list):
            raise Exception('Expected "inputs" to be a list')
        return request

    def _inputs_to_list(self, request):
        """:391-426 -- `*_dict` operations: name -> value dicts become rows in model_spec order"""
        spec = self.model_spec
        if not (spec and spec.inputs):
            raise MLRunInvalidArgumentError(
                "In order to use predict_dict or infer_dict operation you have to provide `model_path` "
                "to the model server and to load it by `load()` function")
        order = [f.name for f in spec.inputs]
        given = request.get("inputs")
        many = isinstance(given, list) and all(isinstance(x, dict) for x in given)
        if not many and not isinstance(given, dict):
            raise MLRunInvalidArgumentError(
                "When using predict_dict or infer_dict operation the inputs must be of type `list[dict]` or `dict`")
        try:
            request["inputs"] = [[row[k] for k in order] for row in given] if many else [given[k] for k in order]
        except KeyError:
            raise MLRunInvalidArgumentError(f"Input dictionary don't contain all the necessary input keys : {order}")
        return request

    # ---- the event handler (:228-342) ----------------------------------------------------------------
    def do_event(self, event, *args, **kwargs):
        started = now_date()
        whole = event.body
        body = _extract_input_data(self._input_path, whole)
        op, event_id = event.path.strip("/"), event.id
        if body and isinstance(body, dict):
            op = op or body.get("operation")
            event_id = body.get("id", event_id)
        if not op and event.method != "GET":
            op = "infer"
        is_get = event.method == "GET"

        if op in _SCORING_OPS or op == "explain":
            response, request = self._score(event, body, op, event_id, started)
            response = self.postprocess(response)
            self._track(started, request, response, op, event_id)
        elif is_get and op == "ready":
            return self._answer_ready(event, event_id)
        elif is_get and op == "":
            response = self._describe(event)
        elif hasattr(self, "op_" + op):
            response = getattr(self, "op_" + op)(event)
        else:
            raise ValueError(f"illegal model operation {op}, method={event.method}")
        event.body = _update_result_body(self._result_path, whole, response)
        return event

    def _score(self, event, body, op, event_id, started):
        self._await_ready(event)
        if "_dict" in op:
            body = self._inputs_to_list(body)
        request = self.validate(self.preprocess(body, op), op)
        run = self.explain if op == "explain" else self.predict
        try:
            outputs = run(request)
        except Exception as exc:
            request["id"] = event_id
            if self._model_logger:
                self._model_logger.push(started, request, op=op, error=exc)
            raise exc
        response = {"id": event_id, "model_name": self.name, "outputs": outputs}
        if op != "explain":
            response["timestamp"] = _iso(started)
        if self.version:
            response["model_version"] = self.version
        return response, request

    def _track(self, started, request, response, op, event_id):
        if not self._model_logger:
            return
        inputs, outputs = self.logged_results(request, response, op)
        if inputs is None and outputs is None:
            self._model_logger.push(started, request, response, op)
        else:
            self._model_logger.push(started, {"id": event_id, "inputs": inputs or []}, {"outputs": outputs or []}, op)

    def _answer_ready(self, event, event_id):
        event.terminated = True
        if self.ready:
            text = f"Model {self.name} is ready (event_id = {event_id})"
            event.body = self.context.Response(status_code=200, body=bytes(text, encoding="utf-8"))
        else:
            event.body = self.context.Response(status_code=408, body=b"model not ready")
        return event

    def _describe(self, event):
        event.terminated = True
        spec = self.model_spec
        return {"name": self.name, "version": self.version,
                "inputs": spec.inputs.to_dict() if spec else [], "outputs": spec.outputs.to_dict() if spec else []}


class _ModelLogPusher:
    """:429-504 -- sampling (`log_stream_sample`), micro-batching (`log_stream_batch`), record layout"""

    _BATCH_HEADERS = ["request", "op", "resp", "when", "microsec", "metrics"]

    def __init__(self, model, context, output_stream=None):
        self.model = model
        self.verbose = context.verbose
        stream = context.stream
        self.hostname, self.function_uri, self.stream_path = stream.hostname, stream.function_uri, stream.stream_uri
        self.stream_batch = int(context.get_param("log_stream_batch", 1))
        self.stream_sample = int(context.get_param("log_stream_sample", 1))
        self.output_stream = output_stream or stream.output_stream
        self._worker = context.worker_id
        self._sample_iter = self._batch_iter = 0
        self._batch = []

    def base_data(self):
        m = self.model
        record = {"class": type(m).__name__, "worker": self._worker, "model": m.name, "version": m.version,
                  "host": self.hostname, "function_uri": self.function_uri}
        if getattr(m, "labels", None):
            record["labels"] = m.labels
        return record

    def _emit(self, **fields):
        record = self.base_data()
        record.update(fields)
        self.output_stream.push([record])

    def push(self, start, request, resp=None, op=None, error=None):
        when = _iso(start)
        if error:
            text = f"{error}\n{traceback.format_exc()}" if self.verbose else str(error)
            self._emit(request=request, op=op, when=when, error=text)
            return
        self._sample_iter = (self._sample_iter + 1) % self.stream_sample
        if self._sample_iter or not self.output_stream:
            return
        microsec = (now_date() - start).microseconds
        if self.stream_batch <= 1:
            extra = {"metrics": self.model.metrics} if getattr(self.model, "metrics", None) else {}
            self._emit(request=request, op=op, resp=resp, when=when, microsec=microsec, **extra)
            return
        if self._batch_iter == 0:
            self._batch = []
        self._batch.append([request, op, resp, str(start), microsec, self.model.metrics])
        self._batch_iter = (self._batch_iter + 1) % self.stream_batch
        if self._batch_iter == 0:
            self._emit(headers=list(self._BATCH_HEADERS), values=self._batch)
