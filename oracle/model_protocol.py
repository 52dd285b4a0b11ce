"""V2 model-server protocol (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/serving/v2_serving.py:
  V2ModelServer :32-426 (do_event :228-342, validate :362-371, _inputs_to_list :391-426),
  _ModelLogPusher :429-504.
Model-store access (`get_model`) is out of scope: models are in-memory objects or local pickles.
"""

import threading
import time
import traceback
from datetime import datetime, timezone

from .helpers import MLRunInvalidArgumentError, logger
from .step_io import StepToDict, _extract_input_data, _update_result_body


def now_date():
    return datetime.now(timezone.utc)


class V2ModelServer(StepToDict):
    def __init__(self, context=None, name=None, model_path=None, model=None, protocol=None,
                 input_path=None, result_path=None, **kwargs):
        self.name = name
        self.version = ""
        if name and ":" in name:
            self.name, self.version = name.split(":", 1)
        self.context = context
        self.ready = False
        self.error = ""
        self.protocol = protocol or "v2"
        self.model_path = model_path
        self.model_spec = None
        self._input_path = input_path
        self._result_path = result_path
        self._kwargs = kwargs
        self._params = kwargs
        self._model_logger = (
            _ModelLogPusher(self, context) if context and context.stream.enabled else None
        )
        self.metrics = {}
        self.labels = {}
        self.model = None
        if model:
            self.model = model
            self.ready = True
        self.model_endpoint_uid = None

    def _load_and_update_state(self):
        try:
            self.load()
        except Exception as exc:
            self.error = exc
            self.context.logger.error(traceback.format_exc())
            raise RuntimeError(f"failed to load model {self.name}") from exc
        self.ready = True
        self.context.logger.info(f"model {self.name} was loaded")

    def post_init(self, mode="sync"):
        """v2_serving.py:134-154 (endpoint records are control-plane: skipped)"""
        if not self.ready:
            if mode == "async":
                threading.Thread(target=self._load_and_update_state, daemon=True).start()
                self.context.logger.info(f"started async model loading for {self.name}")
            else:
                self._load_and_update_state()

    def get_param(self, key, default=None):
        if key in self._params:
            return self._params.get(key)
        return self.context.get_param(key, default=default)

    def set_metric(self, name, value):
        self.metrics[name] = value

    def get_model(self, suffix=""):
        """local-file stand-in for mlrun.artifacts.get_model (v2_serving.py:166-202)"""
        return self.model_path, {}

    def load(self):
        if not self.ready and not self.model:
            raise ValueError("please specify a load method or a model object")

    def _check_readiness(self, event):
        """v2_serving.py:209-219"""
        if self.ready:
            return
        if not event.trigger or event.trigger.kind in ["http", ""]:
            raise RuntimeError(f"model {self.name} is not ready yet")
        self.context.logger.info(f"waiting for model {self.name} to load")
        for _ in range(50):
            time.sleep(5)
            if self.ready:
                return
        raise RuntimeError(f"model {self.name} is not ready {self.error}")

    def _pre_event_processing_actions(self, event, event_body, op):
        self._check_readiness(event)
        if "_dict" in op:
            event_body = self._inputs_to_list(event_body)
        request = self.preprocess(event_body, op)
        return self.validate(request, op)

    def do_event(self, event, *args, **kwargs):
        """v2_serving.py:228-342"""
        start = now_date()
        original_body = event.body
        event_body = _extract_input_data(self._input_path, event.body)
        event_id = event.id
        op = event.path.strip("/")
        if event_body and isinstance(event_body, dict):
            op = op or event_body.get("operation")
            event_id = event_body.get("id", event_id)
        if not op and event.method != "GET":
            op = "infer"

        if op in ("predict", "infer", "infer_dict", "predict_dict"):
            request = self._pre_event_processing_actions(event, event_body, op)
            try:
                outputs = self.predict(request)
            except Exception as exc:
                request["id"] = event_id
                if self._model_logger:
                    self._model_logger.push(start, request, op=op, error=exc)
                raise exc
            response = {
                "id": event_id,
                "model_name": self.name,
                "outputs": outputs,
                "timestamp": start.isoformat(sep=" ", timespec="microseconds"),
            }
            if self.version:
                response["model_version"] = self.version

        elif op == "ready" and event.method == "GET":
            setattr(event, "terminated", True)
            if self.ready:
                event.body = self.context.Response(
                    status_code=200,
                    body=bytes(f"Model {self.name} is ready (event_id = {event_id})", encoding="utf-8"),
                )
            else:
                event.body = self.context.Response(status_code=408, body=b"model not ready")
            return event

        elif op == "" and event.method == "GET":
            setattr(event, "terminated", True)
            meta = {"name": self.name, "version": self.version, "inputs": [], "outputs": []}
            if self.model_spec:
                meta["inputs"] = self.model_spec.inputs.to_dict()
                meta["outputs"] = self.model_spec.outputs.to_dict()
            event.body = _update_result_body(self._result_path, original_body, meta)
            return event

        elif op == "explain":
            request = self._pre_event_processing_actions(event, event_body, op)
            try:
                outputs = self.explain(request)
            except Exception as exc:
                request["id"] = event_id
                if self._model_logger:
                    self._model_logger.push(start, request, op=op, error=exc)
                raise exc
            response = {"id": event_id, "model_name": self.name, "outputs": outputs}
            if self.version:
                response["model_version"] = self.version

        elif hasattr(self, "op_" + op):
            response = getattr(self, "op_" + op)(event)
            event.body = _update_result_body(self._result_path, original_body, response)
            return event

        else:
            raise ValueError(f"illegal model operation {op}, method={event.method}")

        response = self.postprocess(response)
        if self._model_logger:
            inputs, outputs = self.logged_results(request, response, op)
            if inputs is None and outputs is None:
                self._model_logger.push(start, request, response, op)
            else:
                track_request = {"id": event_id, "inputs": inputs or []}
                track_response = {"outputs": outputs or []}
                self._model_logger.push(start, track_request, track_response, op)
        event.body = _update_result_body(self._result_path, original_body, response)
        return event

    def logged_results(self, request, response, op):
        return None, None

    def validate(self, request, operation):
        """v2_serving.py:362-371"""
        if self.protocol == "v2":
            if "inputs" not in request:
                raise Exception('Expected key "inputs" in request body')
            if not isinstance(request["inputs"], list):
                raise Exception('Expected "inputs" to be a list')
        return request

    def preprocess(self, request, operation):
        return request

    def postprocess(self, request):
        return request

    def predict(self, request):
        raise NotImplementedError()

    def explain(self, request):
        raise NotImplementedError()

    def _inputs_to_list(self, request):
        """v2_serving.py:391-426"""
        if self.model_spec and self.model_spec.inputs:
            order = [feature.name for feature in self.model_spec.inputs]
        else:
            raise MLRunInvalidArgumentError(
                "In order to use predict_dict or infer_dict operation you have to provide `model_path` "
                "to the model server and to load it by `load()` function"
            )
        inputs = request.get("inputs")
        try:
            if isinstance(inputs, list) and all(isinstance(item, dict) for item in inputs):
                new_inputs = [[d[key] for key in order] for d in inputs]
            elif isinstance(inputs, dict):
                new_inputs = [inputs[key] for key in order]
            else:
                raise MLRunInvalidArgumentError(
                    "When using predict_dict or infer_dict operation the inputs must be "
                    "of type `list[dict]` or `dict`"
                )
        except KeyError:
            raise MLRunInvalidArgumentError(
                f"Input dictionary don't contain all the necessary input keys : {order}"
            )
        request["inputs"] = new_inputs
        return request


class _ModelLogPusher:
    """v2_serving.py:429-504"""

    def __init__(self, model, context, output_stream=None):
        self.model = model
        self.verbose = context.verbose
        self.hostname = context.stream.hostname
        self.function_uri = context.stream.function_uri
        self.stream_path = context.stream.stream_uri
        self.stream_batch = int(context.get_param("log_stream_batch", 1))
        self.stream_sample = int(context.get_param("log_stream_sample", 1))
        self.output_stream = output_stream or context.stream.output_stream
        self._worker = context.worker_id
        self._sample_iter = 0
        self._batch_iter = 0
        self._batch = []

    def base_data(self):
        data = {
            "class": self.model.__class__.__name__,
            "worker": self._worker,
            "model": self.model.name,
            "version": self.model.version,
            "host": self.hostname,
            "function_uri": self.function_uri,
        }
        if getattr(self.model, "labels", None):
            data["labels"] = self.model.labels
        return data

    def push(self, start, request, resp=None, op=None, error=None):
        start_str = start.isoformat(sep=" ", timespec="microseconds")
        if error:
            data = self.base_data()
            data["request"] = request
            data["op"] = op
            data["when"] = start_str
            message = str(error)
            if self.verbose:
                message = f"{message}\n{traceback.format_exc()}"
            data["error"] = message
            self.output_stream.push([data])
            return

        self._sample_iter = (self._sample_iter + 1) % self.stream_sample
        if self.output_stream and self._sample_iter == 0:
            microsec = (now_date() - start).microseconds
            if self.stream_batch > 1:
                if self._batch_iter == 0:
                    self._batch = []
                self._batch.append([request, op, resp, str(start), microsec, self.model.metrics])
                self._batch_iter = (self._batch_iter + 1) % self.stream_batch
                if self._batch_iter == 0:
                    data = self.base_data()
                    data["headers"] = ["request", "op", "resp", "when", "microsec", "metrics"]
                    data["values"] = self._batch
                    self.output_stream.push([data])
            else:
                data = self.base_data()
                data["request"] = request
                data["op"] = op
                data["resp"] = resp
                data["when"] = start_str
                data["microsec"] = microsec
                if getattr(self.model, "metrics", None):
                    data["metrics"] = self.model.metrics
                self.output_stream.push([data])
