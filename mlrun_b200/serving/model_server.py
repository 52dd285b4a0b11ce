"""V2 model-server protocol (plugin-API mirror of mlrun/serving/v2_serving.py:32-504).

Subclass it exactly like the reference class: override `load()` / `predict(request)` and optionally
`preprocess`, `validate`, `postprocess`, `explain`, `logged_results`, `op_<name>`.  Device-backed
servers (mlrun_b200.serving.device_models) subclass it too and run `predict` as a CUDA plan.
"""

import threading
import time
import traceback
from datetime import datetime, timezone

from .paths import merge_result, select_input
from .resolve import MLRunInvalidArgumentError
from .step_meta import StepMeta

_PREDICT_OPS = ("predict", "infer", "infer_dict", "predict_dict")


def now_date():
    return datetime.now(timezone.utc)


class V2ModelServer(StepMeta):
    def __init__(self, context=None, name=None, model_path=None, model=None, protocol=None, input_path=None,
                 result_path=None, **kwargs):
        self.name, self.version = name, ""
        if name and ":" in name:  # "<model>:<version>" keys come from /versions/<ver>/ URLs
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
        self._model_logger = _ModelLogPusher(self, context) if context is not None and context.stream.enabled else None
        self.metrics = {}
        self.labels = {}
        self.model = None
        if model:
            self.model = model
            self.ready = True
        self.model_endpoint_uid = None

    # ---- lifecycle ------------------------------------------------------------------------------
    def _load_and_flag(self):
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
        if mode == "async":
            threading.Thread(target=self._load_and_flag, daemon=True).start()
            self.context.logger.info(f"started async model loading for {self.name}")
        else:
            self._load_and_flag()

    def load(self):
        if not self.ready and not self.model:
            raise ValueError("please specify a load method or a model object")

    def get_model(self, suffix=""):
        """(local model file, extra data) -- the local-filesystem slice of mlrun.artifacts.get_model (artifacts/model.py:
        412-483): the path itself when it carries the suffix, else the first entry of the directory that does.  Store URIs,
        model-spec yaml files and extra data items are the artifact store (out of scope)."""
        import os

        suffix = suffix or ".pkl"
        path = str(self.model_path)
        if path.endswith(suffix):
            return path, {}
        found = next((os.path.join(path, f) for f in (os.listdir(path) if os.path.isdir(path) else ()) if f.endswith(suffix)), "")
        if not found:
            raise ValueError(f"cant resolve model file for {path} suffix{suffix}")
        return found, {}

    def get_param(self, key, default=None):
        if key in self._params:
            return self._params.get(key)
        return self.context.get_param(key, default=default)

    def set_metric(self, name, value):
        self.metrics[name] = value

    # ---- hooks ----------------------------------------------------------------------------------
    def preprocess(self, request, operation):
        return request

    def postprocess(self, request):
        return request

    def validate(self, request, operation):
        if self.protocol == "v2":
            if "inputs" not in request:
                raise Exception('Expected key "inputs" in request body')
            if not isinstance(request["inputs"], list):
                raise Exception('Expected "inputs" to be a list')
        return request

    def predict(self, request):
        raise NotImplementedError()

    def explain(self, request):
        raise NotImplementedError()

    def logged_results(self, request, response, op):
        return None, None

    # ---- event handler --------------------------------------------------------------------------
    def _wait_ready(self, event):
        if self.ready:
            return
        if not event.trigger or event.trigger.kind in ("http", ""):
            raise RuntimeError(f"model {self.name} is not ready yet")
        self.context.logger.info(f"waiting for model {self.name} to load")
        for _ in range(50):
            time.sleep(5)
            if self.ready:
                return
        raise RuntimeError(f"model {self.name} is not ready {self.error}")

    def _request_of(self, event, body, op):
        self._wait_ready(event)
        if "_dict" in op:
            body = self._inputs_to_list(body)
        return self.validate(self.preprocess(body, op), op)

    def _run_op(self, fn, request, start, event_id, op):
        try:
            return fn(request)
        except Exception as exc:
            request["id"] = event_id
            if self._model_logger:
                self._model_logger.push(start, request, op=op, error=exc)
            raise

    def do_event(self, event, *args, **kwargs):
        start = now_date()
        original = event.body
        body = select_input(self._input_path, event.body)
        event_id = event.id
        op = event.path.strip("/")
        if body and isinstance(body, dict):
            op = op or body.get("operation")
            event_id = body.get("id", event_id)
        if not op and event.method != "GET":
            op = "infer"

        if op in _PREDICT_OPS:
            request = self._request_of(event, body, op)
            outputs = self._run_op(self.predict, request, start, event_id, op)
            response = {"id": event_id, "model_name": self.name, "outputs": outputs,
                        "timestamp": start.isoformat(sep=" ", timespec="microseconds")}
        elif op == "ready" and event.method == "GET":
            event.terminated = True
            if self.ready:
                text = f"Model {self.name} is ready (event_id = {event_id})"
                event.body = self.context.Response(status_code=200, body=bytes(text, encoding="utf-8"))
            else:
                event.body = self.context.Response(status_code=408, body=b"model not ready")
            return event
        elif op == "" and event.method == "GET":
            event.terminated = True
            meta = {"name": self.name, "version": self.version, "inputs": [], "outputs": []}
            if self.model_spec:
                meta["inputs"] = self.model_spec.inputs.to_dict()
                meta["outputs"] = self.model_spec.outputs.to_dict()
            event.body = merge_result(self._result_path, original, meta)
            return event
        elif op == "explain":
            request = self._request_of(event, body, op)
            outputs = self._run_op(self.explain, request, start, event_id, op)
            response = {"id": event_id, "model_name": self.name, "outputs": outputs}
        elif hasattr(self, "op_" + op):
            event.body = merge_result(self._result_path, original, getattr(self, "op_" + op)(event))
            return event
        else:
            raise ValueError(f"illegal model operation {op}, method={event.method}")

        if self.version:
            response["model_version"] = self.version
        response = self.postprocess(response)
        if self._model_logger:
            inputs, outputs = self.logged_results(request, response, op)
            if inputs is None and outputs is None:
                self._model_logger.push(start, request, response, op)
            else:
                self._model_logger.push(start, {"id": event_id, "inputs": inputs or []}, {"outputs": outputs or []}, op)
        event.body = merge_result(self._result_path, original, response)
        return event

    def _inputs_to_list(self, request):
        if not (self.model_spec and self.model_spec.inputs):
            raise MLRunInvalidArgumentError(
                "In order to use predict_dict or infer_dict operation you have to provide `model_path` "
                "to the model server and to load it by `load()` function")
        order = [f.name for f in self.model_spec.inputs]
        inputs = request.get("inputs")
        try:
            if isinstance(inputs, list) and all(isinstance(i, dict) for i in inputs):
                request["inputs"] = [[d[k] for k in order] for d in inputs]
            elif isinstance(inputs, dict):
                request["inputs"] = [inputs[k] for k in order]
            else:
                raise MLRunInvalidArgumentError(
                    "When using predict_dict or infer_dict operation the inputs must be of type `list[dict]` or `dict`")
        except KeyError:
            raise MLRunInvalidArgumentError(f"Input dictionary don't contain all the necessary input keys : {order}")
        return request


class _ModelLogPusher:
    """tracking-stream records (v2_serving.py:429-504): sampling, micro-batching, record layout"""

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
        rec = {"class": type(self.model).__name__, "worker": self._worker, "model": self.model.name,
               "version": self.model.version, "host": self.hostname, "function_uri": self.function_uri}
        if getattr(self.model, "labels", None):
            rec["labels"] = self.model.labels
        return rec

    def push(self, start, request, resp=None, op=None, error=None):
        when = start.isoformat(sep=" ", timespec="microseconds")
        if error:
            rec = self.base_data()
            message = f"{error}\n{traceback.format_exc()}" if self.verbose else str(error)
            rec.update(request=request, op=op, when=when, error=message)
            self.output_stream.push([rec])
            return
        self._sample_iter = (self._sample_iter + 1) % self.stream_sample
        if not self.output_stream or self._sample_iter:
            return
        microsec = (now_date() - start).microseconds
        if self.stream_batch > 1:
            if self._batch_iter == 0:
                self._batch = []
            self._batch.append([request, op, resp, str(start), microsec, self.model.metrics])
            self._batch_iter = (self._batch_iter + 1) % self.stream_batch
            if self._batch_iter == 0:
                rec = self.base_data()
                rec["headers"] = ["request", "op", "resp", "when", "microsec", "metrics"]
                rec["values"] = self._batch
                self.output_stream.push([rec])
            return
        rec = self.base_data()
        rec.update(request=request, op=op, resp=resp, when=when, microsec=microsec)
        if getattr(self.model, "metrics", None):
            rec["metrics"] = self.model.metrics
        self.output_stream.push([rec])

    def push_batch(self, start, requests, responses, op="infer", microsec=None):
        """tracking records for a batch the engine scored in one launch: exactly the records -- same
        sampling positions, same micro-batch boundaries, same layout (v2_serving.py:457-504) -- that `push` would have
        produced had the events been pushed one by one; `requests[i]` / `responses[i]` are built only for the sampled rows.

        requests / responses: callables i -> dict (lazy), or sequences."""
        n = len(requests) if hasattr(requests, "__len__") else int(requests.n)
        if not self.output_stream or n == 0:
            return 0
        get_req = requests if callable(requests) else requests.__getitem__
        get_resp = responses if callable(responses) else responses.__getitem__
        when = start.isoformat(sep=" ", timespec="microseconds")
        if microsec is None:
            microsec = (now_date() - start).microseconds
        # rows i with (iter0 + i + 1) % sample == 0 are logged
        first = (-(self._sample_iter + 1)) % self.stream_sample
        picked = range(first, n, self.stream_sample)
        self._sample_iter = (self._sample_iter + n) % self.stream_sample
        pushed = 0
        for i in picked:
            request, resp = get_req(i), get_resp(i)
            if self.stream_batch > 1:
                if self._batch_iter == 0:
                    self._batch = []
                self._batch.append([request, op, resp, str(start), microsec, self.model.metrics])
                self._batch_iter = (self._batch_iter + 1) % self.stream_batch
                if self._batch_iter == 0:
                    rec = self.base_data()
                    rec["headers"] = ["request", "op", "resp", "when", "microsec", "metrics"]
                    rec["values"] = self._batch
                    self.output_stream.push([rec])
                    pushed += 1
                continue
            rec = self.base_data()
            rec.update(request=request, op=op, resp=resp, when=when, microsec=microsec)
            if getattr(self.model, "metrics", None):
                rec["metrics"] = self.model.metrics
            self.output_stream.push([rec])
            pushed += 1
        return pushed
