"""Serving host: events, context, GraphServer (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/serving/server.py:
  _StreamContext :48-83, GraphServer :86-312, v2_serving_init/handler :315-409,
  create_graph_server :412-434, MockTrigger/MockEvent/Response :437-490, GraphContext :493-602,
  format_error :605-614; and mlrun/datastore/__init__.py:85-127 (get_stream_pusher/_DummyStream).
"""

import asyncio
import json
import os
import socket
import traceback
import uuid

from .helpers import (
    MLRunInvalidArgumentError,
    ModelObj,
    err_to_str,
    get_caller_globals,
    get_function,
    logger as _default_logger,
)
from .topology import RootFlowStep, RouterStep, graph_root_setter  # noqa: F401
from .step_io import event_id_key, event_path_key

SERVING_SPEC_ENV = "SERVING_SPEC_ENV"


class _DummyStream:
    """stream emulator for tests (datastore/__init__.py:115-127); records pushes in event_list"""

    def __init__(self, event_list=None, **kwargs):
        self.event_list = event_list if event_list is not None else []

    def push(self, data, **kwargs):
        if not isinstance(data, list):
            data = [data]
        for item in data:
            self.event_list.append(item)


class _MockV3ioStream:
    """`v3io://...` with stream_args={"mock": True}: records {"data": json} like OutputStream's mock queue"""

    def __init__(self, **kwargs):
        self._mock_queue = []

    def push(self, data, **kwargs):
        if not isinstance(data, list):
            data = [data]
        for item in data:
            self._mock_queue.append({"data": json.dumps(item, default=_json_default)})


def _json_default(obj):
    try:
        import numpy as np

        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, np.generic):
            return obj.item()
    except ImportError:  # pragma: no cover
        pass
    return str(obj)


def get_stream_pusher(stream_path, **kwargs):
    if stream_path.startswith("dummy://"):
        return _DummyStream(**{k: v for k, v in kwargs.items() if k == "event_list"})
    if stream_path.startswith("v3io") and kwargs.get("mock"):
        return _MockV3ioStream()
    raise ValueError(f"unsupported stream path {stream_path} (oracle supports dummy:// and mocked v3io only)")


class _StreamContext:
    """server.py:48-83"""

    def __init__(self, enabled, parameters, function_uri):
        self.enabled = False
        self.hostname = socket.gethostname()
        self.function_uri = function_uri
        self.output_stream = None
        self.stream_uri = None
        log_stream = parameters.get("log_stream", "")
        if (enabled or log_stream) and function_uri:
            self.enabled = True
            project = function_uri.split("/")[0] if "/" in function_uri else "default"
            self.stream_uri = log_stream.format(project=project) if log_stream else f"dummy://{project}"
            stream_args = parameters.get("stream_args", {})
            self.output_stream = get_stream_pusher(self.stream_uri, **stream_args)


class MockTrigger:
    def __init__(self, kind="", name=""):
        self.kind = kind
        self.name = name


class MockEvent:
    """server.py:445-475"""

    def __init__(self, body=None, content_type=None, headers=None, method=None, path=None,
                 event_id=None, trigger=None, offset=None, time=None):
        self.id = event_id or uuid.uuid4().hex
        self.key = ""
        self.body = body
        self.headers = headers or {}
        self.method = method
        self.path = path or "/"
        self.content_type = content_type
        self.error = None
        self.trigger = trigger or MockTrigger()
        self.offset = offset or 0

    def __str__(self):
        error = f", error={self.error}" if self.error else ""
        return f"Event(id={self.id}, body={self.body}, method={self.method}, path={self.path}{error})"


class Response:
    """server.py:478-490"""

    def __init__(self, headers=None, body=None, content_type=None, status_code=200):
        self.headers = headers or {}
        self.body = body
        self.status_code = status_code
        self.content_type = content_type or "text/plain"

    def __repr__(self):
        args = ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items())
        return f"{self.__class__.__name__}({args})"


class GraphContext:
    """server.py:493-602 (nuclio context optional)"""

    def __init__(self, level="info", logger=None, server=None, nuclio_context=None):
        self.state = None
        self.logger = logger
        self.worker_id = 0
        self.Response = Response
        self.verbose = False
        self.stream = None
        self.root = None
        if nuclio_context:
            self.logger = nuclio_context.logger
            self.Response = nuclio_context.Response
            if hasattr(nuclio_context, "trigger") and hasattr(nuclio_context.trigger, "kind"):
                self.trigger = nuclio_context.trigger.kind
            self.worker_id = nuclio_context.worker_id
            if hasattr(nuclio_context, "platform"):
                self.platform = nuclio_context.platform
        elif not logger:
            self.logger = _default_logger
        self._server = server
        self.current_function = None
        self.get_store_resource = None
        self.get_table = None
        self.is_mock = False
        self.monitoring_mock = False

    @property
    def server(self):
        return self._server

    @property
    def project(self):
        uri = self._server.function_uri or ""
        return uri.split("/")[0] if "/" in uri else ""

    def push_error(self, event, message, source=None, **kwargs):
        if self.verbose:
            self.logger.error(f"got error from {source} state:\n{event.body}\n{message}")
        if self._server and self._server._error_stream_object:
            try:
                record = {
                    "function_uri": self._server.function_uri,
                    "worker": self.worker_id,
                    "host": socket.gethostname(),
                    "source": source,
                    "event": {"id": event.id, "body": event.body},
                    "message": message,
                    "args": kwargs,
                }
                self._server._error_stream_object.push(record)
            except Exception as ex:
                self.logger.error(f"failed to write to error stream: {ex}\n{traceback.format_exc()}")

    def get_param(self, key, default=None):
        if self._server and self._server.parameters:
            return self._server.parameters.get(key, default)
        return default

    def get_secret(self, key):
        return None


class GraphServer(ModelObj):
    """server.py:86-312"""

    kind = "server"

    def __init__(self, graph=None, parameters=None, load_mode=None, function_uri=None, verbose=False,
                 version=None, functions=None, graph_initializer=None, error_stream=None,
                 track_models=None, tracking_policy=None, secret_sources=None, default_content_type=None):
        self._graph = None
        self.graph = graph
        self.function_uri = function_uri
        self.parameters = parameters or {}
        self.verbose = verbose
        self.load_mode = load_mode or "sync"
        self.version = version or "v2"
        self.context = None
        self._current_function = None
        self.functions = functions or {}
        self.graph_initializer = graph_initializer
        self.error_stream = error_stream
        self.track_models = track_models
        self.tracking_policy = tracking_policy
        self._error_stream_object = None
        self.secret_sources = secret_sources
        self.default_content_type = default_content_type
        self.http_trigger = True

    def set_current_function(self, function):
        self._current_function = function

    @property
    def graph(self):
        return self._graph

    @graph.setter
    def graph(self, graph):
        graph_root_setter(self, graph)

    def set_error_stream(self, error_stream):
        self.error_stream = error_stream
        self._error_stream_object = get_stream_pusher(error_stream) if error_stream else None

    def init_states(self, context, namespace, resource_cache=None, logger=None, is_mock=False, monitoring_mock=False):
        """server.py:150-191"""
        if self.error_stream:
            self._error_stream_object = get_stream_pusher(self.error_stream)
        context = GraphContext(server=self, nuclio_context=context, logger=logger)
        context.is_mock = is_mock
        context.monitoring_mock = monitoring_mock
        context.root = self.graph
        context.stream = _StreamContext(self.track_models, self.parameters, self.function_uri)
        context.current_function = self._current_function
        context.verbose = self.verbose
        self.context = context
        if self.graph_initializer:
            handler = self.graph_initializer if callable(self.graph_initializer) else get_function(
                self.graph_initializer, namespace or []
            )
            handler(self)
        context.root = self.graph

    def init_object(self, namespace):
        self.graph.init_object(self.context, namespace, self.load_mode, reset=True)

    def test(self, path="/", body=None, method="", headers=None, content_type=None, silent=False,
             get_body=True, event_id=None, trigger=None, offset=None, time=None):
        """server.py:196-250"""
        if not self.graph:
            raise MLRunInvalidArgumentError(
                "no models or steps were set, use function.set_topology() and add steps"
            )
        if not method:
            method = "POST" if body else "GET"
        event = MockEvent(body=body, path=path, method=method, headers=headers, content_type=content_type,
                          event_id=event_id, trigger=trigger, offset=offset, time=time)
        resp = self.run(event, get_body=get_body)
        if hasattr(resp, "status_code") and resp.status_code >= 300 and not silent:
            raise RuntimeError(f"failed ({resp.status_code}): {resp.body}")
        return resp

    def run(self, event, context=None, get_body=False, extra_args=None):
        """server.py:252-293"""
        server_context = self.context
        context = context or server_context
        event.content_type = event.content_type or self.default_content_type or ""
        if event.headers:
            if event_id_key in event.headers:
                event.id = event.headers.get(event_id_key)
            if event_path_key in event.headers:
                event.path = event.headers.get(event_path_key)

        if isinstance(event.body, (str, bytes)) and (
            not event.content_type or event.content_type in ["json", "application/json"]
        ):
            try:
                event.body = json.loads(event.body)
            except (json.decoder.JSONDecodeError, UnicodeDecodeError) as exc:
                if event.content_type in ["json", "application/json"]:
                    message = f"failed to json decode event, {err_to_str(exc)}"
                    context.logger.error(message)
                    server_context.push_error(event, message, source="_handler")
                    return context.Response(body=message, content_type="text/plain", status_code=400)
        try:
            response = self.graph.run(event, **(extra_args or {}))
        except Exception as exc:
            message = f"{exc.__class__.__name__}: {err_to_str(exc)}"
            if server_context.verbose:
                message += "\n" + str(traceback.format_exc())
            context.logger.error(f"run error, {traceback.format_exc()}")
            server_context.push_error(event, message, source="_handler")
            return context.Response(body=message, content_type="text/plain", status_code=400)

        if asyncio.iscoroutine(response):
            response = asyncio.get_event_loop().run_until_complete(response)
        return self._process_response(context, response, get_body)

    def _process_response(self, context, response, get_body):
        """server.py:298-308"""
        body = response.body
        if isinstance(body, context.Response) or get_body:
            return body
        if body and not isinstance(body, (str, bytes)):
            body = json.dumps(body)  # strict, as upstream: a numpy value in the response is a TypeError for the caller
            return context.Response(body=body, content_type="application/json", status_code=200)
        return body

    def wait_for_completion(self):
        return self.graph.wait_for_completion()


def v2_serving_init(context, namespace=None):
    """nuclio init hook (server.py:315-350): spec comes from env SERVING_SPEC_ENV"""
    spec = json.loads(os.environ[SERVING_SPEC_ENV])
    server = GraphServer.from_dict(spec)
    if hasattr(context, "trigger"):
        server.http_trigger = getattr(context.trigger, "kind", "http") == "http"
    server.set_current_function(os.getenv("SERVING_CURRENT_FUNCTION", ""))
    kwargs = {}
    if hasattr(context, "is_mock"):
        kwargs["is_mock"] = context.is_mock
    ns = namespace or get_caller_globals()
    server.init_states(context, ns, **kwargs)
    server.init_object(ns)
    setattr(context, "mlrun_handler", v2_serving_handler)
    setattr(context, "_server", server)


def nuclio_init_hook(context, data, kind):
    """runtimes/nuclio/nuclio.py:30-39 (serving_v2 only)"""
    if kind != "serving_v2":
        raise ValueError("failed to init serving function, unsupported kind")
    v2_serving_init(context, data)


def v2_serving_handler(context, event, get_body=False):
    """server.py:387-409"""
    if context._server.http_trigger and event.body == b"":
        event.body = None
    event.stream_path = getattr(event, "topic", event.path)
    if hasattr(event, "trigger") and event.trigger.kind in (
        "kafka", "kafka-cluster", "v3ioStream", "v3io-stream", "rabbit-mq", "rabbitMq",
    ):
        event.path = "/"
    return context._server.run(event, context, get_body)


def create_graph_server(parameters=None, load_mode=None, graph=None, verbose=False, current_function=None, **kwargs):
    """server.py:412-434"""
    server = GraphServer(graph, parameters or {}, load_mode, verbose=verbose, **kwargs)
    server.set_current_function(current_function or os.getenv("SERVING_CURRENT_FUNCTION", ""))
    return server
