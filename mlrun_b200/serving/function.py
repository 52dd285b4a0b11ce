"""`new_function(kind="serving")`: the graph-building boundary object.

Plugin-API mirror of ServingRuntime (mlrun/runtimes/nuclio/serving.py: set_topology :245-306,
set_tracking :308, add_model :356-445, _get_serving_spec :645-666, to_mock_server :668-724) and of the
mock branch of `invoke` (runtimes/nuclio/function.py:925-936).  Deployment (nuclio, k8s) is out of scope.
"""

import json
from copy import deepcopy

from .graph import RootFlowStep, RouterStep, StepKinds, TaskStep, params_to_step
from .host import create_graph_server
from .resolve import MLRunInvalidArgumentError, caller_globals, logger


class ServingSpec:
    def __init__(self):
        self.graph = None
        self.parameters = {}
        self.load_mode = None
        self.graph_initializer = None
        self.error_stream = None
        self.track_models = None
        self.secret_sources = None
        self.default_content_type = None
        self.default_class = None
        self.command = ""  # path of the function's own code file: its classes / handlers are step candidates


class ServingRuntime:
    kind = "serving"

    def __init__(self, name="", project="", tag=""):
        self.name = name
        self.project = project or "default"
        self.tag = tag
        self.spec = ServingSpec()
        self.verbose = False
        self._mock = None

    def _function_uri(self):
        uri = f"{self.project}/{self.name}"
        return f"{uri}:{self.tag}" if self.tag else uri

    def set_topology(self, topology=None, class_name=None, engine=None, exist_ok=False, **class_args):
        topology = topology or StepKinds.router
        if self.spec.graph and not exist_ok:
            raise MLRunInvalidArgumentError("graph topology is already set, cannot be overwritten")
        if topology == StepKinds.router:
            if class_name is not None and hasattr(class_name, "to_dict"):
                _, step = params_to_step(class_name, None)
                if step.kind != StepKinds.router:
                    raise MLRunInvalidArgumentError(
                        "provided class is not a router step, must provide a router class in router topology")
            else:
                step = RouterStep(class_name=class_name, class_args=class_args)
            self.spec.graph = step
        elif topology == StepKinds.flow:
            self.spec.graph = RootFlowStep(engine=engine)
        else:
            raise MLRunInvalidArgumentError(f"unsupported topology {topology}, use 'router' or 'flow'")
        return self.spec.graph

    def set_tracking(self, stream_path=None, batch=None, sample=None, stream_args=None, tracking_policy=None,
                     enable_tracking=True):
        self.spec.track_models = enable_tracking  # `tracking_policy`: deprecated upstream, no effect
        for key, val in (("log_stream", stream_path), ("log_stream_batch", batch), ("log_stream_sample", sample),
                         ("stream_args", stream_args)):
            if val:
                self.spec.parameters[key] = val

    def add_model(self, key, model_path=None, class_name=None, model_url=None, handler=None, router_step=None,
                  child_function=None, **class_args):
        graph = self.spec.graph or self.set_topology()
        if graph.kind != StepKinds.router:
            if router_step:
                if router_step not in graph:
                    raise ValueError(f"router step {router_step} not present in the graph")
                graph = graph[router_step]
            else:
                routers = [s for s in graph.steps.values() if s.kind == StepKinds.router]
                if not routers:
                    raise ValueError("graph does not contain any router, add_model can only be used when there is a router step")
                if len(routers) > 1:
                    raise ValueError(f"found {len(routers)} routers, please specify the router_step you would like to add this model to")
                graph = routers[0]
        if class_name is not None and hasattr(class_name, "to_dict"):
            if model_path:
                class_name.model_path = model_path
            key, state = params_to_step(class_name, key)
        else:
            if not model_path and not model_url:
                raise ValueError("model_path or model_url must be provided")
            class_name = class_name or self.spec.default_class
            if class_name and not isinstance(class_name, str):
                raise ValueError("class name must be a string (name of module.submodule.name)")
            if model_path and not class_name:
                raise ValueError("model_path must be provided with class_name")
            if model_url:
                raise MLRunInvalidArgumentError("remote model endpoints ($remote) are out of scope of this engine")
            args = deepcopy(class_args)
            args["model_path"] = str(model_path)
            state = TaskStep(class_name, args, handler=handler, function=child_function)
        return graph.add_route(key, state)

    def _get_serving_spec(self):
        return json.dumps({
            "function_uri": self._function_uri(), "version": "v2", "parameters": self.spec.parameters,
            "graph": self.spec.graph.to_dict() if self.spec.graph else {}, "load_mode": self.spec.load_mode,
            "functions": {}, "graph_initializer": self.spec.graph_initializer, "error_stream": self.spec.error_stream,
            "track_models": self.spec.track_models, "tracking_policy": None,
            "default_content_type": self.spec.default_content_type,
        }, default=str)

    def to_mock_server(self, namespace=None, current_function="*", track_models=False, workdir=None, **kwargs):
        namespace = namespace or []
        if not isinstance(namespace, list):
            namespace = [namespace]
        module = _code_module(self.spec.command, workdir)  # the function's own code (mlrun.run.function_to_module, silent)
        if module is not None:
            namespace.append(module)
        namespace.append(caller_globals())
        server = create_graph_server(
            parameters=self.spec.parameters, load_mode=self.spec.load_mode, graph=self.spec.graph, verbose=self.verbose,
            current_function=current_function, graph_initializer=self.spec.graph_initializer,
            track_models=self.spec.track_models, function_uri=self._function_uri(), secret_sources=self.spec.secret_sources,
            default_content_type=self.spec.default_content_type, **kwargs)
        server.init_states(context=None, namespace=namespace, logger=logger, is_mock=True, monitoring_mock=track_models)
        server.init_object(namespace)
        return server

    def invoke(self, path, body=None, method=None, headers=None, **kwargs):
        if self._mock is None:
            self._mock = self.to_mock_server()
        return self._mock.test(path, body, method or ("POST" if body else "GET"), headers)


def _code_module(command, workdir=None):
    """the function's code file as a module (mlrun/run.py:77-127 with silent=True: nothing to load -> None)"""
    import importlib.util
    import os

    if not command:
        return None
    path = os.path.join(workdir or "", command)
    spec = importlib.util.spec_from_file_location(os.path.splitext(os.path.basename(path))[0], path)
    if spec is None:
        raise OSError(f"cannot import from {path!r}")
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def new_function(name="", project="", tag="", kind="", command="", **kwargs):
    if kind != "serving":
        raise MLRunInvalidArgumentError("mlrun_b200 implements kind='serving' functions only")
    fn = ServingRuntime(name=name, project=project, tag=tag)
    fn.spec.command = command or ""
    return fn


ServingFunction = ServingRuntime  # the name this class had before it took the reference's (mlrun.runtimes.ServingRuntime)
