"""`new_function(kind="serving")` boundary object (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/runtimes/nuclio/serving.py:
  set_topology :245-306, set_tracking :308-354, add_model :356-445, _get_serving_spec :645-666,
  to_mock_server :668-724; and the mock branch of invoke (runtimes/nuclio/function.py:925-936).
Deployment (`deploy`, k8s, nuclio) is out of scope.
"""

import json
from copy import deepcopy

from .helpers import MLRunInvalidArgumentError, get_caller_globals, logger
from .host import create_graph_server
from .topology import RootFlowStep, RouterStep, StepKinds, TaskStep, params_to_step


class _ServingSpec:
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
        self.function_refs = {}


class ServingRuntime:
    kind = "serving"

    def __init__(self, name="", project="", tag=""):
        self.metadata_name = name
        self.project = project or "default"
        self.tag = tag
        self.spec = _ServingSpec()
        self.verbose = False
        self._mock_server = None

    def _function_uri(self):
        uri = f"{self.project}/{self.metadata_name}"
        return f"{uri}:{self.tag}" if self.tag else uri

    def set_topology(self, topology=None, class_name=None, engine=None, exist_ok=False, **class_args):
        """serving.py:245-306 -- the graph root: a router (default) or a flow"""
        topology = topology or StepKinds.router
        if self.spec.graph and not exist_ok:
            raise MLRunInvalidArgumentError("graph topology is already set, cannot be overwritten")
        if topology == StepKinds.flow:
            root = RootFlowStep(engine=engine)
        elif topology != StepKinds.router:
            raise MLRunInvalidArgumentError(f"unsupported topology {topology}, use 'router' or 'flow'")
        elif class_name and hasattr(class_name, "to_dict"):  # a router instance
            root = params_to_step(class_name, None)[1]
            if root.kind != StepKinds.router:
                raise MLRunInvalidArgumentError(
                    "provided class is not a router step, must provide a router class in router topology")
        else:
            root = RouterStep(class_name=class_name, class_args=class_args)
        self.spec.graph = root
        return root

    def set_tracking(self, stream_path=None, batch=None, sample=None, stream_args=None, tracking_policy=None,
                     enable_tracking=True):
        """serving.py:308-354 -- tracking is a flag plus stream parameters read by the model servers (`tracking_policy` is
        deprecated upstream and has no effect)"""
        self.spec.track_models = enable_tracking
        given = {"log_stream": stream_path, "log_stream_batch": batch, "log_stream_sample": sample, "stream_args": stream_args}
        self.spec.parameters.update({k: v for k, v in given.items() if v})

    def _router_for(self, router_step):
        """the router a model is added to: the root, a named step of a flow, or the flow's only router"""
        graph = self.spec.graph or self.set_topology()
        if graph.kind == StepKinds.router:
            return graph
        if router_step:
            if router_step not in graph:
                raise ValueError(f"router step {router_step} not present in the graph")
            return graph[router_step]
        routers = [step for step in graph.steps.values() if step.kind == StepKinds.router]
        if not routers:
            raise ValueError("graph does not contain any router, add_model can only be used when there is a router step")
        if len(routers) > 1:
            raise ValueError(f"found {len(routers)} routers, please specify the router_step you would like to add this model to")
        return routers[0]

    def add_model(self, key, model_path=None, class_name=None, model_url=None, handler=None,
                  router_step=None, child_function=None, **class_args):
        """serving.py:356-445"""
        router = self._router_for(router_step)
        if class_name and hasattr(class_name, "to_dict"):  # a model-server instance
            if model_path:
                class_name.model_path = model_path
            key, route = params_to_step(class_name, key)
            return router.add_route(key, route)
        if not (model_path or model_url):
            raise ValueError("model_path or model_url must be provided")
        class_name = class_name or self.spec.default_class
        if class_name and not isinstance(class_name, str):
            raise ValueError("class name must be a string (name of module.submodule.name)")
        if model_path and not class_name:
            raise ValueError("model_path must be provided with class_name")
        args = deepcopy(class_args)
        args["model_path"] = str(model_path) if model_path else model_path
        return router.add_route(key, TaskStep(class_name, args, handler=handler, function=child_function))

    def _get_serving_spec(self):
        return json.dumps(
            {
                "function_uri": self._function_uri(),
                "version": "v2",
                "parameters": self.spec.parameters,
                "graph": self.spec.graph.to_dict() if self.spec.graph else {},
                "load_mode": self.spec.load_mode,
                "functions": {},
                "graph_initializer": self.spec.graph_initializer,
                "error_stream": self.spec.error_stream,
                "track_models": self.spec.track_models,
                "tracking_policy": None,
                "default_content_type": self.spec.default_content_type,
            }
        )

    def to_mock_server(self, namespace=None, current_function="*", track_models=False, workdir=None, **kwargs):
        namespace = namespace or []
        if not isinstance(namespace, list):
            namespace = [namespace]
        code = self._own_code(workdir)
        if code is not None:
            namespace.append(code)
        namespace.append(get_caller_globals())
        server = create_graph_server(
            parameters=self.spec.parameters,
            load_mode=self.spec.load_mode,
            graph=self.spec.graph,
            verbose=self.verbose,
            current_function=current_function,
            graph_initializer=self.spec.graph_initializer,
            track_models=self.spec.track_models,
            function_uri=self._function_uri(),
            secret_sources=self.spec.secret_sources,
            default_content_type=self.spec.default_content_type,
            **kwargs,
        )
        server.init_states(context=None, namespace=namespace, logger=logger, is_mock=True,
                           monitoring_mock=track_models)
        server.init_object(namespace)
        return server

    def _own_code(self, workdir):
        """mlrun.run.function_to_module(self, silent=True) (run.py:77-127): the function's code file, loaded as a module
        whose classes and functions are step candidates; no code -> None"""
        import importlib.util as loader
        import os
        from pathlib import Path

        command = getattr(self.spec, "command", "")
        if not command:
            return None
        location = os.path.join(workdir or "", command)
        found = loader.spec_from_file_location(Path(location).stem, location)
        if found is None:
            raise OSError(f"cannot import from {location!r}")
        module = loader.module_from_spec(found)
        found.loader.exec_module(module)
        return module

    def invoke(self, path, body=None, method=None, headers=None, **kwargs):
        """mock invoke only (function.py:925-936)"""
        if self._mock_server is None:
            self._mock_server = self.to_mock_server()
        if not method:
            method = "POST" if body else "GET"
        return self._mock_server.test(path, body, method, headers)


def new_function(name="", project="", tag="", kind="", command="", **kwargs):
    """mlrun.run.new_function (run.py:425) for kind="serving" only; `command` = the function's code file"""
    if kind != "serving":
        raise MLRunInvalidArgumentError("the oracle only restates kind='serving' functions")
    fn = ServingRuntime(name=name, project=project, tag=tag)
    fn.spec.command = command or ""
    return fn


ServingFunction = ServingRuntime  # the name this class had before it took the reference's (mlrun.runtimes.ServingRuntime)
