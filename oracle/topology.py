"""Graph topology + per-event executor (oracle restatement; test infrastructure).

Follows (behaviour, not text) mlrun/serving/states.py:
  BaseStep :102-395, TaskStep :398-599, ErrorStep :635-668, RouterStep :671-798,
  QueueStep :801-889, FlowStep :892-1402, RootFlowStep :1405-1409, params_to_step :1548-1619,
  _init_async_objects :1622-1710 (storey DAG -> emulated here by `_AsyncFlow`).

storey (the async engine) is a third-party dependency (storey~=1.8.0) whose source is not under
/root/reference.  `_AsyncFlow` restates the contract visible from the call sites: Map(fn,
full_event, input_path, result_path), fan-out to every outlet, Complete(full_event=True)
resolving the caller's awaitable with the event, recovery steps, emit/await across a loop thread,
terminate/await_termination.  It is pinned only by the literal expectations of the reference's
async tests (tests/serving/test_async_flow.py, test_flow.py async params) -- "parity unpinned"
beyond those.
"""

import copy as _copy
import os
import queue as _queue
import threading
import traceback
from inspect import getfullargspec, signature

from .helpers import (
    MLRunInvalidArgumentError,
    ModelObj,
    ObjectDict,
    err_to_str,
    get_class,
    get_function,
)
from .step_io import StepToDict, _extract_input_data, _update_result_body

callable_prefix = "_"
path_splitter = "/"
previous_step = "$prev"
queue_class_names = [">>", "$queue"]
MAX_ALLOWED_STEPS = 4500  # states.py:87


class GraphError(Exception):
    """error in graph topology or configuration (states.py:52-55)"""


class StepKinds:
    router = "router"
    task = "task"
    flow = "flow"
    queue = "queue"
    choice = "choice"
    root = "root"
    error_step = "error_step"


_task_step_fields = [
    "kind",
    "class_name",
    "class_args",
    "handler",
    "skip_context",
    "after",
    "function",
    "comment",
    "shape",
    "full_event",
    "on_error",
    "responder",
    "input_path",
    "result_path",
]


class MapClass:
    """stand-in for storey.MapClass: the base the feature-store steps derive from.  Holds the
    kwargs storey's flow base keeps (context/name/full_event/input_path/result_path)."""

    def __init__(self, context=None, name=None, full_event=None, input_path=None, result_path=None, **kwargs):
        self.context = context
        self.name = name
        self._full_event = full_event
        self._input_path = input_path
        self._result_path = result_path
        self.logger = getattr(context, "logger", None) if context else None
        self._kwargs = kwargs
        self._outlets = []  # marks a "native" async step (states.py:1678)


def get_current_function(context):
    if context and hasattr(context, "current_function"):
        return context.current_function or ""
    return ""


def get_name(name, class_name):
    if name:
        return name
    if not class_name:
        raise MLRunInvalidArgumentError("name or class_name must be provided")
    if isinstance(class_name, type):
        return class_name.__name__
    return class_name


class BaseStep(ModelObj):
    kind = "BaseStep"
    default_shape = "ellipse"
    _dict_fields = ["kind", "comment", "after", "on_error"]

    def __init__(self, name=None, after=None, shape=None):
        self.name = name
        self._parent = None
        self.comment = None
        self.context = None
        self.after = after or []
        self._next = None
        self.shape = shape
        self.on_error = None
        self._on_error_handler = None

    def set_parent(self, parent):
        self._parent = parent

    @property
    def next(self):
        return self._next

    @property
    def parent(self):
        return self._parent

    def set_next(self, key):
        if not self.next:
            self._next = [key]
        elif key not in self.next:
            self._next.append(key)
        return self

    def after_step(self, *after, append=True):
        if not append:
            self.after = []
        for name in after:
            name = name if isinstance(name, str) else name.name
            if name not in self.after:
                self.after.append(name)
        return self

    def error_handler(
        self,
        name=None,
        class_name=None,
        handler=None,
        before=None,
        function=None,
        full_event=None,
        input_path=None,
        result_path=None,
        **class_args,
    ):
        """states.py:155-231"""
        if not (class_name or handler):
            raise MLRunInvalidArgumentError("class_name or handler must be provided")
        if isinstance(self, RootFlowStep) and before:
            raise MLRunInvalidArgumentError("`before` arg can't be specified for graph error handler")
        name = get_name(name, class_name)
        step = ErrorStep(
            class_name,
            class_args,
            handler,
            name=name,
            function=function,
            full_event=full_event,
            input_path=input_path,
            result_path=result_path,
        )
        self.on_error = name
        before = [before] if isinstance(before, str) else before
        step.before = before or []
        step.base_step = self.name
        if getattr(self, "_parent", None):
            step = self._parent._steps.update(name, step)
            step.set_parent(self._parent)
        else:
            step = self._steps.update(name, step)
            step.set_parent(self)
        return self

    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        self.context = context

    def _is_local_function(self, context):
        return True

    def get_children(self):
        return []

    def __iter__(self):
        yield from []

    @property
    def fullname(self):
        name = self.name or ""
        if self._parent and self._parent.fullname:
            name = path_splitter.join([self._parent.fullname, name])
        return name.replace(":", "_")

    def _post_init(self, mode="sync"):
        pass

    def _set_error_handler(self):
        if self.on_error:
            error_step = self.context.root.path_to_step(self.on_error)
            self._on_error_handler = error_step.run

    def _log_error(self, event, err, **kwargs):
        message = err_to_str(err)
        self.context.logger.error(
            f"step {self.name} got error {message} when processing an event:\n {event.body}"
        )
        trace = traceback.format_exc()
        self.context.logger.error(trace)
        self.context.push_error(event, f"{message}\n{trace}", source=self.fullname, **kwargs)

    def _call_error_handler(self, event, err, **kwargs):
        """states.py:276-282"""
        if not event.error:
            event.error = {}
        event.error[self.name] = err_to_str(err)
        event.origin_state = self.fullname
        return self._on_error_handler(event)

    def path_to_step(self, path):
        path = path or ""
        level = self
        for part in path.split(path_splitter):
            if part not in level:
                raise GraphError(f"step {part} doesnt exist in the graph under {level.fullname}")
            level = level[part]
        return level

    def to(
        self,
        class_name=None,
        name=None,
        handler=None,
        graph_shape=None,
        function=None,
        full_event=None,
        input_path=None,
        result_path=None,
        **class_args,
    ):
        """append a step after this one (states.py:297-362)"""
        if hasattr(self, "steps"):
            parent = self
        elif self._parent:
            parent = self._parent
        else:
            raise GraphError(f"step {self.name} parent is not set or it's not part of a graph")
        name, step = params_to_step(
            class_name,
            name,
            handler,
            graph_shape=graph_shape,
            function=function,
            full_event=full_event,
            input_path=input_path,
            result_path=result_path,
            class_args=class_args,
        )
        step = parent._steps.update(name, step)
        step.set_parent(parent)
        if not hasattr(self, "steps"):
            step.after_step(self.name)
        parent._last_added = step
        return step

    def set_flow(self, steps, force=False):
        raise NotImplementedError("set_flow() can only be called on a FlowStep")

    def supports_termination(self):
        return False


class TaskStep(BaseStep):
    """runs a class or a handler (states.py:398-599)"""

    kind = "task"
    _dict_fields = _task_step_fields
    _default_class = ""

    def __init__(
        self,
        class_name=None,
        class_args=None,
        handler=None,
        name=None,
        after=None,
        full_event=None,
        function=None,
        responder=None,
        input_path=None,
        result_path=None,
    ):
        super().__init__(name, after)
        self.class_name = class_name
        self.class_args = class_args or {}
        self.handler = handler
        self.function = function
        self._handler = None
        self._object = None
        self._async_object = None
        self.skip_context = None
        self.context = None
        self._class_object = None
        self.responder = responder
        self.full_event = full_event
        self.input_path = input_path
        self.result_path = result_path
        self.on_error = None
        self._inject_context = False
        self._call_with_event = False

    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        self.context = context
        self._async_object = None
        if not self._is_local_function(context):
            return

        if self.handler and not self.class_name:
            if callable(self.handler):
                self._handler = self.handler
                self.handler = self.handler.__name__
            else:
                self._handler = get_function(self.handler, namespace)
            try:
                params = signature(self._handler).parameters
            except (TypeError, ValueError):
                params = {}
            if params and "context" in list(params.keys()):
                self._inject_context = True
            self._set_error_handler()
            return

        self._class_object, self.class_name = self.get_step_class_object(namespace)
        if not self._object or reset:
            ctor_args = self.get_full_class_args(namespace, self._class_object, **extra_kwargs)
            try:
                self._object = self._class_object(**ctor_args)
            except TypeError as exc:
                raise TypeError(f"failed to init step {self.name}\n args={self.class_args}") from exc

            handler = self.handler
            if handler:
                if not hasattr(self._object, handler):
                    raise GraphError(
                        f"handler ({handler}) specified but doesnt exist in class {self.class_name}"
                    )
            elif hasattr(self._object, "do_event"):
                handler = "do_event"
                self._call_with_event = True
            elif hasattr(self._object, "do"):
                handler = "do"
            if handler:
                self._handler = getattr(self._object, handler, None)

        self._set_error_handler()
        if mode != "skip":
            self._post_init(mode)

    def get_full_class_args(self, namespace, class_object, **extra_kwargs):
        """states.py:494-512: `_x` args resolve to callables; name/context/... only when accepted"""
        args = {}
        for key, val in self.class_args.items():
            if key.startswith(callable_prefix):
                args[key[1:]] = get_function(val, namespace)
            else:
                args[key] = val
        args.update(extra_kwargs)
        spec = getfullargspec(class_object)
        for key in ["name", "context", "input_path", "result_path", "full_event"]:
            if spec.varkw or key in spec.args:
                args[key] = getattr(self, key)
        if spec.varkw or "graph_step" in spec.args:
            args["graph_step"] = self
        return args

    def get_step_class_object(self, namespace):
        class_name = self.class_name
        class_object = self._class_object
        if isinstance(class_name, type):
            class_object = class_name
            class_name = class_name.__name__
        elif not class_object:
            class_object = get_class(class_name or self._default_class, namespace)
        return class_object, class_name

    def _is_local_function(self, context):
        """states.py:529-540"""
        current = get_current_function(context)
        if current == "*":
            return True
        if not self.function and not current:
            return True
        if (self.function and self.function == "*") or self.function == current:
            return True
        return False

    @property
    def async_object(self):
        return self._async_object or self._object

    def clear_object(self):
        self._object = None

    def _post_init(self, mode="sync"):
        if self._object and hasattr(self._object, "post_init"):
            self._object.post_init(mode)

    def respond(self):
        self.responder = True
        return self

    def run(self, event, *args, **kwargs):
        """per-event call convention (states.py:564-599)"""
        if not self._is_local_function(self.context):
            return event
        if self._inject_context:
            kwargs["context"] = self.context
        elif kwargs and "context" in kwargs:
            del kwargs["context"]

        try:
            if self.full_event or self._call_with_event:
                return self._handler(event, *args, **kwargs)
            if self._handler is None:
                raise MLRunInvalidArgumentError(f"step {self.name} does not have a handler")
            result = self._handler(_extract_input_data(self.input_path, event.body), *args, **kwargs)
            event.body = _update_result_body(self.result_path, event.body, result)
        except Exception as exc:
            if self._on_error_handler:
                self._log_error(event, exc)
                result = self._call_error_handler(event, exc)
                event.body = _update_result_body(self.result_path, event.body, result)
            else:
                raise exc
        return event


class ErrorStep(TaskStep):
    kind = "error_step"
    _dict_fields = _task_step_fields + ["before", "base_step"]
    _default_class = ""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.before = None
        self.base_step = None


class RouterStep(TaskStep):
    """router with child routes (states.py:671-798)"""

    kind = "router"
    default_shape = "doubleoctagon"
    _dict_fields = _task_step_fields + ["routes"]
    _default_class = "mlrun.serving.ModelRouter"

    def __init__(
        self,
        class_name=None,
        class_args=None,
        handler=None,
        routes=None,
        name=None,
        function=None,
        input_path=None,
        result_path=None,
    ):
        super().__init__(
            class_name,
            class_args,
            handler,
            name=name,
            function=function,
            input_path=input_path,
            result_path=result_path,
        )
        self._routes = None
        self.routes = routes

    def get_children(self):
        return self._routes.values()

    @property
    def routes(self):
        return self._routes

    @routes.setter
    def routes(self, routes):
        self._routes = ObjectDict.from_dict(classes_map, routes, "task")

    def add_route(self, key, route=None, class_name=None, handler=None, function=None, **class_args):
        if not route and not class_name and not handler:
            raise MLRunInvalidArgumentError("route or class_name must be specified")
        if not route:
            route = TaskStep(class_name, class_args, handler=handler)
        route.function = function or route.function
        if len(self._routes) >= MAX_ALLOWED_STEPS:
            raise MLRunInvalidArgumentError(
                f"Cannot create the serving graph: the maximum number of steps is {MAX_ALLOWED_STEPS}"
            )
        route = self._routes.update(key, route)
        route.set_parent(self)
        return route

    def clear_children(self, routes=None):
        for key in list(routes or self._routes.keys()):
            del self._routes[key]

    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        if not self._is_local_function(context):
            return
        self.class_args = self.class_args or {}
        super().init_object(context, namespace, "skip", reset=reset, routes=self._routes, **extra_kwargs)
        for route in self._routes.values():
            if self.function and not route.function:
                route.function = self.function
            route.set_parent(self)
            route.init_object(context, namespace, mode, reset=reset)
        self._set_error_handler()
        self._post_init(mode)

    def __getitem__(self, name):
        return self._routes[name]

    def __setitem__(self, name, route):
        self.add_route(name, route)

    def __delitem__(self, key):
        del self._routes[key]

    def __iter__(self):
        yield from self._routes.keys()

    def __contains__(self, name):
        return name in self._routes


class QueueStep(BaseStep):
    """queue / stream step (states.py:801-889); in the oracle only `dummy://` and path-less queues exist"""

    kind = "queue"
    default_shape = "cds"
    _dict_fields = BaseStep._dict_fields + ["path", "shards", "retention_in_hours", "trigger_args", "options"]

    def __init__(self, name=None, path=None, after=None, shards=None, retention_in_hours=None, trigger_args=None, **options):
        super().__init__(name, after)
        self.path = path
        self.shards = shards
        self.retention_in_hours = retention_in_hours
        self.options = options
        self.trigger_args = trigger_args
        self._stream = None
        self._async_object = None

    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        self.context = context
        if self.path:
            from .host import get_stream_pusher

            self._stream = get_stream_pusher(self.path, **self.options)
        self._set_error_handler()

    @property
    def async_object(self):
        return self._async_object

    def to(self, class_name=None, name=None, handler=None, graph_shape=None, function=None,
           full_event=None, input_path=None, result_path=None, **class_args):
        if not function:
            name = get_name(name, class_name)
            raise MLRunInvalidArgumentError(
                f"step '{name}' must specify a function, because it follows a queue step"
            )
        return super().to(class_name, name, handler, graph_shape, function, full_event,
                          input_path, result_path, **class_args)

    def run(self, event, *args, **kwargs):
        data = event.body
        if not data:
            return event
        if self._stream:
            self._stream.push(data)
            event.terminated = True
            event.body = None
        return event


class FlowStep(BaseStep):
    """workflow / DAG (states.py:892-1402)"""

    kind = "flow"
    _dict_fields = BaseStep._dict_fields + ["steps", "engine", "default_final_step"]

    def __init__(self, name=None, steps=None, after=None, engine=None, final_step=None):
        super().__init__(name, after)
        self._steps = None
        self.steps = steps
        self.engine = engine
        self.from_step = os.environ.get("START_FROM_STEP", None)
        self.final_step = final_step
        self._last_added = None
        self._controller = None
        self._wait_for_result = False
        self._source = None
        self._start_steps = []
        self._async_flow = None

    def get_children(self):
        return self._steps.values()

    @property
    def steps(self):
        return self._steps

    @steps.setter
    def steps(self, steps):
        self._steps = ObjectDict.from_dict(classes_map, steps, "task")

    @property
    def controller(self):
        return self._controller

    def add_step(self, class_name=None, name=None, handler=None, after=None, before=None, graph_shape=None,
                 function=None, full_event=None, input_path=None, result_path=None, **class_args):
        name, step = params_to_step(
            class_name, name, handler, graph_shape=graph_shape, function=function, full_event=full_event,
            input_path=input_path, result_path=result_path, class_args=class_args,
        )
        for item in after if isinstance(after, list) else [after]:
            self.insert_step(name, step, item, before)
        return step

    def insert_step(self, key, step, after, before=None):
        """states.py:1003-1036"""
        step = self._steps.update(key, step)
        step.set_parent(self)
        if after == "$prev" and len(self._steps) == 1:
            after = None
        previous = ""
        if after:
            if after == "$prev" and self._last_added:
                previous = self._last_added.name
            else:
                if after not in self._steps.keys():
                    raise MLRunInvalidArgumentError(f"cant set after, there is no step named {after}")
                previous = after
            step.after_step(previous)
        if before:
            if before not in self._steps.keys():
                raise MLRunInvalidArgumentError(f"cant set before, there is no step named {before}")
            if before == step.name or before == previous:
                raise GraphError(f"graph loop, step {before} is specified in before and/or after {key}")
            self[step.name].after_step(*self[before].after, append=False)
            self[before].after_step(step.name, append=False)
        self._last_added = step
        return step

    def clear_children(self, steps=None):
        for key in list(steps or self._steps.keys()):
            del self._steps[key]

    def __getitem__(self, name):
        return self._steps[name]

    def __setitem__(self, name, step):
        self.add_step(name, step)

    def __delitem__(self, key):
        del self._steps[key]

    def __iter__(self):
        yield from self._steps.keys()

    def __contains__(self, name):
        return name in self._steps

    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        self.context = context
        self._insert_all_error_handlers()
        self.check_and_process_graph()
        for step in self._steps.values():
            step.set_parent(self)
            step.init_object(context, namespace, mode, reset=reset)
        self._set_error_handler()
        self._post_init(mode)
        if self.engine != "sync":
            self._build_async_flow()
            self._run_async_flow()

    def check_and_process_graph(self, allow_empty=False):
        """validate the DAG and set the .next links (states.py:1073-1184)"""
        if self.is_empty() and allow_empty:
            self._start_steps = []
            return [], None, []

        def find_loop(step, seen):
            for prev in step.after or []:
                if prev in seen:
                    return step.name
                found = find_loop(self[prev], seen + [prev])
                if found:
                    return found
            return None

        start_steps = []
        for step in self._steps.values():
            step._next = None
            step._visited = False
            if step.after:
                loop = find_loop(step, [])
                if loop:
                    raise GraphError(f"Error, loop detected in step {loop}, graph must be acyclic (DAG)")
            else:
                start_steps.append(step.name)

        responders = []
        for step in self._steps.values():
            if getattr(step, "responder", None) and step.kind != "error_step":
                responders.append(step.name)
            if step.on_error and step.on_error in start_steps:
                start_steps.remove(step.on_error)
            for prev in step.after or []:
                self[prev].set_next(step.name)
        if self.on_error and self.on_error in start_steps:
            start_steps.remove(self.on_error)

        if len(responders) > 1:
            raise GraphError(
                f'there are more than one responder steps in the graph ({",".join(responders)})'
            )

        if self.from_step:
            if self.from_step not in self.steps:
                raise GraphError(f"from_step ({self.from_step}) specified and not found in graph steps")
            start_steps = [self.from_step]

        self._start_steps = [self[name] for name in start_steps]

        def first_in_function(step, current):
            if getattr(step, "function", None) and step.function == current:
                return step
            for item in step.next or []:
                found = first_in_function(self[item], current)
                if found:
                    return found
            return None

        current = get_current_function(self.context)
        if current and current != "*":
            new_starts = []
            for start in self._start_steps:
                step = first_in_function(start, current)
                if step:
                    new_starts.append(step)
            if not new_starts:
                raise GraphError(f"did not find steps pointing to current function ({current})")
            self._start_steps = new_starts

        if self.engine == "sync" and len(self._start_steps) > 1:
            raise GraphError("sync engine can only have one starting step (without .after)")

        default_final_step = None
        if self.final_step:
            if self.final_step not in self.steps:
                raise GraphError(f"final_step ({self.final_step}) specified and not found in graph steps")
            default_final_step = self.final_step
        elif len(self._start_steps) == 1:
            cur = self._start_steps[0]
            while cur:
                nxt = cur.next
                if not nxt:
                    default_final_step = cur.name
                    break
                cur = self[nxt[0]] if len(nxt) == 1 else None
        return self._start_steps, default_final_step, responders

    # ---- async engine (storey emulation) ------------------------------------------------------
    def set_flow_source(self, source):
        self._source = source

    def _build_async_flow(self):
        """states.py:1190-1226"""
        self._wait_for_result = _init_async_nodes(self.context, self._steps.values())
        flow = _AsyncFlow(self.context)

        def link(state, node):
            if not state._is_local_function(self.context) or state._visited:
                return
            for item in state.next or []:
                nxt = self[item]
                if getattr(nxt, "_node", None) is not None:
                    node.outlets.append(nxt._node)
                    link(nxt, nxt._node)
            state._visited = True

        for start in self._start_steps:
            if getattr(start, "_node", None) is not None:
                flow.start_nodes.append(start._node)
                link(start, start._node)

        for step in self._steps.values():
            node = getattr(step, "_node", None)
            if (step.on_error or self.on_error) and node is not None:
                err_step = self._steps[step.on_error or self.on_error]
                if step is not err_step and getattr(err_step, "_node", None) is not None:
                    node.recovery = err_step._node
                    for item in err_step.next or []:
                        nxt = self[item]
                        nnode = getattr(nxt, "_node", None)
                        if nnode is not None and nnode not in err_step._node.outlets:
                            err_step._node.outlets.append(nnode)
        self._async_flow = flow

    def _run_async_flow(self):
        self._controller = self._async_flow.run()

    def is_empty(self):
        return len(self.steps) == 0

    def list_child_functions(self):
        out = []
        for step in self.get_children():
            fn = getattr(step, "function", None)
            if fn and fn not in out:
                out.append(fn)
        return out

    def run(self, event, *args, **kwargs):
        """states.py:1279-1323"""
        if self._controller:
            event._awaitable_result = None
            resp = self._controller.emit(event, return_awaitable_result=self._wait_for_result)
            if self._wait_for_result and resp:
                return resp.await_result()
            event = _copy.copy(event)
            event.body = {"id": event.id}
            return event

        if len(self._start_steps) == 0:
            return event
        cur = self._start_steps[0]
        while cur:
            try:
                event = cur.run(event, *args, **kwargs)
            except Exception as exc:
                if self._on_error_handler:
                    self._log_error(event, exc, failed_step=cur.name)
                    event.body = self._call_error_handler(event, exc)
                    event.terminated = True
                    return event
                raise exc
            if getattr(event, "terminated", None):
                return event
            if isinstance(getattr(event, "error", None), dict) and cur.name in event.error:
                cur = self._steps[cur.on_error]
            nxt = cur.next
            if nxt and len(nxt) > 1:
                raise GraphError(
                    f"synchronous flow engine doesnt support branches use async, step={cur.name}"
                )
            cur = self[nxt[0]] if nxt else None
        return event

    def wait_for_completion(self):
        if self._controller:
            self._controller.terminate()
            return self._controller.await_termination()

    def _insert_all_error_handlers(self):
        for name, step in self._steps.items():
            if step.kind == "error_step":
                self._insert_error_step(name, step)

    def _insert_error_step(self, name, step):
        """states.py:1362-1378"""
        if not step.before and not any(step.name in other.after for other in self._steps.values()):
            step.responder = True
            return
        for step_name in step.before:
            if step_name not in self._steps.keys():
                raise MLRunInvalidArgumentError(f"cant set before, there is no step named {step_name}")
            self[step_name].after_step(name)

    def set_flow(self, steps, force=False):
        if not force and self.steps:
            raise MLRunInvalidArgumentError(
                "set_flow() called on a step that already has downstream steps. "
                "If you want to overwrite existing steps, set force=True."
            )
        self.steps = None
        step = self
        for nxt in steps:
            step = step.to(**nxt) if isinstance(nxt, dict) else step.to(nxt)
        return step

    def supports_termination(self):
        return self.engine != "sync"


class RootFlowStep(FlowStep):
    kind = "root"
    _dict_fields = ["steps", "engine", "final_step", "on_error"]


classes_map = {
    "task": TaskStep,
    "router": RouterStep,
    "flow": FlowStep,
    "queue": QueueStep,
    "error_step": ErrorStep,
}


def graph_root_setter(server, graph):
    """states.py:1520-1534"""
    if graph:
        if isinstance(graph, dict):
            kind = graph.get("kind")
        elif hasattr(graph, "kind"):
            kind = graph.kind
        else:
            raise MLRunInvalidArgumentError("graph must be a dict or a valid object")
        if kind == StepKinds.router:
            server._graph = server._verify_dict(graph, "graph", RouterStep)
        elif not kind or kind == StepKinds.root:
            server._graph = server._verify_dict(graph, "graph", RootFlowStep)
        else:
            raise GraphError(f"illegal root step {kind}")


def params_to_step(class_name, name, handler=None, graph_shape=None, function=None, full_event=None,
                   input_path=None, result_path=None, class_args=None):
    """states.py:1548-1619"""
    class_args = class_args or {}
    if class_name and hasattr(class_name, "to_dict"):
        struct = class_name.to_dict()
        kind = struct.get("kind", StepKinds.task)
        name = name or struct.get("name", struct.get("class_name"))
        cls = classes_map.get(kind, RootFlowStep)
        step = cls.from_dict(struct)
        step.function = function
        step.full_event = full_event or step.full_event
        step.input_path = input_path or step.input_path
        step.result_path = result_path or step.result_path
    elif class_name and class_name in queue_class_names:
        if "path" not in class_args:
            raise MLRunInvalidArgumentError("path=<stream path or None> must be specified for queues")
        if not name:
            raise MLRunInvalidArgumentError("queue name must be specified")
        if full_event is not None:
            class_args = class_args.copy()
            class_args["full_event"] = full_event
        step = QueueStep(name, **class_args)
    elif class_name and isinstance(class_name, str) and class_name.startswith("*"):
        routes = class_args.get("routes", None)
        class_name = class_name[1:]
        name = get_name(name, class_name or "router")
        step = RouterStep(class_name, class_args, handler, name=name, function=function, routes=routes,
                          input_path=input_path, result_path=result_path)
    elif class_name or handler:
        name = get_name(name, class_name)
        step = TaskStep(class_name, class_args, handler, name=name, function=function, full_event=full_event,
                        input_path=input_path, result_path=result_path)
    else:
        raise MLRunInvalidArgumentError("class_name or handler must be provided")
    if graph_shape:
        step.shape = graph_shape
    return name, step


# =============================================================================== async emulation
class _Node:
    """one storey step: Map(fn, full_event, input_path, result_path) or Complete"""

    def __init__(self, name, fn=None, full_event=False, input_path=None, result_path=None,
                 pass_context=False, context=None, complete=False):
        self.name = name
        self.fn = fn
        self.full_event = full_event
        self.input_path = input_path
        self.result_path = result_path
        self.pass_context = pass_context
        self.context = context
        self.complete = complete
        self.outlets = []
        self.recovery = None
        self.fullname = name


class _Awaitable:
    def __init__(self):
        self._q = _queue.Queue(1)

    def set(self, value):
        self._q.put(value)

    def await_result(self):
        value = self._q.get()
        if isinstance(value, BaseException):
            raise value
        return value


class _Controller:
    """SyncEmitSource controller contract: emit / terminate / await_termination"""

    def __init__(self, flow):
        self._flow = flow
        self._q = _queue.Queue()
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        self._terminated = False

    def _loop(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            event, awaitable = item
            try:
                self._flow._dispatch_all(event, awaitable)
                if awaitable is not None and not awaitable._done:
                    # flow finished without reaching Complete
                    awaitable._done = True
                    awaitable.set(None)
            except BaseException as exc:  # noqa: BLE001
                if awaitable is not None and not awaitable._done:
                    awaitable._done = True
                    awaitable.set(exc)
                else:
                    self._flow.context.logger.error(f"async flow error: {exc}")

    def emit(self, event, return_awaitable_result=None):
        awaitable = None
        if return_awaitable_result:
            awaitable = _Awaitable()
            awaitable._done = False
        self._q.put((event, awaitable))
        return awaitable

    def terminate(self):
        if not self._terminated:
            self._terminated = True
            self._q.put(None)

    def await_termination(self):
        self._thread.join()
        return None


from .merger import DROP  # noqa: E402


def _branch_copy(event):
    """what an extra outlet receives: storey hands every additional branch its own deep copy of the event, so branches
    never see each other's in-place edits (pinned by tests/serving/test_merger.py:107-128: [10, 11], not [13, 13])"""
    ev = _copy.copy(event)
    ev.body = _copy.deepcopy(event.body)
    return ev


class _AsyncFlow:
    def __init__(self, context):
        self.context = context
        self.start_nodes = []

    def run(self):
        return _Controller(self)

    def _dispatch_all(self, event, awaitable):
        for i, node in enumerate(self.start_nodes):
            ev = event if i == 0 else _copy.copy(event)
            self._run_node(node, ev, awaitable)

    def _run_node(self, node, event, awaitable):
        if node.complete:
            if awaitable is not None and not awaitable._done:
                awaitable._done = True
                awaitable.set(event)
            return
        try:
            if node.full_event:
                kwargs = {"context": node.context} if node.pass_context else {}
                result = node.fn(event, **kwargs)
                if result is DROP:
                    return
                out = result if result is not None else event
            else:
                kwargs = {"context": node.context} if node.pass_context else {}
                element = _extract_input_data(node.input_path, event.body)
                result = node.fn(element, **kwargs)
                if result is DROP:
                    return
                out = _copy.copy(event)
                out.body = _update_result_body(node.result_path, event.body, result)
        except Exception as exc:
            if node.recovery is None:
                raise
            if not getattr(event, "error", None):
                event.error = {}
            event.error[node.name] = err_to_str(exc)
            event.origin_state = node.fullname
            self._run_node(node.recovery, event, awaitable)
            return
        payloads = [out] + [_branch_copy(out) for _ in node.outlets[1:]]  # copied before any branch runs
        for outlet, ev in zip(node.outlets, payloads):
            self._run_node(outlet, ev, awaitable)


def _init_async_nodes(context, steps):
    """build one `_Node` per local step (states.py:1622-1710); returns wait_for_result"""
    wait_for_result = False
    trigger = getattr(context, "trigger", None)
    respond_supported = trigger is None or trigger == "http"

    for step in steps:
        step._node = None
        if not (hasattr(step, "async_object") and step._is_local_function(context)):
            continue
        if step.kind == StepKinds.queue:
            skip_stream = context.is_mock and step.next
            if step.path and not skip_stream:
                stream = step._stream

                def _push(body, _stream=stream):
                    _stream.push(body)
                    return body

                node = _Node(step.name, fn=_push)
            else:
                node = _Node(step.name, fn=lambda x: x)
            step._async_object = node
        else:
            obj = step.async_object
            if obj is not None and hasattr(obj, "_outlets") and hasattr(obj, "do"):
                # native (MapClass-derived) step: its own kwargs decide the call convention
                node = _Node(
                    step.name,
                    fn=obj.do,
                    full_event=bool(getattr(obj, "_full_event", None)),
                    input_path=getattr(obj, "_input_path", None),
                    result_path=getattr(obj, "_result_path", None),
                    context=context,
                )
            else:
                node = _Node(
                    step.name,
                    fn=step._handler,
                    full_event=bool(step.full_event or step._call_with_event),
                    input_path=step.input_path,
                    result_path=step.result_path,
                    pass_context=step._inject_context,
                    context=context,
                )
        node.fullname = step.fullname
        step._node = node
        if respond_supported and not step.next and getattr(step, "responder", None):
            node.outlets.append(_Node("complete", complete=True))
            wait_for_result = True
    return wait_for_result
