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
from .step_io import _extract_input_data, _update_result_body

callable_prefix = "_"
path_splitter = "/"
previous_step = "$prev"
queue_class_names = [">>", "$queue"]
MAX_ALLOWED_STEPS = 4500  # states.py:87


class GraphError(Exception):
    """error in graph topology or configuration (states.py:52-55)"""


class StepKinds:
    router, task, flow, queue, choice, root, error_step = "router", "task", "flow", "queue", "choice", "root", "error_step"


# wire fields of a task-like step, in the order the reference serialises them
_task_step_fields = ("kind class_name class_args handler skip_context after function comment shape full_event on_error "
                     "responder input_path result_path").split()


class MapClass:
    """stand-in for storey.MapClass: the base the feature-store steps derive from.  Holds the
    kwargs storey's flow base keeps (context/name/full_event/input_path/result_path)."""

    def __init__(self, context=None, name=None, full_event=None, input_path=None, result_path=None, **kwargs):
        self.context, self.name = context, name
        self._full_event, self._input_path, self._result_path = full_event, input_path, result_path
        self.logger = getattr(context, "logger", None) if context else None
        self._kwargs = kwargs
        self._outlets = []  # marks a "native" async step (states.py:1678)


def get_current_function(context):
    return (getattr(context, "current_function", None) or "") if context else ""


def get_name(name, class_name):
    """a step is named explicitly, or after its class"""
    if name:
        return name
    if not class_name:
        raise MLRunInvalidArgumentError("name or class_name must be provided")
    return class_name.__name__ if isinstance(class_name, type) else class_name


class BaseStep(ModelObj):
    """what every node of the graph has: a name, a parent, predecessors (`after`), successors (`next`), an optional error
    handler (states.py:102-395)"""

    kind = "BaseStep"
    default_shape = "ellipse"
    _dict_fields = ["kind", "comment", "after", "on_error"]

    def __init__(self, name=None, after=None, shape=None):
        self.name, self.shape = name, shape
        self.after = after or []
        self._parent = self._next = None
        self.comment = self.context = None
        self.on_error = self._on_error_handler = None

    # ---- wiring -----------------------------------------------------------------------------------
    @property
    def next(self):
        return self._next

    @property
    def parent(self):
        return self._parent

    def set_parent(self, parent):
        self._parent = parent

    def set_next(self, key):
        if not self._next:
            self._next = [key]
        elif key not in self._next:
            self._next.append(key)
        return self

    def after_step(self, *after, append=True):
        if not append:
            self.after = []
        for item in after:
            key = item if isinstance(item, str) else item.name
            if key not in self.after:
                self.after.append(key)
        return self

    @property
    def fullname(self):
        own = self.name or ""
        up = self._parent.fullname if self._parent else ""
        return (path_splitter.join([up, own]) if up else own).replace(":", "_")

    def path_to_step(self, path):
        node = self
        for key in (path or "").split(path_splitter):
            if key not in node:
                raise GraphError(f"step {key} doesnt exist in the graph under {node.fullname}")
            node = node[key]
        return node

    def _owner(self):
        """the flow new steps are registered in: this step when it is a flow, else its parent"""
        if hasattr(self, "steps"):
            return self
        if not self._parent:
            raise GraphError(f"step {self.name} parent is not set or it's not part of a graph")
        return self._parent

    def to(self, class_name=None, name=None, handler=None, graph_shape=None, function=None, full_event=None,
           input_path=None, result_path=None, **class_args):
        """append a step after this one (states.py:297-362)"""
        owner = self._owner()
        key, step = params_to_step(class_name, name, handler, graph_shape=graph_shape, function=function,
                                   full_event=full_event, input_path=input_path, result_path=result_path, class_args=class_args)
        step = owner._steps.update(key, step)
        step.set_parent(owner)
        if owner is not self:
            step.after_step(self.name)
        owner._last_added = step
        return step

    def set_flow(self, steps, force=False):
        raise NotImplementedError("set_flow() can only be called on a FlowStep")

    # ---- error routing (states.py:155-231, 263-282) ---------------------------------------------------------
    def error_handler(self, name=None, class_name=None, handler=None, before=None, function=None, full_event=None,
                      input_path=None, result_path=None, **class_args):
        if not (class_name or handler):
            raise MLRunInvalidArgumentError("class_name or handler must be provided")
        if before and isinstance(self, RootFlowStep):
            raise MLRunInvalidArgumentError("`before` arg can't be specified for graph error handler")
        key = get_name(name, class_name)
        catcher = ErrorStep(class_name, class_args, handler, name=key, function=function, full_event=full_event,
                            input_path=input_path, result_path=result_path)
        catcher.before = ([before] if isinstance(before, str) else before) or []
        catcher.base_step = self.name
        self.on_error = key
        flow = self._parent if getattr(self, "_parent", None) else self
        flow._steps.update(key, catcher).set_parent(flow)
        return self

    def _set_error_handler(self):
        if self.on_error:
            self._on_error_handler = self.context.root.path_to_step(self.on_error).run

    def _log_error(self, event, err, **kwargs):
        text, trace = err_to_str(err), traceback.format_exc()
        log = self.context.logger
        log.error(f"step {self.name} got error {text} when processing an event:\n {event.body}")
        log.error(trace)
        self.context.push_error(event, f"{text}\n{trace}", source=self.fullname, **kwargs)

    def _call_error_handler(self, event, err, **kwargs):
        """the failure is recorded on the event, which then continues at the handler step"""
        event.error = event.error or {}
        event.error[self.name] = err_to_str(err)
        event.origin_state = self.fullname
        return self._on_error_handler(event)

    # ---- defaults the subclasses override ----------------------------------------------------------------
    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        self.context = context

    def _post_init(self, mode="sync"):
        pass

    def _is_local_function(self, context):
        return True

    def get_children(self):
        return []

    def __iter__(self):
        yield from []

    def supports_termination(self):
        return False


_INJECTABLE = ("name", "context", "input_path", "result_path", "full_event")


class TaskStep(BaseStep):
    """runs a class or a handler (states.py:398-599)"""

    kind = "task"
    _dict_fields = _task_step_fields
    _default_class = ""

    def __init__(self, class_name=None, class_args=None, handler=None, name=None, after=None, full_event=None,
                 function=None, responder=None, input_path=None, result_path=None):
        super().__init__(name, after)
        self.class_name, self.class_args, self.handler = class_name, class_args or {}, handler
        self.function, self.responder, self.full_event = function, responder, full_event
        self.input_path, self.result_path = input_path, result_path
        self.skip_context = self.context = self.on_error = None
        self._handler = self._object = self._async_object = self._class_object = None
        self._inject_context = self._call_with_event = False

    # ---- resolution: class_name / handler -> the callable run() uses ------------------------------------------
    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        self.context, self._async_object = context, None
        if not self._is_local_function(context):
            return
        if self.handler and not self.class_name:
            self._bind_function(namespace)
            self._set_error_handler()
            return
        self._class_object, self.class_name = self.get_step_class_object(namespace)
        if reset or not self._object:
            self._build_object(namespace, extra_kwargs)
        self._set_error_handler()
        if mode != "skip":
            self._post_init(mode)

    def _bind_function(self, namespace):
        """a bare handler: a callable, a name to look up, or a "(expr)" lambda body (helpers.get_function)"""
        if callable(self.handler):
            self._handler, self.handler = self.handler, self.handler.__name__
        else:
            self._handler = get_function(self.handler, namespace)
        try:
            accepted = list(signature(self._handler).parameters)
        except (TypeError, ValueError):
            accepted = []
        self._inject_context = "context" in accepted

    def _build_object(self, namespace, extra_kwargs):
        ctor_args = self.get_full_class_args(namespace, self._class_object, **extra_kwargs)
        try:
            self._object = self._class_object(**ctor_args)
        except TypeError as exc:
            raise TypeError(f"failed to init step {self.name}\n args={self.class_args}") from exc
        method = self.handler
        if method:
            if not hasattr(self._object, method):
                raise GraphError(f"handler ({method}) specified but doesnt exist in class {self.class_name}")
        elif hasattr(self._object, "do_event"):  # model servers / routers: they take the whole event
            method, self._call_with_event = "do_event", True
        elif hasattr(self._object, "do"):
            method = "do"
        if method:
            self._handler = getattr(self._object, method, None)

    def get_full_class_args(self, namespace, class_object, **extra_kwargs):
        """states.py:494-512: `_x` args resolve to callables; name/context/... only when accepted"""
        args = {}
        for key, value in self.class_args.items():
            if key.startswith(callable_prefix):
                key, value = key[1:], get_function(value, namespace)
            args[key] = value
        args.update(extra_kwargs)
        spec = getfullargspec(class_object)
        takes = (lambda k: True) if spec.varkw else (lambda k: k in spec.args)
        args.update({k: getattr(self, k) for k in _INJECTABLE if takes(k)})
        if takes("graph_step"):
            args["graph_step"] = self
        return args

    def get_step_class_object(self, namespace):
        if isinstance(self.class_name, type):
            return self.class_name, self.class_name.__name__
        found = self._class_object or get_class(self.class_name or self._default_class, namespace)
        return found, self.class_name

    def _is_local_function(self, context):
        """states.py:529-540 -- does this step run in the current function of a multi-function graph?"""
        here, mine = get_current_function(context), self.function
        if here == "*" or (not mine and not here):
            return True
        return bool(mine and mine == "*") or mine == here

    @property
    def async_object(self):
        return self._async_object or self._object

    def clear_object(self):
        self._object = None

    def _post_init(self, mode="sync"):
        hook = getattr(self._object, "post_init", None) if self._object else None
        if hook:
            hook(mode)

    def respond(self):
        self.responder = True
        return self

    # ---- per-event call convention (states.py:564-599) --------------------------------------------------------
    def run(self, event, *args, **kwargs):
        if not self._is_local_function(self.context):
            return event
        if self._inject_context:
            kwargs["context"] = self.context
        elif kwargs:
            kwargs.pop("context", None)
        try:
            if self.full_event or self._call_with_event:
                return self._handler(event, *args, **kwargs)  # the handler's return value IS the result event
            if self._handler is None:
                raise MLRunInvalidArgumentError(f"step {self.name} does not have a handler")
            outcome = self._handler(_extract_input_data(self.input_path, event.body), *args, **kwargs)
            event.body = _update_result_body(self.result_path, event.body, outcome)
        except Exception as exc:
            if not self._on_error_handler:
                raise exc
            self._log_error(event, exc)
            outcome = self._call_error_handler(event, exc)  # states.py:592-597: the handler runs BEFORE event.body is read again
            event.body = _update_result_body(self.result_path, event.body, outcome)
        return event


class ErrorStep(TaskStep):
    """a task that other steps name in `on_error`"""

    kind = "error_step"
    _dict_fields = _task_step_fields + ["before", "base_step"]
    _default_class = ""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.before = self.base_step = None


class RouterStep(TaskStep):
    """a task whose object dispatches to child routes (states.py:671-798)"""

    kind = "router"
    default_shape = "doubleoctagon"
    _dict_fields = _task_step_fields + ["routes"]
    _default_class = "mlrun.serving.ModelRouter"

    def __init__(self, class_name=None, class_args=None, handler=None, routes=None, name=None, function=None,
                 input_path=None, result_path=None):
        super().__init__(class_name, class_args, handler, name=name, function=function, input_path=input_path,
                         result_path=result_path)
        self._routes = None
        self.routes = routes

    @property
    def routes(self):
        return self._routes

    @routes.setter
    def routes(self, routes):
        self._routes = ObjectDict.from_dict(classes_map, routes, "task")

    def get_children(self):
        return self._routes.values()

    def add_route(self, key, route=None, class_name=None, handler=None, function=None, **class_args):
        if not (route or class_name or handler):
            raise MLRunInvalidArgumentError("route or class_name must be specified")
        route = route or TaskStep(class_name, class_args, handler=handler)
        route.function = function or route.function
        if len(self._routes) >= MAX_ALLOWED_STEPS:  # states.py:740-743
            raise MLRunInvalidArgumentError(
                f"Cannot create the serving graph: the maximum number of steps is {MAX_ALLOWED_STEPS}")
        route = self._routes.update(key, route)
        route.set_parent(self)
        return route

    def clear_children(self, routes=None):
        for key in list(routes or self._routes.keys()):
            del self._routes[key]

    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        """the router object first (it receives the routes), then every child, then the router's post_init"""
        if not self._is_local_function(context):
            return
        self.class_args = self.class_args or {}
        super().init_object(context, namespace, "skip", reset=reset, routes=self._routes, **extra_kwargs)
        for child in self._routes.values():
            child.function = child.function or self.function or child.function
            child.set_parent(self)
            child.init_object(context, namespace, mode, reset=reset)
        self._set_error_handler()
        self._post_init(mode)

    # mapping protocol over the routes
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
        self.path, self.shards, self.retention_in_hours = path, shards, retention_in_hours
        self.trigger_args, self.options = trigger_args, options
        self._stream = self._async_object = None

    @property
    def async_object(self):
        return self._async_object

    def init_object(self, context, namespace, mode="sync", reset=False, **extra_kwargs):
        self.context = context
        if self.path:
            from .host import get_stream_pusher

            self._stream = get_stream_pusher(self.path, **self.options)
        self._set_error_handler()

    def to(self, class_name=None, name=None, handler=None, graph_shape=None, function=None,
           full_event=None, input_path=None, result_path=None, **class_args):
        if not function:  # what follows a queue runs in another function: it has to say which (states.py:860-864)
            raise MLRunInvalidArgumentError(
                f"step '{get_name(name, class_name)}' must specify a function, because it follows a queue step")
        return super().to(class_name, name, handler, graph_shape, function, full_event, input_path, result_path, **class_args)

    def run(self, event, *args, **kwargs):
        payload = event.body
        if payload and self._stream:
            self._stream.push(payload)
            event.terminated, event.body = True, None
        return event


class FlowStep(BaseStep):
    """workflow / DAG of steps (states.py:892-1402): wiring, validation, and the two executors -- the synchronous walk
    of a single chain and the hand-off to the async engine"""

    kind = "flow"
    _dict_fields = BaseStep._dict_fields + ["steps", "engine", "default_final_step"]

    def __init__(self, name=None, steps=None, after=None, engine=None, final_step=None):
        super().__init__(name, after)
        self._steps = None
        self.steps = steps
        self.engine, self.final_step = engine, final_step
        self.from_step = os.environ.get("START_FROM_STEP", None)
        self._last_added = self._controller = self._source = self._async_flow = None
        self._wait_for_result = False
        self._start_steps = []

    # ---- container protocol -------------------------------------------------------------------------
    @property
    def steps(self):
        return self._steps

    @steps.setter
    def steps(self, steps):
        self._steps = ObjectDict.from_dict(classes_map, steps, "task")

    @property
    def controller(self):
        return self._controller

    def get_children(self):
        return self._steps.values()

    def is_empty(self):
        return len(self.steps) == 0

    def clear_children(self, steps=None):
        for key in list(steps or self._steps.keys()):
            del self._steps[key]

    def list_child_functions(self):
        seen = []
        for child in self.get_children():
            fn = getattr(child, "function", None)
            if fn and fn not in seen:
                seen.append(fn)
        return seen

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

    # ---- building ------------------------------------------------------------------------------------
    def add_step(self, class_name=None, name=None, handler=None, after=None, before=None, graph_shape=None,
                 function=None, full_event=None, input_path=None, result_path=None, **class_args):
        key, step = params_to_step(class_name, name, handler, graph_shape=graph_shape, function=function,
                                   full_event=full_event, input_path=input_path, result_path=result_path, class_args=class_args)
        for predecessor in (after if isinstance(after, list) else [after]):
            self.insert_step(key, step, predecessor, before)
        return step

    def _must_exist(self, key, what):
        if key not in self._steps.keys():
            raise MLRunInvalidArgumentError(f"cant set {what}, there is no step named {key}")

    def insert_step(self, key, step, after, before=None):
        """states.py:1003-1036 -- register `step`, hang it after `after` ("$prev": the last added) and before `before`"""
        step = self._steps.update(key, step)
        step.set_parent(self)
        if after == previous_step and len(self._steps) == 1:
            after = None
        hung_after = ""
        if after:
            if after == previous_step and self._last_added:
                hung_after = self._last_added.name
            else:
                self._must_exist(after, "after")
                hung_after = after
            step.after_step(hung_after)
        if before:
            self._must_exist(before, "before")
            if before in (step.name, hung_after):
                raise GraphError(f"graph loop, step {before} is specified in before and/or after {key}")
            successor = self[before]
            self[step.name].after_step(*successor.after, append=False)  # the new step inherits the successor's inputs
            successor.after_step(step.name, append=False)
        self._last_added = step
        return step

    def set_flow(self, steps, force=False):
        if self.steps and not force:
            raise MLRunInvalidArgumentError(
                "set_flow() called on a step that already has downstream steps. "
                "If you want to overwrite existing steps, set force=True.")
        self.steps = None
        tail = self
        for spec in steps:
            tail = tail.to(**spec) if isinstance(spec, dict) else tail.to(spec)
        return tail

    def _insert_all_error_handlers(self):
        for key, step in self._steps.items():
            if step.kind == StepKinds.error_step:
                self._insert_error_step(key, step)

    def _insert_error_step(self, name, step):
        """states.py:1362-1378 -- an error step nothing follows answers the caller; else it precedes its `before` steps"""
        followed = any(step.name in other.after for other in self._steps.values())
        if not step.before and not followed:
            step.responder = True
            return
        for target in step.before:
            self._must_exist(target, "before")
            self[target].after_step(name)

    # ---- validation (states.py:1073-1184) -----------------------------------------------------------------
    def _cycle_through(self, step, trail=()):
        """name of a step that closes a cycle over `after` links, if any"""
        for prev in step.after or []:
            if prev in trail:
                return step.name
            found = self._cycle_through(self[prev], trail + (prev,))
            if found:
                return found
        return None

    def _first_owned(self, step, function):
        """first step on the way down from `step` that runs in `function`"""
        if getattr(step, "function", None) and step.function == function:
            return step
        for key in step.next or []:
            found = self._first_owned(self[key], function)
            if found:
                return found
        return None

    def check_and_process_graph(self, allow_empty=False):
        """validate the DAG and set the .next links -> (start steps, default final step, responders)"""
        if allow_empty and self.is_empty():
            self._start_steps = []
            return [], None, []

        starts = []
        for step in self._steps.values():
            step._next, step._visited = None, False
            if not step.after:
                starts.append(step.name)
                continue
            looped = self._cycle_through(step)
            if looped:
                raise GraphError(f"Error, loop detected in step {looped}, graph must be acyclic (DAG)")

        responders = []
        for step in self._steps.values():
            if getattr(step, "responder", None) and step.kind != StepKinds.error_step:
                responders.append(step.name)
            if step.on_error and step.on_error in starts:
                starts.remove(step.on_error)  # an error handler is entered on failure only
            for prev in step.after or []:
                self[prev].set_next(step.name)
        if self.on_error and self.on_error in starts:
            starts.remove(self.on_error)
        if len(responders) > 1:
            raise GraphError(f'there are more than one responder steps in the graph ({",".join(responders)})')

        if self.from_step:
            if self.from_step not in self.steps:
                raise GraphError(f"from_step ({self.from_step}) specified and not found in graph steps")
            starts = [self.from_step]
        self._start_steps = [self[key] for key in starts]

        here = get_current_function(self.context)
        if here and here != "*":  # a child function of a multi-function graph starts at its own first steps
            owned = [s for s in (self._first_owned(start, here) for start in self._start_steps) if s]
            if not owned:
                raise GraphError(f"did not find steps pointing to current function ({here})")
            self._start_steps = owned

        if self.engine == "sync" and len(self._start_steps) > 1:
            raise GraphError("sync engine can only have one starting step (without .after)")
        return self._start_steps, self._default_final(), responders

    def _default_final(self):
        if self.final_step:
            if self.final_step not in self.steps:
                raise GraphError(f"final_step ({self.final_step}) specified and not found in graph steps")
            return self.final_step
        if len(self._start_steps) != 1:
            return None
        node = self._start_steps[0]
        while node:  # follow the chain while it does not branch
            if not node.next:
                return node.name
            node = self[node.next[0]] if len(node.next) == 1 else None
        return None

    # ---- initialisation ---------------------------------------------------------------------------------
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

    # ---- async engine (storey emulation) ------------------------------------------------------
    def set_flow_source(self, source):
        self._source = source

    def _build_async_flow(self):
        """states.py:1190-1226"""
        self._wait_for_result = _init_async_nodes(self.context, self._steps.values())
        flow = _AsyncFlow(self.context)

        def link(state, node):
            if state._visited or not state._is_local_function(self.context):
                return
            for key in state.next or []:
                follower = self[key]
                if getattr(follower, "_node", None) is not None:
                    node.outlets.append(follower._node)
                    link(follower, follower._node)
            state._visited = True

        for start in self._start_steps:
            if getattr(start, "_node", None) is not None:
                flow.start_nodes.append(start._node)
                link(start, start._node)

        for step in self._steps.values():
            node, target = getattr(step, "_node", None), step.on_error or self.on_error
            if node is None or not target:
                continue
            catcher = self._steps[target]
            cnode = getattr(catcher, "_node", None)
            if catcher is step or cnode is None:
                continue
            node.recovery = cnode
            for key in catcher.next or []:
                fnode = getattr(self[key], "_node", None)
                if fnode is not None and fnode not in cnode.outlets:
                    cnode.outlets.append(fnode)
        self._async_flow = flow

    def _run_async_flow(self):
        self._controller = self._async_flow.run()

    def wait_for_completion(self):
        if self._controller:
            self._controller.terminate()
            return self._controller.await_termination()

    def supports_termination(self):
        return self.engine != "sync"

    # ---- execution (states.py:1279-1323) -------------------------------------------------------------------
    def run(self, event, *args, **kwargs):
        if self._controller:
            return self._run_async(event)
        if not self._start_steps:
            return event
        step = self._start_steps[0]
        while step:
            try:
                event = step.run(event, *args, **kwargs)
            except Exception as exc:
                if not self._on_error_handler:
                    raise exc
                self._log_error(event, exc, failed_step=step.name)  # graph-level handler: answer and stop
                event.body = self._call_error_handler(event, exc)
                event.terminated = True
                return event
            if getattr(event, "terminated", None):
                return event
            failed_here = isinstance(getattr(event, "error", None), dict) and step.name in event.error
            if failed_here:
                step = self._steps[step.on_error]  # the step's own handler ran: continue after it
            followers = step.next
            if followers and len(followers) > 1:
                raise GraphError(f"synchronous flow engine doesnt support branches use async, step={step.name}")
            step = self[followers[0]] if followers else None
        return event

    def _run_async(self, event):
        event._awaitable_result = None
        pending = self._controller.emit(event, return_awaitable_result=self._wait_for_result)
        if self._wait_for_result and pending:
            return pending.await_result()
        receipt = _copy.copy(event)
        receipt.body = {"id": receipt.id}
        return receipt


class RootFlowStep(FlowStep):
    """the flow at the top of a function's graph"""

    kind = "root"
    _dict_fields = ["steps", "engine", "final_step", "on_error"]


classes_map = {StepKinds.task: TaskStep, StepKinds.router: RouterStep, StepKinds.flow: FlowStep,
               StepKinds.queue: QueueStep, StepKinds.error_step: ErrorStep}


def graph_root_setter(server, graph):
    """states.py:1520-1534 -- a function's graph is a router or a (root) flow, given as object or wire dict"""
    if not graph:
        return
    if isinstance(graph, dict):
        kind = graph.get("kind")
    elif hasattr(graph, "kind"):
        kind = graph.kind
    else:
        raise MLRunInvalidArgumentError("graph must be a dict or a valid object")
    if kind not in (None, "", StepKinds.root, StepKinds.router):
        raise GraphError(f"illegal root step {kind}")
    server._graph = server._verify_dict(graph, "graph", RouterStep if kind == StepKinds.router else RootFlowStep)


def _step_from_object(obj, name, function, full_event, input_path, result_path):
    """an object with to_dict() (a step / model / router instance) travels as its wire form"""
    wire = obj.to_dict()
    step = classes_map.get(wire.get("kind", StepKinds.task), RootFlowStep).from_dict(wire)
    step.function = function
    step.full_event = full_event or step.full_event
    step.input_path = input_path or step.input_path
    step.result_path = result_path or step.result_path
    return name or wire.get("name", wire.get("class_name")), step


def params_to_step(class_name, name, handler=None, graph_shape=None, function=None, full_event=None,
                   input_path=None, result_path=None, class_args=None):
    """states.py:1548-1619 -- the `to()` / `add_step()` arguments -> (step name, step): an object, a queue marker (">>",
    "$queue"), a "*RouterClass", or a task class / handler"""
    class_args = class_args or {}
    if class_name and hasattr(class_name, "to_dict"):
        name, step = _step_from_object(class_name, name, function, full_event, input_path, result_path)
    elif class_name and class_name in queue_class_names:
        if "path" not in class_args:
            raise MLRunInvalidArgumentError("path=<stream path or None> must be specified for queues")
        if not name:
            raise MLRunInvalidArgumentError("queue name must be specified")
        extra = {"full_event": full_event} if full_event is not None else {}
        step = QueueStep(name, **{**class_args, **extra})
    elif class_name and isinstance(class_name, str) and class_name.startswith("*"):
        router_class = class_name[1:]
        name = get_name(name, router_class or "router")
        step = RouterStep(router_class, class_args, handler, name=name, function=function,
                          routes=class_args.get("routes", None), input_path=input_path, result_path=result_path)
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
