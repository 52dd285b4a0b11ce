"""Implementation-agnostic serving scenarios.

Every scenario takes an `api` namespace (see `tests/apis.py`) that exposes the reference's plugin
surface -- `new_function`, `V2ModelServer`, `VotingEnsemble`, `MockEvent`, the feature-store steps
... -- and returns a JSON-serialisable result.  The same scenarios are run through

  * the REAL reference (`/root/reference`, imported with mocked third-party deps) by
    `tests/golden/gen_golden.py`  -> `tests/golden/scenarios.json`   (committed fixture),
  * the CPU oracle (`oracle/`)                                           (-m "not gpu"),
  * the product host layer (`mlrun_b200`)                                 (-m "not gpu" for pure
    host-logic scenarios, -m gpu for scenarios whose steps lower to CUDA).

Each scenario mirrors a reference test; the docstring cites it and, where the reference test
asserts a literal, `EXPECT` holds that literal too (checked for every implementation).
"""

import copy
import json
import math

import numpy as np

TESTDATA = '{"inputs": [5]}'
TESTDATA_2 = '{"inputs": [5, 5]}'


# --------------------------------------------------------------------------- helper classes
def make_namespace(api):
    """step/model classes the reference tests define (tests/serving/test_serving.py:159-215,
    tests/serving/demo_states.py, tests/serving/test_flow.py:36-61), built on `api`'s base classes"""
    V2 = api.V2ModelServer

    class ModelTestingClass(V2):
        def load(self):
            pass

        def predict(self, request):
            return request["inputs"][0] * self.get_param("multiplier")

        def explain(self, request):
            return {"explained": request["inputs"][0]}

        def op_myop(self, event):
            return event.body

    class EnsembleModelTestingClass(ModelTestingClass):
        def predict(self, request):
            return {"predictions": [x * self.get_param("multiplier") for x in request["inputs"]]}

    class EnsembleModelTestingClassClassification(ModelTestingClass):
        def predict(self, request):
            return {"predictions": [self.get_param("predict") for _ in request["inputs"]]}

    class RaiserTestingClass(V2):
        def load(self):
            pass

        def predict(self, request):
            raise ValueError("simulated error..")

    class EchoModel(V2):
        def load(self):
            pass

        def predict(self, request):
            return request["inputs"]

    class ModelClass(V2):
        def load(self):
            pass

        def predict(self, request):
            return request["inputs"][0] * self.get_param("multiplier", 1)

    class ModelClassList(V2):
        def load(self):
            pass

        def predict(self, request):
            return [request["inputs"][0][0] * self.get_param("multiplier", 1)]

    class TrackedModel(V2):
        def load(self):
            pass

        def predict(self, request):
            m = self.get_param("multiplier", 1)
            return np.array([v[0] * m for v in request["inputs"]])

    class BaseClass:
        def __init__(self, context, name=None):
            self.context = context
            self.name = name

    class Echo(BaseClass):
        def __init__(self, name=None):
            self.name = name

        def do(self, x):
            return x

    class RespName(BaseClass):
        def __init__(self, **kwargs):
            self.name = kwargs.get("name")

        def do(self, x):
            return [x, self.name]

    class EchoError(BaseClass):
        def do(self, x):
            x.body = {"body": x.body, "origin_state": x.origin_state, "error": x.error}
            return x

    class Chain(BaseClass):
        def do(self, x):
            x = copy.copy(x)
            x.append(self.name)
            return x

    class ChainWithContext(BaseClass):
        def do(self, x):
            visits = self.context.visits.get(self.name, 0)
            self.context.visits[self.name] = visits + 1
            x = copy.copy(x)
            x.append(self.name)
            return x

    class Raiser:
        def __init__(self, msg="", context=None, name=None):
            self.context = context
            self.name = name
            self.msg = msg

        def do(self, x):
            raise ValueError(f" this is an error, {x}")

    class ParEcho:
        def __init__(self, context, name=None, data=None):
            self.context = context
            self.name = name
            self.data = data or {}

        def do(self, x):
            return self.data

    class Mul(api.MapClass):
        def __init__(self, **kwargs):
            super().__init__(**kwargs)

        def do(self, event):
            return event * 2

    def my_hnd(event):
        return {"mul": event["x"] * 2}

    def multiply_input(request):
        request["inputs"][0] = request["inputs"][0] * 2
        return request

    def myfunc1(x, context=None):
        assert isinstance(context, api.GraphContext), "didnt get a valid context"
        return x * 2

    def myfunc2(x):
        return x * 2

    def return_type(event):
        return event.__class__.__name__

    def extract_meta(event):
        event.body = {"id": event.id, "key": event.key}
        return event

    def double(x):  # tests/serving/test_merger.py:27-42
        return x * 2

    class Adder:
        def __init__(self, add=1, **kwargs):
            self.add = add

        def do(self, x):
            return x + self.add

    ns = dict(locals())
    ns.pop("api")
    ns.pop("V2")
    ns.pop("BaseClass")
    for obj in ns.values():
        if isinstance(obj, type):
            # serialise as a bare class name (like the reference tests' module-level classes) so that
            # `to_dict()`-built steps resolve through the namespace passed to the server
            obj.__module__ = "__main__"
            obj.__qualname__ = obj.__name__
    ns["json"] = json
    return ns


def _routes(api, model_class, arg, values):
    names = ["m1", "m2", "m3:v1", "m3:v2"]
    return {n: api.TaskStep(model_class, class_args={"model_path": "", arg: v}) for n, v in zip(names, values)}


def _spec(graph, mode="sync", params=None):
    return {"version": "v2", "parameters": params or {}, "graph": graph, "load_mode": mode,
            "verbose": True, "function_uri": "default/func"}


def _clean(resp, drop=("id", "timestamp")):
    """drop volatile keys (ids, timestamps) so results are comparable across runs"""
    if isinstance(resp, dict):
        return {k: _clean(v, drop) for k, v in resp.items() if k not in drop}
    if isinstance(resp, list):
        return [_clean(v, drop) for v in resp]
    if isinstance(resp, np.ndarray):
        return resp.tolist()
    if isinstance(resp, np.generic):
        return resp.item()
    if isinstance(resp, float) and math.isnan(resp):
        return "NaN"
    if isinstance(resp, bytes):
        return resp.decode()
    return resp


def _resp(resp):
    """normalise a handler result: Response objects -> {status, body}; python objects as-is"""
    if hasattr(resp, "status_code"):
        body = resp.body
        if isinstance(body, (bytes, str)):
            try:
                body = json.loads(body)
            except Exception:
                body = body.decode() if isinstance(body, bytes) else body
        return {"status": resp.status_code, "body": _clean(body)}
    return _clean(resp)


def _first_line(text):
    return str(text).split("\n")[0]


# =========================================================================== router scenarios
def router_protocol(api):
    """tests/serving/test_serving.py:231-239, 428-456, 504-509, 567-609 -- ModelRouter over a spec"""
    ns = make_namespace(api)
    router = api.RouterStep()
    router.routes = _routes(api, "ModelTestingClass", "multiplier", [100, 200, 300, 400])
    ctx = api.init_from_spec(_spec(router.to_dict()), ns)
    out = {}

    def call(key, *args, **kw):
        ev = api.MockEvent(*args, **kw)
        resp = ctx.mlrun_handler(ctx, ev)
        out[key] = _resp(resp)
        return ev, resp

    call("get_models", "", path="/v2/models/", method="GET")
    for url in ["m1", "m2", "m3/versions/v1", "m3/versions/v2"]:
        call(f"infer_{url}", TESTDATA, path=f"/v2/models/{url}/infer")
    call("stream_body_model", '{"model": "m2", "inputs": [5]}')
    call("stream_body_op", '{"model": "m3:v2", "operation": "explain", "inputs": [5]}', path="")
    call("explain", TESTDATA, path="/v2/models/m1/explain")
    call("custom_op", '{"test": "ok"}', path="/v2/models/m1/myop")
    call("bad_op", '{"test": "ok"}', path="/v2/models/m1/xx")
    call("bad_model", '{"test": "ok"}', path="/v2/models/m5/xx")
    ev, resp = call("ready", "", path="/v2/models/m1/ready", method="GET")
    out["ready_text_ok"] = resp.body.decode("utf-8") == f"Model m1 is ready (event_id = {ev.id})"
    out["ready"]["body"] = "<text>"
    call("health_root", None, path="/", method="GET")
    call("health", "", path="/v2/health", method="GET")
    call("bad_prefix", TESTDATA, path="/v3/models/m1/infer")
    out["bad_op"]["body"] = _first_line(out["bad_op"]["body"])
    out["bad_model"]["body"] = _first_line(out["bad_model"]["body"])
    out["bad_prefix"]["body"] = _first_line(out["bad_prefix"]["body"])
    return out


router_protocol.EXPECT = {
    ("infer_m1", "body", "outputs"): 500,
    ("infer_m2", "body", "outputs"): 1000,
    ("infer_m3/versions/v1", "body", "outputs"): 1500,
    ("infer_m3/versions/v2", "body", "outputs"): 2000,
    ("stream_body_model", "body", "outputs"): 1000,
    ("ready", "status"): 200,
    ("bad_op", "status"): 400,
    ("bad_model", "status"): 400,
}


def router_raised_error(api):
    """tests/serving/test_serving.py:459-468 -- predict raises -> 400"""
    ns = make_namespace(api)
    spec = _spec({"kind": "router", "routes": {"m6": {"class_name": "RaiserTestingClass",
                                                     "class_args": {"model_path": "."}}}})
    ctx = api.init_from_spec(spec, ns)
    resp = ctx.mlrun_handler(ctx, api.MockEvent(TESTDATA, path="/v2/models/m6/infer"))
    r = _resp(resp)
    r["body"] = _first_line(r["body"])
    return r


router_raised_error.EXPECT = {("status",): 400}


def ensemble_regression(api):
    """tests/serving/test_serving.py:324-353 -- 4 models x{100..400} on 5 => 1250.0; batch of 2"""
    ns = make_namespace(api)
    out = {}
    for executor in ["array", "thread"]:
        ens = api.RouterStep(class_name="mlrun.serving.routers.VotingEnsemble",
                             class_args={"vote_type": "regression", "prediction_col_name": "predictions",
                                         "format_response_with_col_name_flag": True, "executor_type": executor})
        ens.routes = _routes(api, "EnsembleModelTestingClass", "multiplier", [100, 200, 300, 400])
        ctx = api.init_from_spec(_spec(ens.to_dict()), ns)
        for url in ["m1", "m2", "m3/versions/v1", "m3/versions/v2", "VotingEnsemble", ""]:
            path = f"/v2/models/{url}/infer" if url else "/v2/models/infer"
            r1 = ctx.mlrun_handler(ctx, api.MockEvent(TESTDATA, path=path, method="POST"))
            r2 = ctx.mlrun_handler(ctx, api.MockEvent(TESTDATA_2, path=path))
            out[f"{executor}:{url}"] = [_resp(r1), _resp(r2)]
    return out


ensemble_regression.EXPECT = {
    ("array:VotingEnsemble", 0, "body", "outputs"): {"predictions": [1250.0]},
    ("array:", 1, "body", "outputs"): {"predictions": [1250.0, 1250.0]},
    ("thread:", 0, "body", "outputs"): {"predictions": [1250.0]},
    ("array:m1", 0, "body", "outputs"): {"predictions": [500]},
}


def ensemble_classification(api):
    """tests/serving/test_serving.py:356-388 -- predictions {1,2,3,4}, equal weights => 1 (first max)"""
    ns = make_namespace(api)
    out = {}
    for executor in ["array", "thread"]:
        ens = api.RouterStep(class_name="mlrun.serving.routers.VotingEnsemble",
                             class_args={"vote_type": "classification", "prediction_col_name": "predictions",
                                         "format_response_with_col_name_flag": True, "executor_type": executor})
        ens.routes = _routes(api, "EnsembleModelTestingClassClassification", "predict", [1, 2, 3, 4])
        ctx = api.init_from_spec(_spec(ens.to_dict()), ns)
        for url in ["m1", "m3/versions/v2", "VotingEnsemble", ""]:
            path = f"/v2/models/{url}/infer" if url else "/v2/models/infer"
            r1 = ctx.mlrun_handler(ctx, api.MockEvent(TESTDATA, path=path))
            r2 = ctx.mlrun_handler(ctx, api.MockEvent(TESTDATA_2, path=path))
            out[f"{executor}:{url}"] = [_resp(r1), _resp(r2)]
    return out


ensemble_classification.EXPECT = {
    ("array:VotingEnsemble", 0, "body", "outputs"): {"predictions": [1]},
    ("thread:", 1, "body", "outputs"): {"predictions": [1, 1]},
}


def ensemble_weights(api):
    """tests/serving/test_serving.py:391-425 -- weights {.1,.2,.3,.4}: regression 1500.0, classification 4"""
    ns = make_namespace(api)
    out = {}
    weights = {"m1": 0.1, "m2": 0.2, "m3:v1": 0.3, "m3:v2": 0.4}
    for vote_type, cls, arg, vals in [
        ("regression", "EnsembleModelTestingClass", "multiplier", [100, 200, 300, 400]),
        ("classification", "EnsembleModelTestingClassClassification", "predict", [1, 2, 3, 4]),
    ]:
        ens = api.RouterStep(class_name="mlrun.serving.routers.VotingEnsemble",
                             class_args={"vote_type": vote_type, "prediction_col_name": "predictions",
                                         "format_response_with_col_name_flag": True, "weights": weights,
                                         "executor_type": "array"})
        ens.routes = _routes(api, cls, arg, vals)
        ctx = api.init_from_spec(_spec(ens.to_dict()), ns)
        for url in ["VotingEnsemble", ""]:
            path = f"/v2/models/{url}/infer" if url else "/v2/models/infer"
            r1 = ctx.mlrun_handler(ctx, api.MockEvent(TESTDATA, path=path))
            r2 = ctx.mlrun_handler(ctx, api.MockEvent(TESTDATA_2, path=path))
            out[f"{vote_type}:{url}"] = [_resp(r1), _resp(r2)]
    return out


ensemble_weights.EXPECT = {
    ("regression:", 0, "body", "outputs"): {"predictions": [1500.0]},
    ("classification:VotingEnsemble", 1, "body", "outputs"): {"predictions": [4, 4]},
}


def ensemble_metadata(api):
    """tests/serving/test_serving.py:242-321 -- model list, per-model metadata, weights echo"""
    ns = make_namespace(api)
    out = {}
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("router", api.VotingEnsemble(vote_type="regression", prediction_col_name="predictions"))
    graph.routes = _routes(api, "EnsembleModelTestingClass", "multiplier", [100, 200, 300, 400])
    server = fn.to_mock_server(namespace=ns)
    out["models"] = _clean(server.test("/v2/models/"))
    out["m1"] = _clean(server.test("/v2/models/m1"))
    out["m3v2"] = _clean(server.test("/v2/models/m3/versions/v2"))
    out["ens"] = _clean(server.test("/v2/models/VotingEnsemble"))

    models = ["m1", "m2", "m3:v1", "m3:v2"]
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("router", api.VotingEnsemble(vote_type="regression", prediction_col_name="predictions",
                                                        format_response_with_col_name_flag=True,
                                                        weights=dict(zip(models, [1, 1, 1, 1]))))
    graph.routes = _routes(api, "EnsembleModelTestingClass", "multiplier", [100, 200, 300, 400])
    server = fn.to_mock_server(namespace=ns)
    out["weights_ones"] = _clean(server.test("/v2/models/"))
    # [1,1,1,1] are used as given => the "mean" is a sum (SURVEY App.A item 20)
    out["ones_infer"] = _clean(server.test("/v2/models/infer", body={"inputs": [5]}))
    fn.spec.graph.class_args["weights"] = dict(zip(models, [0.1, 0.2, 0.3, 0.4]))
    server = fn.to_mock_server(namespace=ns)
    out["weights_frac"] = _clean(server.test("/v2/models/"))
    return out


ensemble_metadata.EXPECT = {
    ("m1",): {"name": "m1", "version": "", "inputs": [], "outputs": []},
    ("m3v2",): {"name": "m3", "version": "v2", "inputs": [], "outputs": []},
    ("ens",): {"name": "VotingEnsemble", "version": "v1", "inputs": [], "outputs": []},
    ("weights_ones", "weights"): {"m1": 1, "m2": 1, "m3:v1": 1, "m3:v2": 1},
}


def ensemble_weight_sum_below_one(api):
    """SURVEY §3.4 quirk (serving/routers.py:962-980): weights summing < 1 crash at init"""
    ns = make_namespace(api)
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("router", api.VotingEnsemble(vote_type="regression",
                                                        weights={"m1": 0.1, "m2": 0.2}))
    graph.routes = _routes(api, "EnsembleModelTestingClass", "multiplier", [100, 200, 300, 400])
    try:
        fn.to_mock_server(namespace=ns)
        return {"raised": None}
    except Exception as exc:  # noqa: BLE001
        return {"raised": exc.__class__.__name__}


def ensemble_vote_type_inference(api):
    """serving/routers.py:756-775 -- vote type is inferred once from the first request and sticks"""
    ns = make_namespace(api)
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("router", api.VotingEnsemble(executor_type="array"))
    for k, m in [("a", 1), ("b", 1), ("c", 2)]:
        graph.add_route(k, class_name="EchoTimes", model_path="", multiplier=m)

    class EchoTimes(api.V2ModelServer):
        def load(self):
            pass

        def predict(self, request):
            return [x * self.get_param("multiplier") for x in request["inputs"]]

    ns["EchoTimes"] = EchoTimes
    server = fn.to_mock_server(namespace=ns)
    first = _clean(server.test("/v2/models/infer", body={"inputs": [1, 2, 3]}))  # all ints -> classification
    second = _clean(server.test("/v2/models/infer", body={"inputs": [1.5, 2.5]}))  # stays classification (int cast)
    fn2 = api.new_function("tests", kind="serving")
    graph = fn2.set_topology("router", api.VotingEnsemble(executor_type="array"))
    for k, m in [("a", 1), ("b", 1), ("c", 2)]:
        graph.add_route(k, class_name="EchoTimes", model_path="", multiplier=m)
    server = fn2.to_mock_server(namespace=ns)
    third = _clean(server.test("/v2/models/infer", body={"inputs": [1.5, 2.0]}))  # floats -> regression
    return {"first": first, "second": second, "third": third}


def router_mock_direct(api):
    """tests/serving/test_serving.py:611-622 -- create_graph_server + add_route + test()"""
    ns = make_namespace(api)
    host = api.create_graph_server(graph=api.RouterStep())
    host.graph.add_route("my", class_name=ns["ModelTestingClass"], model_path="", multiplier=100)
    host.init_states(None, namespace=ns)
    host.init_object(ns)
    return _clean(host.test("/v2/models/my/infer", TESTDATA))


router_mock_direct.EXPECT = {("outputs",): 500}


def echo_plumbing(api):
    """BASELINE.json configs[0]: single V2ModelServer echo-model, MockEvent batch=1, CPU mock server"""
    ns = make_namespace(api)
    fn = api.new_function("tests", kind="serving")
    fn.set_topology("router")
    fn.add_model("m1", ".", class_name="EchoModel")
    server = fn.to_mock_server(namespace=ns)
    ev = api.MockEvent('{"inputs":[5]}', path="/v2/models/m1/infer")
    out = {"run": _clean(server.run(ev, get_body=True))}
    out["test"] = _clean(server.test("/v2/models/m1/infer", body={"inputs": [[1.5, 2.5], [3.0, 4.0]]}))
    return out


echo_plumbing.EXPECT = {("run", "outputs"): [5], ("run", "model_name"): "m1"}


def tracking(api):
    """tests/serving/test_tracking.py:41-95 + test_serving.py:625-637 -- records pushed to the stream"""
    ns = make_namespace(api)
    out = {}
    fn = api.new_function("tests", kind="serving")
    fn.set_topology("router")
    fn.add_model("my", ".", class_name=ns["ModelTestingClass"](multiplier=100))
    fn.set_tracking("dummy://")
    server = fn.to_mock_server(namespace=ns)
    out["resp"] = _clean(server.test("/v2/models/my/infer", TESTDATA))
    out["n_records"] = len(server.context.stream.output_stream.event_list)

    fn = api.new_function("tests", kind="serving")
    fn.set_topology("router", api.VotingEnsemble(vote_type="regression"))
    fn.add_model("1", ".", class_name=ns["TrackedModel"](multiplier=2))
    fn.add_model("2", ".", class_name=ns["TrackedModel"](multiplier=3))
    fn.set_tracking("dummy://")
    server = fn.to_mock_server(namespace=ns)
    resp = server.test("/v2/models/infer", '{"inputs": [[5, 6]]}')
    out["ens_outputs"] = _clean(resp)["outputs"]
    recs = {}
    for rec in server.context.stream.output_stream.event_list:
        recs[rec["model"]] = [rec["class"], _clean(rec["request"]["inputs"]), _clean(rec["resp"]["outputs"])]
    out["records"] = recs
    return out


tracking.EXPECT = {
    ("resp", "outputs"): 500,
    ("n_records",): 1,
    ("records",): {"1": ["TrackedModel", [[5, 6]], [10]], "2": ["TrackedModel", [[5, 6]], [15]],
                   "VotingEnsemble": ["VotingEnsemble", [[5, 6]], [12.5]]},
}


def parallel_run(api):
    """tests/serving/test_parallel.py:39-56"""
    ns = make_namespace(api)
    ns["Echo"] = ns["ParEcho"]
    out = {}
    for executor in ["array", "thread"]:
        fn = api.new_function("tests", kind="serving")
        graph = fn.set_topology("router", api.ParallelRun(extend_event=True, executor_type=executor))
        graph.add_route("c1", class_name="Echo", data={"a": 1, "b": 2})
        graph.add_route("c2", class_name="Echo", data={"c": 7})
        graph.add_route("c3", handler="my_hnd")
        server = fn.to_mock_server(namespace=ns)
        out[executor] = [_clean(server.test(body={"x": 8})), _clean(server.test("", {"x": 9}))]
    return out


parallel_run.EXPECT = {
    ("array", 0): {"x": 8, "a": 1, "b": 2, "c": 7, "mul": 16},
    ("thread", 1): {"x": 9, "a": 1, "b": 2, "c": 7, "mul": 18},
}


# =========================================================================== flow scenarios (sync engine)
def flow_basic_sync(api):
    """tests/serving/test_flow.py:63-94 -- ordering / before / after"""
    ns = make_namespace(api)
    out = []
    fn = api.new_function("tests", kind="serving", project="x")
    graph = fn.set_topology("flow", engine="sync")
    graph.add_step(name="s1", class_name="Chain")
    graph.add_step(name="s2", class_name="Chain", after="$prev")
    graph.add_step(name="s3", class_name="Chain", after="$prev")
    server = fn.to_mock_server(namespace=ns)
    out.append(server.test(body=[]))

    graph = fn.set_topology("flow", exist_ok=True, engine="sync")
    graph.add_step(name="s2", class_name="Chain")
    graph.add_step(name="s1", class_name="Chain", before="s2")
    graph.add_step(name="s3", class_name="Chain", after="s2")
    out.append(fn.to_mock_server(namespace=ns).test(body=[]))

    graph = fn.set_topology("flow", exist_ok=True, engine="sync")
    graph.add_step(name="s1", class_name="Chain")
    graph.add_step(name="s3", class_name="Chain", after="$prev")
    graph.add_step(name="s2", class_name="Chain", after="s1", before="s3")
    server = fn.to_mock_server(namespace=ns)
    out.append(server.test(body=[]))
    return {"flows": out, "project": server.context.project}


flow_basic_sync.EXPECT = {("flows",): [["s1", "s2", "s3"]] * 3, ("project",): "x"}


def flow_handlers_sync(api):
    """tests/serving/test_flow.py:97-132, 401-431 -- expression handlers, context injection, classes, set_flow"""
    ns = make_namespace(api)
    out = {}
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="s1", handler="(event + 1)").to(name="s2", handler="json.dumps")
    out["expr"] = fn.to_mock_server(namespace=ns).test(body=5)

    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="s1", handler=ns["myfunc1"]).to(name="s2", handler=ns["myfunc2"]).to(name="s3", handler=ns["myfunc1"])
    out["context"] = fn.to_mock_server(namespace=ns).test(body=5)

    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="s1", class_name="Echo").to(name="s2", class_name="RespName")
    out["classes"] = fn.to_mock_server(namespace=ns).test(body=5)

    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="s1", handler="(event + 1)").to(name="s2", handler="json.dumps")
    try:
        graph.set_flow(steps=[dict(name="r1", handler="(event + 10)")])
        out["set_flow_error"] = None
    except Exception as exc:  # noqa: BLE001
        out["set_flow_error"] = str(exc)
    graph.set_flow(steps=[dict(name="r1", handler="(event + 10)"), dict(name="r2", handler="json.dumps")], force=True)
    out["set_flow"] = fn.to_mock_server(namespace=ns).test(body=5)
    return out


flow_handlers_sync.EXPECT = {("expr",): "6", ("context",): 40, ("classes",): [5, "s2"], ("set_flow",): "15"}


def flow_on_error_sync(api):
    """tests/serving/test_flow.py:135-149"""
    ns = make_namespace(api)
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.add_step(name="s1", class_name="Chain")
    graph.add_step(name="raiser", class_name="Raiser", after="$prev").error_handler(
        name="catch", class_name="EchoError", full_event=True)
    graph.add_step(name="s3", class_name="Chain", after="$prev")
    resp = fn.to_mock_server(namespace=ns).test(body=[])
    # in sync mode the full-event error handler's return value (the event) becomes event.body, i.e. the
    # response is the (self-referencing) event object; the reference test accepts both forms
    if not isinstance(resp, dict):
        resp = {"origin_state": resp.origin_state, "error": resp.error}
    return _clean({"origin_state": resp["origin_state"], "error": resp["error"]})


flow_on_error_sync.EXPECT = {("origin_state",): "raiser"}


def flow_content_type(api):
    """tests/serving/test_flow.py:156-186"""
    ns = make_namespace(api)
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="totype", handler=ns["return_type"])
    server = fn.to_mock_server(namespace=ns)
    out = [
        server.test(body={"a": 1}),
        server.test(body="[1,2]"),
        server.test(body={"a": 1}, content_type="application/json"),
        server.test(body="[1,2]", content_type="application/json"),
        server.test(body="[1,2]", content_type="application/text"),
        server.test(body="xx [1,2]"),
        server.test(body="xx [1,2]", content_type="application/json", silent=True).status_code,
    ]
    fn = api.new_function("tests", kind="serving")
    fn.spec.default_content_type = "application/json"
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="totype", handler=ns["return_type"])
    out.append(fn.to_mock_server(namespace=ns).test(body="[1,2]"))
    return out


flow_content_type.EXPECT = {(): ["dict", "list", "dict", "list", "str", "str", 400, "list"]}


def flow_model_no_router(api):
    """tests/serving/test_serving.py:640-655 -- a model server as a plain flow step"""
    ns = make_namespace(api)
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to("ModelTestingClass", "my2", model_path=".", multiplier=100).respond()
    server = fn.to_mock_server(namespace=ns)
    return {
        "meta": _clean(server.test("/", method="GET")),
        "ready": server.test("/ready", method="GET").status_code,
        "infer": _clean(server.test("/", TESTDATA)),
    }


flow_model_no_router.EXPECT = {("meta", "name"): "my2", ("ready",): 200, ("infer", "outputs"): 500}


def flow_multi_function_sync(api):
    """tests/serving/test_flow.py:217-238 -- queue step + child function, sync engine"""
    ns = make_namespace(api)
    ns["ModelTestingClass"] = ns["EchoModel"]
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to("Echo", "e1").to("$queue", "q1", path="").to("*", "r1", function="f2").to("Echo", "e2", function="f2")
    fn.add_model("m1", class_name="ModelTestingClass", model_path=".")
    out = {}
    server = fn.to_mock_server(namespace=ns)
    out["root"] = _clean(server.test("/v2/models/m1/infer", body={"inputs": [5]}))
    server = fn.to_mock_server(namespace=ns, current_function="f2")
    out["f2"] = _clean(server.test(body={"inputs": [5]}))
    return out


flow_multi_function_sync.EXPECT = {("root", "outputs"): [5], ("f2", "outputs"): [5]}


def flow_path_control_sync(api):
    """tests/serving/test_flow.py:244-294 (sync halves) -- input_path/result_path; ensemble in a flow => [75]"""
    ns = make_namespace(api)
    out = {}
    for kind, (handler, cls) in {"handler": (ns["myfunc2"], None), "class": (None, "Mul")}.items():
        fn = api.new_function("test", kind="serving")
        flow = fn.set_topology("flow", engine="sync")
        flow.to(cls, handler=handler, name="x2", input_path="x", result_path="y.z").respond()
        out[kind] = fn.to_mock_server(namespace=ns).test(body={"x": 5})

    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="s1", class_name="Echo").to(
        "*mlrun.serving.routers.VotingEnsemble", name="r1", input_path="x", result_path="y", vote_type="regression",
    ).to(name="s3", class_name="Echo").respond()
    fn.add_model("m1", class_name="ModelClassList", model_path=".", multiplier=10)
    fn.add_model("m2", class_name="ModelClassList", model_path=".", multiplier=20)
    resp = fn.to_mock_server(namespace=ns).test("/v2/models/infer", body={"x": {"inputs": [[5]]}})
    out["ensemble"] = _clean(resp)
    return out


flow_path_control_sync.EXPECT = {
    ("handler",): {"x": 5, "y": {"z": 10}},
    ("class",): {"x": 5, "y": {"z": 10}},
    ("ensemble", "y", "outputs"): [75],
}


def step_to_dict(api):
    """tests/serving/test_flow.py:297-327 (V2ModelServer half) + router/flow to_dict round trip"""
    ms = api.V2ModelServer(name="ms", model_path="./xx", multiplier=7)
    d = ms.to_dict()
    d["class_name"] = d["class_name"].rsplit(".", 1)[-1]
    ns = make_namespace(api)
    router = api.RouterStep()
    router.routes = _routes(api, "ModelTestingClass", "multiplier", [100, 200, 300, 400])
    return {"model": d, "router": router.to_dict()}


step_to_dict.EXPECT = {
    ("model",): {"class_args": {"model_path": "./xx", "multiplier": 7, "protocol": "v2"},
                 "class_name": "V2ModelServer", "name": "ms"},
}


def route_cap(api):
    """tests/serving/test_serving.py:754-760 -- more than 4500 routes is rejected"""
    host = api.create_graph_server(graph=api.RouterStep())
    n = 0
    try:
        for key in range(4501):
            host.graph.add_route(f"my{key}", class_name="X", model_path="")
            n += 1
        return {"added": n, "raised": None}
    except Exception as exc:  # noqa: BLE001
        return {"added": n, "raised": exc.__class__.__name__}


route_cap.EXPECT = {("added",): 4500, ("raised",): "MLRunInvalidArgumentError"}


# =========================================================================== async-engine scenarios
# The reference's async engine is storey (not in /root/reference and not installed), so these cannot
# be generated from the real reference; they are pinned by the literal expectations of its tests.
def flow_async_basic(api):
    """tests/serving/test_async_flow.py:30-61 -- branches, responder, visit counts"""
    ns = make_namespace(api)
    fn = api.new_function("tests", kind="serving")
    flow = fn.set_topology("flow", engine="async")
    queue = flow.to(name="s1", class_name="ChainWithContext").to("$queue", "q1", path="")
    s2 = queue.to(name="s2", class_name="ChainWithContext", function="some_function")
    s2.to(name="s4", class_name="ChainWithContext")
    s2.to(name="s5", class_name="ChainWithContext").respond()
    queue.to(name="s3", class_name="ChainWithContext", function="some_other_function")
    server = fn.to_mock_server(namespace=ns)
    server.context.visits = {}
    resp = server.test(body=[])
    server.wait_for_completion()
    return {"resp": resp, "visits": dict(sorted(server.context.visits.items()))}


flow_async_basic.EXPECT = {
    ("resp",): ["s1", "s2", "s5"],
    ("visits",): {"s1": 1, "s2": 1, "s3": 1, "s4": 1, "s5": 1},
}
flow_async_basic.ASYNC = True


def flow_async_misc(api):
    """tests/serving/test_flow.py:97-110, 244-261 (async halves); test_async_flow.py:63-126;
    test_serving.py:658-677; tests/feature-store/test_steps.py:47-74"""
    ns = make_namespace(api)
    out = {}
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="async")
    graph.to(name="s1", handler="(event + 1)").to(name="s2", handler="json.dumps")
    graph["s2"].respond()
    server = fn.to_mock_server(namespace=ns)
    out["expr"] = server.test(body=5)
    server.wait_for_completion()

    for kind, (handler, cls) in {"handler": (ns["myfunc2"], None), "class": (None, "Mul")}.items():
        fn = api.new_function("test", kind="serving")
        flow = fn.set_topology("flow", engine="async")
        flow.to(cls, handler=handler, name="x2", input_path="x", result_path="y.z").respond()
        server = fn.to_mock_server(namespace=ns)
        out[f"path_{kind}"] = server.test(body={"x": 5})
        server.wait_for_completion()

    # queue must be followed by a step with function=
    fn = api.new_function("tests", kind="serving")
    flow = fn.set_topology("flow", engine="async")
    queue = flow.to(name="s1", class_name="ChainWithContext").to("$queue", "q1", path="")
    try:
        queue.to(name="s2", class_name="ChainWithContext")
        out["queue_needs_function"] = None
    except Exception as exc:  # noqa: BLE001
        out["queue_needs_function"] = str(exc)

    # nested router in an async flow: 5 * 2 * 200
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="async")
    graph.add_step(name="s1", class_name="Echo")
    graph.add_step(name="s2", handler="multiply_input", after="s1")
    graph.add_step(name="s3", class_name="Echo", after="s2")
    router = graph.add_step("*", name="ensemble", after="s2")
    router.add_route("m1", class_name="ModelClass", model_path=".", multiplier=100)
    router.add_route("m2", class_name="ModelClass", model_path=".", multiplier=200)
    router.add_route("m3:v1", class_name="ModelClass", model_path=".", multiplier=300)
    graph.add_step(name="final", class_name="Echo", after="ensemble").respond()
    server = fn.to_mock_server(namespace=ns)
    out["nested"] = _clean(server.test("/v2/models/m2/infer", body={"inputs": [5]}))
    server.wait_for_completion()

    # error handler in an async flow
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="async")
    chain = graph.to("Chain", name="s1")
    chain.to("Raiser").error_handler(name="catch", class_name="EchoError", full_event=True).to("Chain", name="s3")
    server = fn.to_mock_server(namespace=ns)
    resp = server.test(body=[])
    server.wait_for_completion()
    if not isinstance(resp, dict):
        resp = {"origin_state": resp.origin_state, "error": resp.error}
    out["on_error_origin"] = resp["origin_state"]
    out["on_error_has_error"] = bool(resp["error"])

    # chained models with input/result paths
    fn = api.new_function("demo", kind="serving")
    graph = fn.set_topology("flow", engine="async")
    graph.to(ns["ModelTestingClass"](name="m1", model_path=".", multiplier=2), result_path="m1", input_path="req").to(
        ns["ModelTestingClass"](name="m2", model_path=".", result_path="m2", multiplier=3, input_path="req")
    ).respond()
    server = fn.to_mock_server(namespace=ns)
    resp = server.test(body={"req": {"inputs": [5]}})
    server.wait_for_completion()
    out["chained_keys"] = list(resp.keys())
    out["chained"] = [resp["m1"]["outputs"], resp["m2"]["outputs"]]

    # SetEventMetadata (default engine = async)
    fn = api.new_function("test1", kind="serving")
    flow = fn.set_topology("flow")
    flow.to(api.SetEventMetadata(id_path="myid", key_path="mykey")).to(
        name="e", handler="extract_meta", full_event=True).respond()
    server = fn.to_mock_server(namespace=ns)
    out["event_meta"] = server.test(body={"myid": "34", "mykey": "123"})
    server.wait_for_completion()
    return out


flow_async_misc.EXPECT = {
    ("expr",): "6",
    ("path_handler",): {"x": 5, "y": {"z": 10}},
    ("path_class",): {"x": 5, "y": {"z": 10}},
    ("queue_needs_function",): "step 's2' must specify a function, because it follows a queue step",
    ("nested", "outputs"): 2000,
    ("on_error_origin",): "Raiser",
    ("on_error_has_error",): True,
    ("chained_keys",): ["req", "m1", "m2"],
    ("chained",): [10, 15],
    ("event_meta",): {"id": "34", "key": "123"},
}
flow_async_misc.ASYNC = True


# =========================================================================== feature steps (storey engine = dict events)
def _steps_frame():
    """tests/feature-store/test_steps.py:790-816 (get_data) with a fixed timestamp"""
    return {
        "name": ["A", "B", "C", "D", "E"],
        "age": [33, 4, 76, 90, 24],
        "department": ["IT", "RD", "RD", "Marketing", "IT"],
        "timestamp": ["2021-07-01 00:00:00"] * 5,  # a Thursday: day_of_week 3, hour 0 (test_steps.py:420-421)
        "id": ["a", "v", "h", "g", "j"],
    }


def _rows(frame):
    keys = list(frame.keys())
    return [{k: frame[k][i] for k in keys} for i in range(len(frame[keys[0]]))]


def _run_flow_rows(api, ns, steps, rows):
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    cur = graph
    for step in steps:
        cur = cur.to(step)
    server = fn.to_mock_server(namespace=ns)
    return [_clean(server.test(body=dict(r)), drop=()) for r in rows]


def steps_dict_events(api):
    """tests/feature-store/test_steps.py:113-153, 196-267, 311-329, 418-421, 660-662, 703-734 --
    every step run per event (storey engine semantics) through a sync flow"""
    ns = make_namespace(api)
    frame = _steps_frame()
    rows = _rows(frame)
    out = {}
    out["onehot"] = _run_flow_rows(api, ns, [api.OneHotEncoder(mapping={"department": list(frame["department"])})], rows)
    rows_none = [dict(r) for r in rows]
    rows_none[0]["department"] = None
    out["imputer"] = _run_flow_rows(api, ns, [api.Imputer(mapping={"department": "IT"})], rows_none)
    for with_original in (False, True):
        out[f"mapval_{with_original}"] = _run_flow_rows(
            api, ns,
            [api.MapValues(mapping={"age": {"ranges": {"child": [0, 30], "adult": [30, "inf"]}},
                                    "department": {"IT": 1, "Marketing": 2, "RD": 3}},
                           with_original_features=with_original)],
            rows)
    out["date"] = _run_flow_rows(api, ns, [api.DateExtractor(parts=["hour", "day_of_week"], timestamp_col="timestamp")], rows)
    out["date_default_col"] = _run_flow_rows(api, ns, [api.DateExtractor(parts=["hour", "day_of_week"])], rows[:1])
    out["drop"] = _run_flow_rows(api, ns, [api.DropFeatures(features=["age"])], rows)
    nones = [{"id": 1, "height": None, "age": 20}, {"id": 2, "height": 160, "age": None},
             {"id": 3, "height": float("nan"), "age": 19}]
    out["imputer_default"] = _run_flow_rows(api, ns, [api.Imputer(default_value=1)], nones)
    # unknown category -> all zeros; "-"/" " sanitised in key names (steps.py:465-468, 508-513)
    out["onehot_unknown"] = _run_flow_rows(
        api, ns, [api.OneHotEncoder(mapping={"c": ["a b", "c-d", 3]})], [{"c": "a b"}, {"c": "zz"}, {"c": 3}])
    # MapValues edge cases: first matching range wins; -inf; unmapped value passes through
    out["mapval_edges"] = _run_flow_rows(
        api, ns,
        [api.MapValues(mapping={"v": {"ranges": {"neg": ["-inf", 0], "lo": [0, 10], "lo2": [5, 20]}}, "k": {1: 10}})],
        [{"v": -3.5, "k": 1}, {"v": 0, "k": 2}, {"v": 7, "k": 1}, {"v": 10, "k": 1}, {"v": 25, "k": 3}])
    # chained: Imputer -> OneHot -> Drop
    out["chain"] = _run_flow_rows(
        api, ns,
        [api.Imputer(mapping={"department": "IT"}, default_value=0),
         api.OneHotEncoder(mapping={"department": ["IT", "RD", "Marketing"]}),
         api.DropFeatures(features=["name", "timestamp", "id"])],
        rows_none)
    try:
        _run_flow_rows(api, ns, [api.DropFeatures(features=["nope"])], rows[:1])
        out["drop_missing"] = None
    except Exception as exc:  # noqa: BLE001
        out["drop_missing"] = _first_line(str(exc).split("): ", 1)[-1])
    return out


steps_dict_events.EXPECT = {
    ("onehot", 0): {"name": "A", "age": 33, "department_IT": 1, "department_RD": 0, "department_Marketing": 0,
                    "timestamp": "2021-07-01 00:00:00", "id": "a"},
    ("imputer", 0, "department"): "IT",
    ("mapval_False", 1): {"age": "child", "department": 3},
    ("date", 0, "timestamp_day_of_week"): 3,
    ("date", 0, "timestamp_hour"): 0,
}


def steps_pandas_engine(api):
    """tests/feature-store/test_steps.py (pandas halves) -- `_do_pandas` on the 5-row frame"""
    import pandas as pd

    frame = _steps_frame()

    def df():
        return pd.DataFrame(frame).set_index("id")

    def dump(d):
        return {"columns": [str(c) for c in d.columns], "dtypes": [str(t) for t in d.dtypes],
                "index": [str(i) for i in d.index], "values": _clean(d.astype(object).where(d.notna(), None).values.tolist())}

    out = {}
    out["onehot"] = dump(api.OneHotEncoder(mapping={"department": list(frame["department"])}).do(df()))
    d = df()
    d.loc["a", "department"] = None
    out["imputer"] = dump(api.Imputer(mapping={"department": "IT"}).do(d))
    for with_original in (False, True):
        out[f"mapval_{with_original}"] = dump(
            api.MapValues(mapping={"age": {"ranges": {"child": [0, 30], "adult": [30, "inf"]}},
                                   "department": {"IT": 1, "Marketing": 2, "RD": 3}},
                          with_original_features=with_original).do(df()))
    out["date"] = dump(api.DateExtractor(parts=["hour", "day_of_week"], timestamp_col="timestamp").do(df()))
    out["drop"] = dump(api.DropFeatures(features=["age"]).do(df()))
    return out


# the reference's pandas Imputer does `event[col].fillna(val, inplace=True)` (steps.py:412), which is a
# silent no-op under the pandas 3.0 copy-on-write installed here (the reference pins pandas<2.2, where it
# works); the golden run therefore shows the un-imputed frame.  That key is pinned by the reference
# test's literal instead (tests/feature-store/test_steps.py:196-238: department[0] == "IT").
steps_pandas_engine.GOLDEN_SKIP = [("imputer",)]
steps_pandas_engine.EXPECT = {
    ("imputer", "values", 0): ["A", 33, "IT", "2021-07-01 00:00:00"],
    ("onehot", "columns"): ["name", "age", "department_IT", "department_RD", "department_Marketing", "timestamp"],
    ("mapval_False", "values"): [["adult", 1], ["child", 3], ["adult", 3], ["adult", 2], ["child", 1]],
}


def steps_validate_args(api):
    """feature_store/steps.py:331-370, 737-753 + tests/feature-store/test_steps.py:466-547 -- the ingest-time argument checks
    of MapValues and DropFeatures (`validate_args`, called by FeatureSet.validate_steps), driven directly"""
    import types

    fset = types.SimpleNamespace(spec=types.SimpleNamespace(entities={"id": None, "name": None}, label_column="y",
                                                            timestamp_key="when"))

    def attempt(cls, **kw):
        try:
            cls.validate_args(fset, **kw)
            return None
        except Exception as exc:  # noqa: BLE001
            return f"{type(exc).__name__}: {_first_line(exc)}"

    nan = float("nan")
    return {
        "map_ranges_mixed": attempt(api.MapValues, mapping={"age": {"ranges": {"one": [0, 30], "two": ["a", "inf"]}}}),
        "map_values_mixed": attempt(api.MapValues, mapping={"names": {"A": 1, "B": False}}),
        "map_combined": attempt(api.MapValues, mapping={"age": {"ranges": {"one": [0, 30], "two": ["a", "inf"]}, 4: "kid"}}),
        "map_ok_ranges": attempt(api.MapValues, mapping={"age": {"ranges": {"child": [0, 18], "adult": [18, "inf"]}}}),
        "map_ok_inf_spellings": attempt(api.MapValues, mapping={"x": {"ranges": {0: ["-inf", 0], 1: [0, "inf"]}}}),
        "map_int_and_float_bounds": attempt(api.MapValues, mapping={"x": {"ranges": {0: [0, 0.5], 1: [0.5, 1]}}}),
        "map_nan_value_is_ignored": attempt(api.MapValues, mapping={"x": {1: nan, 2: 5}}),
        "map_ok_strings": attempt(api.MapValues, mapping={"dep": {"IT": "tech", "RD": "tech"}}),
        "map_second_column_bad": attempt(api.MapValues, mapping={"a": {1: 2}, "b": {1: "x", 2: 3}}),
        "map_no_mapping": attempt(api.MapValues),
        "drop_entity": attempt(api.DropFeatures, features=["age", "id"]),
        "drop_label": attempt(api.DropFeatures, features=["y"]),
        "drop_timestamp": attempt(api.DropFeatures, features=["age", "when"]),
        "drop_ok": attempt(api.DropFeatures, features=["age", "dep"]),
        "drop_nothing": attempt(api.DropFeatures),
    }


steps_validate_args.EXPECT = {
    ("map_ranges_mixed",): "MLRunInvalidArgumentError: MapValues - mapping values of the same column must be in the same type, "
                           "which was not the case for Column 'age'",
    ("map_combined",): "MLRunInvalidArgumentError: MapValues - mapping values of the same column can not combine ranges and single "
                       "replacement, which is the case for column 'age'",
    ("map_ok_ranges",): None,
}


def _captured(fn):
    import contextlib
    import io

    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = fn()
    return res, buf.getvalue().splitlines()


_VALIDATOR_RULES = {"age": dict(min=30, severity="info"), "bid": dict(min=1.5, max=9.0, severity="warn"),
                    "qty": dict(max=100), "name": dict(min=3, severity="info")}


def validator_events(api):
    """feature_store/steps.py:94-128 + mlrun/features.py:265-321 -- FeaturesetValidator on storey events: events pass through
    unchanged, violations are printed (message, key, arguments); a comparison that raises is a violation too"""
    import types

    step = api.validator_step(_VALIDATOR_RULES, None)
    only_age = api.validator_step(_VALIDATOR_RULES, ["age"])
    out = {}
    events = [("k1", {"age": 25, "bid": 2.0, "qty": 5, "name": 7}), (None, {"age": 31, "bid": 0.5, "qty": 500, "other": 1}),
              ("k3", {"age": 30, "bid": 9.0}), ("k4", {"age": float("nan"), "bid": 9.5, "qty": None, "name": "bob"}),
              ("k5", {"bid": "x" * 300}), ("k6", {"qty": 10**50})]  # the reported value is cut at 40 characters
    for i, (key, body) in enumerate(events):
        ev = types.SimpleNamespace(body=dict(body), key=key)
        res, lines = _captured(lambda: step.do(ev))
        out[f"e{i}"] = {"same_event": res is ev, "body_unchanged": _clean(ev.body) == _clean(body), "printed": lines}
    ev = types.SimpleNamespace(body={"age": 1, "bid": 0.1}, key="k")
    _res, lines = _captured(lambda: only_age.do(ev))
    out["columns_filter"] = lines
    return out


validator_events.EXPECT = {
    ("e0", "printed"): ["info! age value is smaller than min, key=k1 args={'min': 30, 'value': '25'}"],
    ("e2", "printed"): [],
}


def validator_pandas(api):
    """feature_store/steps.py:130-149 + tests/feature-store/test_steps.py:550-627 -- the pandas engine of FeaturesetValidator:
    one report per violating column, the frame passes through unchanged"""
    import types

    import pandas as pd

    step = api.validator_step(_VALIDATOR_RULES, None)
    frame = pd.DataFrame({"age": [25, 31, 29, 40], "bid": [2.0, 0.5, 9.5, 3.0], "qty": [5, 500, 7, 101], "other": [1, 2, 3, 4]},
                         index=pd.Index([10, 11, 12, 13], name="id"))
    ev = types.SimpleNamespace(body=frame.copy(), key=None)
    res, lines = _captured(lambda: step.do(ev))
    return {"same_event": res is ev, "frame_unchanged": bool(ev.body.equals(frame)), "printed": lines}


def set_event_metadata_logic(api):
    """feature_store/steps.py:635-696 + tests/feature-store/test_steps.py:47-76 -- SetEventMetadata driven directly:
    post_init builds the taggers, do() copies id / key from (nested) body paths as strings, random_id draws a hex id"""
    import types

    def run(step, body):
        step.post_init()
        ev = types.SimpleNamespace(body=body, id="orig-id", key="orig-key")
        res = step.do(ev)
        return {"same_event": res is ev, "id": ev.id, "key": ev.key}

    out = {
        "both": run(api.SetEventMetadata(id_path="myid", key_path="mykey"), {"myid": "34", "mykey": "123"}),
        "nested_and_numeric": run(api.SetEventMetadata(id_path="meta.id", key_path="k"), {"meta": {"id": 7}, "k": 2.5}),
        "key_only": run(api.SetEventMetadata(key_path="mykey"), {"mykey": ["a", 1]}),
        "missing_path": run(api.SetEventMetadata(id_path="nope.deeper"), {"x": 1}),
        "nothing": run(api.SetEventMetadata(), {"x": 1}),
    }
    rnd = run(api.SetEventMetadata(random_id=True, key_path="k"), {"k": "kk"})
    out["random"] = {"key": rnd["key"], "id_is_hex32": len(rnd["id"]) == 32 and all(c in "0123456789abcdef" for c in rnd["id"])}
    rnd_wins = run(api.SetEventMetadata(id_path="myid", random_id=True), {"myid": "34"})
    out["random_after_path"] = rnd_wins["id"] != "34" and len(rnd_wins["id"]) == 32
    step = api.SetEventMetadata(id_path="a", key_path="b")
    out["full_event"] = bool(getattr(step, "_full_event", None))
    return out


set_event_metadata_logic.EXPECT = {("both",): {"same_event": True, "id": "34", "key": "123"},
                                   ("random", "id_is_hex32"): True}


def pickle_model_from_path(api):
    """frameworks/_ml_common/pkl_model_server.py:40-50 + serving/v2_serving.py:166-202 + artifacts/model.py:412-483 (local
    slice) -- a pickled scikit-learn model served from `model_path`: the file itself, or a directory searched for the suffix"""
    import os
    import tempfile

    import cloudpickle
    from sklearn.linear_model import LinearRegression

    model = LinearRegression()
    model.coef_, model.intercept_, model.n_features_in_ = np.array([0.5, -2.0, 4.0]), 1.25, 3
    ns = {"SKLearnModelServer": api.SKLearnModelServer}
    out = {}

    def serve(path):
        fn = api.new_function("t", kind="serving")
        fn.set_topology("router")
        fn.add_model("m1", model_path=path, class_name="SKLearnModelServer")
        try:
            server = fn.to_mock_server(namespace=ns)
        except Exception as exc:  # noqa: BLE001
            return {"load_failed": type(exc).__name__}
        return _clean(server.test("/v2/models/m1/infer", {"inputs": [[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]]}))

    with tempfile.TemporaryDirectory() as tmp:
        full = os.path.join(tmp, "with_model")
        os.mkdir(full)
        with open(os.path.join(full, "model.pkl"), "wb") as fp:
            cloudpickle.dump(model, fp)
        with open(os.path.join(full, "notes.txt"), "w") as fp:
            fp.write("not a model")
        empty = os.path.join(tmp, "without_model")
        os.mkdir(empty)
        out["file"] = serve(os.path.join(full, "model.pkl"))
        out["directory"] = serve(full)
        out["directory_without_model"] = serve(empty)
        out["missing_file"] = serve(os.path.join(tmp, "nope.pkl"))
    return out


pickle_model_from_path.EXPECT = {("file", "outputs"): [0.5 - 4.0 + 12.0 + 1.25, 1.25], ("directory", "outputs"): [9.75, 1.25]}


def model_async_load(api):
    """tests/serving/test_serving.py:471-501 -- load_mode="async": the model loads in a thread; until it is done `ready` is 408
    and an HTTP infer fails, afterwards requests (here a stream-triggered one) are served"""
    import threading
    import time

    gate = threading.Event()

    class SlowModel(api.V2ModelServer):
        def load(self):
            gate.wait(20)

        def predict(self, request):
            return request["inputs"][0]

    router = api.RouterStep()
    router.routes = {"m5": api.TaskStep("SlowModel", class_args={"model_path": ""})}
    ctx = api.init_from_spec(_spec(router.to_dict(), mode="async"), {"SlowModel": SlowModel})
    out = {}

    def text(resp):
        return _first_line(resp.body if isinstance(resp.body, str) else resp.body.decode())

    resp = ctx.mlrun_handler(ctx, api.MockEvent("", path="/v2/models/m5/ready", method="GET"))
    out["ready_before"] = resp.status_code
    resp = ctx.mlrun_handler(ctx, api.MockEvent(TESTDATA, path="/v2/models/m5/infer"))
    out["infer_before"] = [resp.status_code, text(resp)]
    gate.set()
    for _ in range(500):
        resp = ctx.mlrun_handler(ctx, api.MockEvent("", path="/v2/models/m5/ready", method="GET"))
        if resp.status_code == 200:
            break
        time.sleep(0.02)
    out["ready_after"] = resp.status_code
    resp = ctx.mlrun_handler(ctx, api.MockEvent('{"model": "m5", "inputs": [5]}', trigger=api.MockTrigger(kind="stream")))
    out["stream_after"] = _resp(resp)
    return out


model_async_load.EXPECT = {("ready_before",): 408, ("ready_after",): 200, ("stream_after", "body", "outputs"): 5}


def class_args_protocol(api):
    """serving/states.py:437-512 (init_object / get_full_class_args) -- what a step class's constructor receives: `_x` class
    args arrive as callables under `x`; name / context / input_path / result_path / full_event / graph_step only when the
    signature names them or takes **kwargs"""

    def triple(x):
        return x * 3

    class TakesNothing:
        def __init__(self):
            self.got = sorted(vars(self))

        def do(self, x):
            return [x, "nothing"]

    class TakesSome:
        def __init__(self, name=None, fn=None, factor=1):
            self.name, self.fn, self.factor = name, fn, factor

        def do(self, x):
            return {"name": self.name, "value": self.fn(x) * self.factor}

    class TakesKwargs:
        def __init__(self, **kwargs):
            self.kw = kwargs

        def do(self, x):
            step = self.kw.get("graph_step")
            return {"keys": sorted(self.kw), "step_name": getattr(step, "name", None), "ctx": self.kw.get("context") is not None,
                    "fn": self.kw["fn"](x), "full_event": self.kw.get("full_event"), "input_path": self.kw.get("input_path")}

    class PrefersDoEvent:
        def do(self, x):
            return "do"

        def do_event(self, event):
            event.body = ["do_event", event.body]
            return event

    ns = {"TakesNothing": TakesNothing, "TakesSome": TakesSome, "TakesKwargs": TakesKwargs, "PrefersDoEvent": PrefersDoEvent,
          "triple": triple}
    out = {}

    def serve(*args, **kw):
        fn = api.new_function("t", kind="serving")
        graph = fn.set_topology("flow", engine="sync")
        graph.to(*args, **kw).respond()
        return fn.to_mock_server(namespace=ns)

    out["nothing"] = serve("TakesNothing", "s").test(body=5)
    out["some_named_fn"] = serve("TakesSome", "s", _fn="triple", factor=2).test(body=5)
    out["some_lambda_fn"] = serve("TakesSome", "lam", _fn="(event + 1)").test(body=5)
    out["kwargs"] = serve("TakesKwargs", "kw", _fn="triple", extra=1).test(body=5)
    out["kwargs_paths"] = serve("TakesKwargs", "kw2", _fn="triple", input_path="a").test(body={"a": 2})
    out["do_event_wins"] = serve("PrefersDoEvent", "p").test(body=1)
    try:
        serve("TakesNothing", "s", unexpected=1)
        out["unexpected_arg"] = None
    except Exception as exc:  # noqa: BLE001
        out["unexpected_arg"] = type(exc).__name__
    try:
        serve("TakesSome", "s", _fn="no_such_function")
        out["unknown_callable"] = None
    except Exception as exc:  # noqa: BLE001
        out["unknown_callable"] = type(exc).__name__
    return out


class_args_protocol.EXPECT = {("some_named_fn",): {"name": "s", "value": 30}, ("some_lambda_fn",): {"name": "lam", "value": 6},
                              ("do_event_wins",): ["do_event", 1]}


def model_hooks(api):
    """serving/v2_serving.py:228-342, 344-383 -- the model-class hooks around predict: preprocess / validate / postprocess,
    `logged_results` deciding what the tracking stream records, get_param falling back to the server's parameters,
    set_metric"""

    class Hooked(api.V2ModelServer):
        def load(self):
            self.set_metric("loaded", 1)

        def preprocess(self, request, operation):
            request["inputs"] = [v + self.get_param("shift", 0) for v in request["inputs"]]
            request["seen_op"] = operation
            return request

        def validate(self, request, operation):
            request = super().validate(request, operation)
            if any(v < 0 for v in request["inputs"]):
                raise ValueError("negative input")
            return request

        def predict(self, request):
            return [v * self.get_param("multiplier") for v in request["inputs"]]

        def postprocess(self, response):
            response["rounded"] = [round(v) for v in response["outputs"]]
            return response

        def logged_results(self, request, response, op):
            if self.get_param("filter_logs", False):
                return [request["inputs"][0]], [response["outputs"][0]]
            return None, None

    out = {}
    for tag, filter_logs in (("full_records", False), ("filtered_records", True)):
        fn = api.new_function("hooks", kind="serving")
        fn.set_topology("router")
        fn.add_model("m", ".", class_name="Hooked", multiplier=2.5, filter_logs=filter_logs)
        fn.set_tracking("dummy://")
        fn.spec.parameters["shift"] = 1  # a server parameter: reached through get_param's fallback
        server = fn.to_mock_server(namespace={"Hooked": Hooked})
        resp = server.test("/v2/models/m/infer", {"inputs": [1, 2, 3]}, event_id="e1")
        bad = server.test("/v2/models/m/infer", {"inputs": [-5]}, silent=True)
        records = [{k: r[k] for k in ("model", "op", "request", "resp")} for r in server.context.stream.output_stream.event_list]
        model = server.graph.routes["m"]._object
        out[tag] = {"response": _clean(resp), "bad_status": bad.status_code,
                    "bad_text": _first_line(bad.body if isinstance(bad.body, str) else bad.body.decode()),
                    "records": _clean(records), "metrics": dict(model.metrics)}
    return out


model_hooks.EXPECT = {("full_records", "response", "outputs"): [5.0, 7.5, 10.0], ("full_records", "response", "rounded"): [5, 8, 10],
                      ("filtered_records", "bad_status"): 400}


def custom_router(api):
    """serving/routers.py:46-211 -- the router class protocol: a ModelRouter subclass with its own url / health prefixes and
    pre / post hooks, served from a router topology built with the class object"""
    ns = make_namespace(api)

    class TaggingRouter(api.ModelRouter):
        def preprocess(self, event):
            if isinstance(event.body, dict) and "inputs" in event.body:
                event.body["inputs"] = [v + 1 for v in event.body["inputs"]]
            return event

        def postprocess(self, event):
            if isinstance(event.body, dict):
                event.body["via"] = self.name
            return event

    ns["TaggingRouter"] = TaggingRouter
    fn = api.new_function("r", kind="serving")
    graph = fn.set_topology("router", "TaggingRouter", url_prefix="/api/models", health_prefix="/api/health", name="tagger")
    graph.add_route("a", class_name="ModelTestingClass", model_path=".", multiplier=10)
    graph.add_route("b:v2", class_name="ModelTestingClass", model_path=".", multiplier=100)
    server = fn.to_mock_server(namespace=ns)

    def call(path, body=None, method=None):
        resp = server.test(path, body, method=method or ("POST" if body is not None else "GET"), silent=True)
        if hasattr(resp, "status_code"):
            text = resp.body if isinstance(resp.body, str) else (resp.body or b"").decode()
            return {"status": resp.status_code, "text": _first_line(text)}
        return _clean(resp)

    return {
        "infer_a": call("/api/models/a/infer", {"inputs": [5]}),
        "infer_b_version": call("/api/models/b/versions/v2/infer", {"inputs": [5]}),
        "default_prefix_is_gone": call("/v2/models/a/infer", {"inputs": [5]}),
        "list": call("/api/models/"),
        "health": call("/api/health"),
        "old_health": call("/v2/health"),
        "unknown_model": call("/api/models/zz/infer", {"inputs": [5]}),
        "body_model": call("/api/models", {"model": "b:v2", "inputs": [1]}),
        "to_dict": {k: v for k, v in graph.to_dict().items() if k in ("kind", "class_name", "class_args", "name")},
    }


custom_router.EXPECT = {("infer_a", "outputs"): 60}


def server_run_details(api):
    """serving/server.py:196-308 + serving/utils.py:22-23 -- GraphServer.run / test around the graph: the MLRUN-EVENT-ID and
    MLRUN-EVENT-PATH header overrides, bytes bodies, get_body, non-JSON content types, the JSON encoding of the response"""
    ns = make_namespace(api)
    fn = api.new_function("t", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="meta", handler="extract_meta", full_event=True).respond()
    server = fn.to_mock_server(namespace=ns)
    out = {
        "event_id_arg": server.test(body={"x": 1}, event_id="abc"),
        "id_header_wins": server.test(body={"x": 1}, event_id="abc", headers={"MLRUN-EVENT-ID": "from-header"}),
        "path_header": _resp(server.run(api.MockEvent(body={"x": 1}, path="/a", headers={"MLRUN-EVENT-PATH": "/b"}))),
    }
    fn = api.new_function("t2", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="ret", handler="return_type").respond()
    server = fn.to_mock_server(namespace=ns)
    out["bytes_json"] = server.test(body=b'{"a": [1, 2]}')
    out["bytes_text"] = server.test(body=b"plain", content_type="text/plain")
    out["str_not_json"] = server.test(body="{not json", content_type="text/plain")
    out["bad_json_default_type"] = _resp(server.test(body="{not json", silent=True))
    out["number_body"] = server.test(body=5)
    fn = api.new_function("t3", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to("Echo", "e").respond()
    server = fn.to_mock_server(namespace=ns)
    resp = server.run(api.MockEvent(body={"k": [1, 2.5, None, "s"]}))
    out["response_is_json_text"] = [type(resp).__name__, resp if isinstance(resp, str) else None]
    out["get_body_keeps_object"] = server.run(api.MockEvent(body={"k": 1}), get_body=True)
    out["test_returns_object"] = server.test(body={"k": [1, 2]})
    out["str_body_passthrough"] = server.test(body="hello", content_type="text/plain")
    return out


def tracking_sampling_batching(api):
    """serving/v2_serving.py:429-504 -- _ModelLogPusher: `log_stream_sample` keeps every n-th request, `log_stream_batch`
    groups the kept ones into one stream record (lists of requests / responses / timings); errors are pushed at once"""
    ns = make_namespace(api)

    def serve(**params):
        fn = api.new_function("trk", kind="serving")
        fn.set_topology("router")
        fn.add_model("my", ".", class_name=ns["ModelTestingClass"](multiplier=10))
        fn.set_tracking("dummy://")
        fn.spec.parameters.update(params)
        return fn.to_mock_server(namespace=ns)

    def shape(rec):
        keep = {k: rec[k] for k in ("model", "op", "class", "function_uri") if k in rec}
        for key in ("request", "resp", "requests", "error"):
            if key in rec:
                keep[key] = _clean(rec[key])
        if "values" in rec:  # one entry per batched request: [request, op, resp, when, microsec, metrics]
            keep["values"] = [[_clean(v[0]), v[1], _clean(v[2]), type(v[3]).__name__, type(v[4]).__name__, v[5]] for v in rec["values"]]
        for key in ("when", "microsec", "metrics"):
            if key in rec:
                keep[f"{key}_kind"] = type(rec[key]).__name__
        if isinstance(rec.get("microsec"), list):
            keep["n_timings"] = len(rec["microsec"])
        return keep

    out = {}
    for tag, params in (("sample3", {"log_stream_sample": 3}), ("batch2", {"log_stream_batch": 2}),
                        ("sample2_batch2", {"log_stream_sample": 2, "log_stream_batch": 2})):
        server = serve(**params)
        for i in range(7):
            server.test("/v2/models/my/infer", {"inputs": [i]}, event_id=f"e{i}")
        out[tag] = [shape(r) for r in server.context.stream.output_stream.event_list]
    server = serve(log_stream_batch=3)
    server.test("/v2/models/my/infer", {"inputs": [1]}, event_id="ok")
    server.test("/v2/models/my/infer", {"inputs": "not-a-list"}, event_id="bad", silent=True)
    out["error_in_a_batch"] = [shape(r) for r in server.context.stream.output_stream.event_list]
    return out


def graph_validation_errors(api):
    """serving/states.py:1073-1184 (check_and_process_graph) -- what a flow refuses at server start: loops, several
    responders, a sync flow with two entry points, unknown from_step / final_step, a child function nobody points at"""
    ns = make_namespace(api)

    def attempt(build, **server_kw):
        fn = api.new_function("g", kind="serving")
        try:
            build(fn)
            fn.to_mock_server(namespace=ns, **server_kw)
            return None
        except Exception as exc:  # noqa: BLE001
            return f"{type(exc).__name__}: {_first_line(exc)}"

    def loop(fn):
        g = fn.set_topology("flow", engine="sync")
        g.add_step("Echo", "a", after="c")
        g.add_step("Echo", "b", after="a")
        g.add_step("Echo", "c", after="b")

    def real_loop(fn):
        g = fn.set_topology("flow", engine="sync")
        g.add_step("Echo", "start")
        g.add_step("Echo", "a", after="start")
        g.add_step("Echo", "b", after="a")
        g["a"].after_step("b")

    def two_responders(fn):
        g = fn.set_topology("flow", engine="async")
        s = g.to("Echo", "a")
        s.to("Echo", "b").respond()
        s.to("Echo", "c").respond()

    def two_starts_sync(fn):
        g = fn.set_topology("flow", engine="sync")
        g.add_step("Echo", "a")
        g.add_step("Echo", "b")

    def unknown_after(fn):
        g = fn.set_topology("flow", engine="sync")
        g.add_step("Echo", "a")
        g.add_step("Echo", "b", after="nope")

    def bad_from_step(fn):
        g = fn.set_topology("flow", engine="sync")
        g.to("Echo", "a").to("Echo", "b")
        g.from_step = "zz"

    def bad_final_step(fn):
        g = fn.set_topology("flow", engine="sync")
        g.to("Echo", "a").to("Echo", "b")
        g.final_step = "zz"

    def ok_chain(fn):
        fn.set_topology("flow", engine="sync").to("Echo", "a").to("Echo", "b")

    def child_function_steps(fn):
        g = fn.set_topology("flow", engine="async")
        g.to("Echo", "a").to("$queue", "q", path="").to("Echo", "b", function="child")

    return {"loop": attempt(loop), "real_loop": attempt(real_loop), "two_responders": attempt(two_responders), "two_starts_sync": attempt(two_starts_sync),
            "unknown_after": attempt(unknown_after), "bad_from_step": attempt(bad_from_step),
            "bad_final_step": attempt(bad_final_step), "ok_chain": attempt(ok_chain),
            "no_step_for_function": attempt(ok_chain, current_function="other"),
            "child_function_found": attempt(child_function_steps, current_function="child")}


def event_envelope(api):
    """serving/server.py:437-490 -- MockEvent / MockTrigger / Response: defaults and what the constructor keeps"""
    def fields(ev):
        d = {k: v for k, v in vars(ev).items() if not k.startswith("_")}
        if "id" in d:
            d["id"] = "uuid32" if isinstance(d["id"], str) and len(d["id"]) == 32 else d["id"]
        if d.get("trigger") is not None:
            d["trigger"] = {k: v for k, v in vars(d["trigger"]).items()}
        for k, v in list(d.items()):
            if not isinstance(v, (str, int, float, bool, type(None), dict, list)):
                d[k] = type(v).__name__
        return d

    ctx = api.GraphContext()
    resp = ctx.Response(body={"a": 1})
    return {
        "bare": fields(api.MockEvent()),
        "full": fields(api.MockEvent(body={"x": 1}, content_type="application/json", headers={"H": "1"}, method="PUT", path="/p",
                                     event_id="my-id", trigger=api.MockTrigger(kind="stream", name="t1"), offset=7, time="now")),
        "path_default_method": fields(api.MockEvent(body="b", path="/x"))["method"],
        "str": str(api.MockEvent(body=[1], event_id="e")),
        "trigger_default": {k: v for k, v in vars(api.MockTrigger()).items()},
        "response_defaults": {k: v for k, v in vars(resp).items()},
        "response_full": {k: v for k, v in vars(ctx.Response(headers={"h": 1}, body="x", content_type="text/plain", status_code=404)).items()},
        "response_repr": repr(resp),
    }


def set_tracking_params(api):
    """runtimes/nuclio/serving.py:308-354 -- what set_tracking leaves in the function spec, and that a model server only
    pushes records while tracking is enabled"""
    ns = make_namespace(api)
    out = {}

    def spec_of(**kw):
        fn = api.new_function("t", kind="serving")
        fn.set_topology("router")
        fn.set_tracking(**kw)
        return {"track_models": fn.spec.track_models, "parameters": dict(fn.spec.parameters)}

    out["defaults"] = spec_of()
    out["all"] = spec_of(stream_path="dummy://x", batch=4, sample=2, stream_args={"mock": True})
    out["disabled"] = spec_of(stream_path="dummy://x", enable_tracking=False)
    out["zero_batch_is_not_set"] = spec_of(batch=0, sample=0)
    for tag, enable in (("records_when_enabled", True), ("records_when_disabled", False)):
        fn = api.new_function("t", kind="serving")
        fn.set_topology("router")
        fn.add_model("my", ".", class_name=ns["ModelTestingClass"](multiplier=100))
        fn.set_tracking("dummy://", enable_tracking=enable)
        server = fn.to_mock_server(namespace=ns)
        server.test("/v2/models/my/infer", TESTDATA)
        stream = getattr(getattr(server.context, "stream", None), "output_stream", None)
        out[tag] = len(stream.event_list) if stream is not None else None
    return out


def add_model_args(api):
    """runtimes/nuclio/serving.py:356-445 -- add_model's argument rules in a router topology: what is required, the default
    class of the function spec, a model object carrying its own arguments, the route's serialised form"""
    ns = make_namespace(api)

    def attempt(**kw):
        fn = api.new_function("t", kind="serving")
        fn.set_topology("router")
        default = kw.pop("default_class", None)
        if default:
            fn.spec.default_class = default
        try:
            route = fn.add_model("m1", **kw)
            d = route.to_dict()
            return {k: d.get(k) for k in ("class_name", "class_args", "handler", "function") if d.get(k) is not None}
        except Exception as exc:  # noqa: BLE001
            return f"{type(exc).__name__}: {_first_line(exc)}"

    return {
        "no_path": attempt(class_name="ModelTestingClass"),
        "path_without_class": attempt(model_path="."),
        "default_class": attempt(model_path=".", default_class="ModelTestingClass", multiplier=3),
        "class_not_a_string": attempt(model_path=".", class_name=5),
        "handler_and_child": attempt(model_path=".", class_name="ModelTestingClass", handler="explain", child_function="child"),
        "pathlike": attempt(model_path=__import__("pathlib").PurePosixPath("/models/m"), class_name="ModelTestingClass"),
        "object_with_path": attempt(model_path="/x/y", class_name=ns["ModelTestingClass"](multiplier=7)),
        "object_without_path": attempt(class_name=ns["ModelTestingClass"](multiplier=7, model_path="own")),
    }


def parallel_run_details(api):
    """serving/routers.py:214-455 -- ParallelRun beyond the reference's own test: extend_event off (only the routes' results),
    later routes overwriting earlier keys, a failing route under the array and the thread executors, input / result paths"""
    ns = make_namespace(api)
    ns["Echo"] = ns["ParEcho"]

    def boom(event):
        raise ValueError("route failed")

    ns["boom"] = boom
    out = {}

    def call(server, body):
        resp = server.test(body=body, silent=True)
        if hasattr(resp, "status_code"):
            text = resp.body if isinstance(resp.body, str) else (resp.body or b"").decode()
            return {"status": resp.status_code, "text": _first_line(text)}
        return _clean(resp)

    for executor in ("array", "thread"):
        fn = api.new_function("t", kind="serving")
        graph = fn.set_topology("router", api.ParallelRun(executor_type=executor))
        graph.add_route("c1", class_name="Echo", data={"a": 1, "k": "first"})
        graph.add_route("c2", class_name="Echo", data={"c": 7, "k": "second"})
        out[f"{executor}_plain"] = call(fn.to_mock_server(namespace=ns), {"x": 8})
        fn = api.new_function("t", kind="serving")
        graph = fn.set_topology("router", api.ParallelRun(executor_type=executor, extend_event=True))
        graph.add_route("ok", class_name="Echo", data={"a": 1})
        graph.add_route("bad", handler="boom")
        out[f"{executor}_failing_route"] = call(fn.to_mock_server(namespace=ns), {"x": 1})
    fn = api.new_function("t", kind="serving")
    graph = fn.set_topology("router", api.ParallelRun(executor_type="array", extend_event=True, input_path="req", result_path="res"))
    graph.add_route("c1", class_name="Echo", data={"a": 1})
    out["paths"] = call(fn.to_mock_server(namespace=ns), {"req": {"x": 2}, "other": 3})
    return out


def ensemble_odd_requests(api):
    """serving/routers.py:457-991 + serving/v2_serving.py:228-371 -- a voting ensemble asked odd things: weights naming an
    unknown model, an invalid vote type, bodies without a usable `inputs`, operations that do not exist, metadata / ready /
    metrics paths of a child model"""
    ns = make_namespace(api)
    out = {}

    def call(server, path, body, **kw):
        resp = server.test(path, body, silent=True, **kw)
        if hasattr(resp, "status_code"):
            text = resp.body if isinstance(resp.body, str) else (resp.body or b"").decode()
            return [resp.status_code, _first_line(text).split(" (event_id")[0]]
        return _clean(resp)

    def ensemble(**kw):
        fn = api.new_function("e", kind="serving")
        graph = fn.set_topology("router", api.VotingEnsemble(**kw))
        graph.add_route("m1", class_name="ModelTestingClass", model_path=".", multiplier=1)
        graph.add_route("m2", class_name="ModelTestingClass", model_path=".", multiplier=3)
        return fn.to_mock_server(namespace=ns)

    for tag, kw in (("weights_unknown_name", dict(vote_type="regression", weights={"m1": 0.7, "zz": 0.3})),
                    ("bad_vote_type", dict(vote_type="banana"))):
        try:
            server = ensemble(**kw)
            out[tag] = [call(server, "/v2/models/infer", {"inputs": [5]}), call(server, "/v2/models/", None, method="GET")]
        except Exception as exc:  # noqa: BLE001
            out[tag] = f"{type(exc).__name__}: {_first_line(exc)}"
    server = ensemble(vote_type="regression")
    out["no_inputs"] = call(server, "/v2/models/infer", {"x": 1})
    out["inputs_not_list"] = call(server, "/v2/models/infer", {"inputs": 5})
    out["empty_inputs"] = call(server, "/v2/models/infer", {"inputs": []})
    out["explain_on_ensemble"] = call(server, "/v2/models/explain", {"inputs": [5]})
    out["unknown_op"] = call(server, "/v2/models/m1/zzz", {"inputs": [5]})
    out["get_on_model"] = call(server, "/v2/models/m1", None, method="GET")
    out["ready"] = call(server, "/v2/models/m1/ready", None, method="GET")
    out["metrics_path"] = call(server, "/v2/models/m1/metrics", None, method="GET")
    return out


def flow_odd_cases(api):
    """serving/states.py:564-599, 1292-1323 -- a sync flow in odd situations: a step returning None, a step terminating the
    event, graph-level and step-level error handlers, nested input / result paths (present, missing, on a scalar body), a
    responder in the middle, an empty flow"""
    ns = dict(make_namespace(api))

    def ret_none(x):
        return None

    def terminate(event):
        event.terminated = True
        event.body = {"stopped": event.body}
        return event

    def add_one(x):
        return x + 1

    def fail(x):
        raise KeyError("bad key")

    def catch(event):
        return {"caught": str(event.error), "origin": event.origin_state, "body": event.body}

    ns.update(ret_none=ret_none, terminate=terminate, add_one=add_one, fail=fail, catch=catch)

    def run(build, body):
        fn = api.new_function("f", kind="serving")
        try:
            build(fn.set_topology("flow", engine="sync"))
            resp = fn.to_mock_server(namespace=ns).test(body=body, silent=True)
        except Exception as exc:  # noqa: BLE001
            return f"{type(exc).__name__}: {_first_line(exc)}"
        if hasattr(resp, "status_code"):
            text = resp.body if isinstance(resp.body, str) else (resp.body or b"").decode()
            return [resp.status_code, _first_line(text)]
        return _clean(resp)

    def graph_handler(g):
        g.to(name="a", handler="add_one").to(name="f", handler="fail").to(name="b", handler="add_one")
        g.error_handler(name="catcher", handler="catch", full_event=True)

    def step_handler(g):
        g.to(name="a", handler="add_one").to(name="f", handler="fail").error_handler(
            name="catcher", handler="catch", full_event=True).to(name="after", handler="(event)")

    def paths(g):
        g.to(name="a", handler="add_one", input_path="x.y", result_path="z.w")

    return {
        "none_result": run(lambda g: g.to(name="a", handler="add_one").to(name="n", handler="ret_none").to(name="b", handler="add_one"), 1),
        "terminated": run(lambda g: g.to(name="a", handler="add_one").to(name="t", handler="terminate", full_event=True).to(
            name="b", handler="add_one"), 1),
        "graph_error_handler": run(graph_handler, 1),
        "step_error_handler": run(step_handler, 1),
        "nested_paths": run(paths, {"x": {"y": 4}}),
        "nested_paths_missing": run(paths, {"x": {}}),
        "result_path_on_scalar_body": run(lambda g: g.to(name="a", handler="add_one", result_path="r"), 5),
        "respond_midway_sync": run(lambda g: g.to(name="a", handler="add_one").respond().to(name="b", handler="add_one"), 1),
        "empty_flow": run(lambda g: None, 1),
    }


flow_odd_cases.EXPECT = {("terminated",): {"stopped": 2}, ("nested_paths",): {"x": {"y": 4}, "z": {"w": 5}}}


def steps_odd_values(api):
    """feature_store/steps.py:152-216, 377-513, 516-602, 699-735 on dict events with odd values: what counts as missing for
    Imputer, bool / None / float values against OneHotEncoder categories, MapValues on None / strings / NaN, DateExtractor on
    bad input, DropFeatures listing a feature twice"""
    import pandas as pd

    def run(make, ev):
        try:
            res = make().do(dict(ev))
            return _clean({k: (None if v is pd.NaT else v) for k, v in res.items()})
        except Exception as exc:  # noqa: BLE001
            return f"{type(exc).__name__}: {_first_line(exc)}"

    nan = float("nan")
    return {
        "imputer_kinds": run(lambda: api.Imputer(mapping={"a": 1, "b": 2, "c": 3, "d": 4, "e": 5}, default_value=9),
                             {"a": None, "b": nan, "c": pd.NaT, "d": "nan", "e": np.float32("nan"), "f": np.nan, "g": 0, "h": ""}),
        "imputer_none_default": run(lambda: api.Imputer(), {"a": None, "b": nan, "c": 1}),
        "onehot_kinds": run(lambda: api.OneHotEncoder(mapping={"c": [0, 1, "x y-z", True]}), {"c": True, "k": 1}),
        "onehot_none": run(lambda: api.OneHotEncoder(mapping={"c": [0, 1]}), {"c": None}),
        "onehot_float_value": run(lambda: api.OneHotEncoder(mapping={"c": [0, 1, 2]}), {"c": 2.0}),
        "onehot_missing_feature": run(lambda: api.OneHotEncoder(mapping={"zz": [0, 1]}), {"c": 1}),
        "onehot_float_cats": run(lambda: api.OneHotEncoder(mapping={"c": [0.5, 1]}), {"c": 1}),
        "map_missing_feature": run(lambda: api.MapValues(mapping={"zz": {1: 2}}), {"a": 1}),
        "map_none_value": run(lambda: api.MapValues(mapping={"a": {None: 5, 1: 2}}), {"a": None}),
        "map_range_on_string": run(lambda: api.MapValues(mapping={"a": {"ranges": {0: [0, 5]}}}), {"a": "text"}),
        "map_range_nan": run(lambda: api.MapValues(mapping={"a": {"ranges": {0: ["-inf", "inf"]}}}), {"a": nan}),
        "map_suffix": run(lambda: api.MapValues(mapping={"a": {1: 2}}, with_original_features=True, suffix="m"), {"a": 1, "b": 3}),
        "date_bad": run(lambda: api.DateExtractor(parts=["hour"]), {"timestamp": "not a date"}),
        "date_missing_col": run(lambda: api.DateExtractor(parts=["hour"], timestamp_col="when"), {"timestamp": "2021-01-01"}),
        "date_parts": run(lambda: api.DateExtractor(parts=["year", "quarter", "is_leap_year", "day_of_year", "minute"]),
                          {"timestamp": "2024-02-29 13:45:10"}),
        "date_unknown_part": run(lambda: api.DateExtractor(parts=["fortnight"]), {"timestamp": "2024-02-29"}),
        "drop_twice": run(lambda: api.DropFeatures(features=["a", "a"]), {"a": 1, "b": 2}),
    }


def steps_random_events(api):
    """feature_store/steps.py:152-246 (MapValues), 377-413 (Imputer), 427-513 (OneHotEncoder), 699-735 (DropFeatures) on 240
    seeded random dict events -- the per-event semantics the fused kernels restate, away from hand-picked cases: missing values
    of every kind the reference recognises (None / nan / float32 nan) and +-inf beside ordinary floats and ints, categorical
    values inside and outside the encoder's categories (ints, floats equal to ints, strings with sanitised characters),
    range tables whose bounds are hit exactly, value maps over ints; every step alone and the chain Imputer -> OneHotEncoder
    -> MapValues -> DropFeatures, one `do(event)` per event"""
    import random

    rnd = random.Random(20260922)
    nan = float("nan")
    num_cols = [f"n{i}" for i in range(6)]
    cat_cols = ["c0", "c1", "c2"]
    cats = {"c0": [0, 1, 2, 3], "c1": ["red", "dark green", "blue-ish", "x"], "c2": [10, 20, 30]}
    bounds = [-2.0, -0.5, 0.0, 0.5, 2.0]

    def number():
        r = rnd.random()
        if r < 0.08:
            return None
        if r < 0.16:
            return nan
        if r < 0.19:
            return np.float32("nan")
        if r < 0.23:
            return rnd.choice([float("inf"), float("-inf")])
        if r < 0.45:
            return rnd.choice(bounds)  # exactly on a range bound
        if r < 0.6:
            return rnd.randint(-3, 3)
        return round(rnd.uniform(-3, 3), 3)

    def category(col):
        r = rnd.random()
        if r < 0.7:
            v = rnd.choice(cats[col])
            return float(v) if isinstance(v, int) and rnd.random() < 0.2 else v  # 2.0 == 2 for the encoder
        if r < 0.8:
            return None
        return rnd.choice([99, "unknown", -1, 2.5])

    events = []
    for _ in range(240):
        ev = {c: number() for c in num_cols}
        ev.update({c: category(c) for c in cat_cols})
        ev["key"] = rnd.randint(0, 5)
        events.append(ev)

    imputer = dict(mapping={"n0": 0.25, "n1": -1, "c0": 1, "c1": "x"}, default_value=7)
    onehot = dict(mapping=cats)
    ranges = {"lowest": ["-inf", -2.0], "low": [-2.0, -0.5], "mid": [-0.5, 0.5], "high": [0.5, 2.0], "highest": [2.0, "inf"]}
    mapval = dict(mapping={"n2": {"ranges": {k: list(v) for k, v in ranges.items()}}, "n3": {"ranges": {0: [-1, 1], 1: [0, 3]}},
                           "key": {0: 100, 1: 101, 2: 102}})

    def run(steps, ev):
        try:
            body = dict(ev)
            for step in steps:
                body = step.do(body)
            return _clean(body)
        except Exception as exc:  # noqa: BLE001
            return f"{type(exc).__name__}: {_first_line(exc)}"

    def mk_chain():
        return [api.Imputer(**imputer), api.OneHotEncoder(**onehot), api.MapValues(with_original_features=True, **mapval),
                api.DropFeatures(features=["n5", "key"])]

    singles = {"imputer": [api.Imputer(**imputer)], "imputer_no_default": [api.Imputer(mapping=imputer["mapping"])],
               "onehot": [api.OneHotEncoder(**onehot)], "mapval": [api.MapValues(**mapval)],
               "mapval_with_originals": [api.MapValues(with_original_features=True, **mapval)]}
    out = {name: [run(steps, ev) for ev in events] for name, steps in singles.items()}
    chain = mk_chain()
    out["chain"] = [run(chain, ev) for ev in events]
    return out


def flow_error_handler_paths(api):
    """serving/states.py:592-597 -- a step that raises under a step-level error handler AND a result_path: the handler runs
    first (it may replace event.body), the step's result_path merge then reads the body the handler left; same with an
    input_path the body does not have, and with a handler that only mutates the body (found by tests/golden/diff_flow_graphs.py)"""

    class Boom:
        def do(self, x):
            raise ValueError(f"boom on {type(x).__name__}")

    class Replace:
        def do_event(self, event):
            event.body = {"handled": str(event.error), "origin": event.origin_state}
            return event

    class Mutate:
        def do_event(self, event):
            event.body["seen"] = sorted(event.error)
            return event

    ns = {"Boom": Boom, "Replace": Replace, "Mutate": Mutate}
    out = {}
    for handler in ("Replace", "Mutate"):
        for result_path in (None, "res", "x.res"):
            for input_path in (None, "x", "nope.deep"):
                fn = api.new_function("f", kind="serving")
                flow = fn.set_topology("flow", engine="sync")
                kw = {k: v for k, v in (("result_path", result_path), ("input_path", input_path)) if v}
                flow.to("Boom", name="s0", **kw).error_handler(name="catch", class_name=handler)
                server = fn.to_mock_server(namespace=ns)
                try:
                    body = server.test(body={"x": {"y": 3}, "q": 2}, silent=True)
                    text = json.dumps(_resp(body), sort_keys=True, default=lambda o: f"<{type(o).__name__}>")
                    out[f"{handler}|{result_path}|{input_path}"] = text
                except Exception as exc:  # noqa: BLE001
                    out[f"{handler}|{result_path}|{input_path}"] = f"{type(exc).__name__}: {_first_line(exc)}"
    return out


def error_text_shapes(api):
    """mlrun/errors.py:126-149 through serving/server.py:278-288 -- how a failing step's exception reads in the 400 response:
    no message (the repr stands in), an empty message, a chain of causes, a cause chain that loops back, a message beyond
    32 000 characters (cut to its two ends); found by tests/golden/diff_flow_with_routers.py"""

    class Raise:
        def __init__(self, how="bare", **kw):
            self.how = how

        def do(self, x):
            if self.how == "bare":
                raise NotImplementedError()
            if self.how == "empty":
                raise ValueError("")
            if self.how == "chain":
                try:
                    try:
                        raise KeyError("inner")
                    except KeyError as inner:
                        raise RuntimeError() from inner
                except RuntimeError as middle:
                    raise ValueError("outer") from middle
            if self.how == "loop":
                a, b = ValueError("a"), TypeError("b")
                a.__cause__, b.__cause__ = b, a
                raise a
            raise ValueError("x" * 20000 + "MIDDLE" + "y" * 20000)

    out = {}
    for how in ("bare", "empty", "chain", "loop", "long"):
        fn = api.new_function("f", kind="serving")
        fn.set_topology("flow", engine="sync").to("Raise", name="s", how=how).respond()
        resp = fn.to_mock_server(namespace={"Raise": Raise}).test(body={"a": 1}, silent=True)
        text = resp.body if isinstance(resp.body, str) else resp.body.decode()
        out[how] = {"status": resp.status_code, "length": len(text), "head": text[:80], "tail": text[-40:], "has_cut": "...truncated..." in text}
    return out


def model_numpy_outputs(api):
    """serving/server.py:298-308 + serving/v2_serving.py:228-342 -- a model returning numpy values: fine for `server.test`
    (the body object comes back), a TypeError on the wire (`GraphServer.run` json-encodes strictly); plus odd requests to a
    model: GET on infer, a custom `op_` handler, an unknown version, a trailing slash, a body naming another model"""

    class NpModel(api.V2ModelServer):
        def load(self):
            pass

        def predict(self, request):
            x = np.asarray(request["inputs"], dtype=np.float64)
            return {"array": x * 2, "scalar": np.float32(x.sum()), "list_np": [np.int64(1), np.float64(2.5)],
                    "nested": (x * 2).tolist(), "tuple": (1, 2)}[self.get_param("kind")]

        def op_echo_headers(self, event):
            return {"method": event.method, "path": event.path, "ct": event.content_type}

    kinds = ("array", "scalar", "list_np", "nested", "tuple")
    fn = api.new_function("m", kind="serving")
    fn.set_topology("router")
    for kind in kinds:
        fn.add_model(kind, ".", class_name="NpModel", kind=kind)
    server = fn.to_mock_server(namespace={"NpModel": NpModel})
    out = {}

    def call(path, body=None, **kw):
        resp = server.test(path, body, silent=True, **kw)
        if hasattr(resp, "status_code"):
            text = resp.body if isinstance(resp.body, str) else (resp.body or b"").decode()
            return [resp.status_code, _first_line(text)]
        return _clean(resp)

    for kind in kinds:
        out[f"object_{kind}"] = call(f"/v2/models/{kind}/infer", {"inputs": [[1, 2], [3, 4]]})
        try:
            resp = server.run(api.MockEvent(body=json.dumps({"inputs": [[1, 2]]}), path=f"/v2/models/{kind}/infer"))
            out[f"wire_{kind}"] = [resp.status_code, _clean(json.loads(resp.body))]
        except Exception as exc:  # noqa: BLE001
            out[f"wire_{kind}"] = f"raised {type(exc).__name__}: {_first_line(exc)}"
    out["id_passthrough"] = server.test("/v2/models/nested/infer", {"id": "req-7", "inputs": [[1]]})["id"]
    out["get_infer"] = call("/v2/models/nested/infer", None, method="GET")
    out["custom_op"] = call("/v2/models/nested/echo_headers", {"x": 1}, content_type="application/json")
    out["custom_op_get"] = call("/v2/models/nested/echo_headers", None, method="GET")
    out["version_in_url_unknown"] = call("/v2/models/nested/versions/v9/infer", {"inputs": [[1]]})
    out["trailing_slash"] = call("/v2/models/nested/infer/", {"inputs": [[1]]})
    out["model_in_body_wrong_url"] = call("/v2/models/array/infer", {"model": "nested", "inputs": [[1]]})
    return out


model_numpy_outputs.EXPECT = {("wire_array",): "raised TypeError: Object of type ndarray is not JSON serializable",
                              ("object_scalar", "outputs"): 10.0}


def vote_odd_predictions(api):
    """serving/routers.py:708-810 -- the vote on odd child predictions: ties, float / negative / string / large labels, custom
    and non-normalised weights, None and ragged predictions, vote-type inference from the first request's values"""

    class Fixed(api.V2ModelServer):
        def load(self):
            pass

        def predict(self, request):
            return list(self.get_param("preds"))

    def ens(preds_list, **kw):
        fn = api.new_function("e", kind="serving")
        graph = fn.set_topology("router", api.VotingEnsemble(**kw))
        for i, preds in enumerate(preds_list):
            graph.add_route(f"m{i}", class_name="Fixed", model_path=".", preds=preds)
        try:
            server = fn.to_mock_server(namespace={"Fixed": Fixed})
            resp = server.test("/v2/models/infer", {"inputs": [[0]] * len(preds_list[0])}, silent=True)
        except Exception as exc:  # noqa: BLE001
            return f"raised {type(exc).__name__}: {_first_line(exc)}"
        if hasattr(resp, "status_code"):
            text = resp.body if isinstance(resp.body, str) else (resp.body or b"").decode()
            return [resp.status_code, _first_line(text)]
        return _clean(resp)["outputs"]

    return {
        "cls_basic": ens([[1, 2, 0], [1, 0, 0], [2, 2, 1]], vote_type="classification"),
        "cls_tie_lowest": ens([[0, 3], [1, 2]], vote_type="classification"),
        "cls_float_labels": ens([[1.0, 2.0], [1.0, 0.0], [0.0, 2.0]], vote_type="classification"),
        "cls_negative": ens([[-1, 2], [1, 2]], vote_type="classification"),
        "cls_weighted": ens([[0, 0], [1, 1], [1, 0]], vote_type="classification", weights={"m0": 0.6, "m1": 0.2, "m2": 0.2}),
        "cls_strings": ens([["a", "b"], ["a", "a"]], vote_type="classification"),
        "cls_big_labels": ens([[1000, 2], [1000, 3]], vote_type="classification"),
        "reg_basic": ens([[1.0, 2.0], [3.0, 5.0]], vote_type="regression"),
        "reg_ints": ens([[1, 2], [3, 5]], vote_type="regression"),
        "reg_weights_not_one": ens([[1.0, 2.0], [3.0, 5.0]], vote_type="regression", weights={"m0": 2, "m1": 2}),
        "reg_none": ens([[1.0, None], [3.0, 5.0]], vote_type="regression"),
        "reg_single_model": ens([[4, 5]], vote_type="regression"),
        "reg_ragged": ens([[1, 2], [1]], vote_type="regression"),
        "inferred_from_ints": ens([[1, 2], [1, 0]]),
        "inferred_from_floats": ens([[1.5, 2.0], [1.0, 0.0]]),
        "inferred_from_integer_floats": ens([[1.0, 2.0], [1.0, 0.0]]),
    }


vote_odd_predictions.EXPECT = {("cls_tie_lowest",): [0, 2], ("reg_weights_not_one",): [8.0, 14.0], ("inferred_from_integer_floats",): [1, 0]}


def graph_serialisation(api):
    """serving/states.py:89-154, 297-362, 940-1001 + runtimes/nuclio/serving.py:645-666 -- `to_dict` of a flow with paths,
    a queue, a router inside the flow, an error handler and an `after` step; of a router topology; the keys of the serving
    spec the function hands to the server"""
    fn = api.new_function("f", kind="serving")
    g = fn.set_topology("flow", engine="async")
    s1 = g.to("Echo", "s1", input_path="a", result_path="b", full_event=True, custom=1)
    queue = s1.to("$queue", "q1", path="v3io:///x", shards=2)
    router = queue.to("*", "router", function="child")
    router.add_route("m1", class_name="ModelTestingClass", model_path=".", multiplier=2)
    router.add_route("m2:v3", handler="my_hnd")
    router.to(name="post", handler="json.dumps").respond()
    s1.error_handler(name="eh", class_name="EchoError", full_event=True)
    g.add_step(name="side", handler="(event)", after="s1")
    out = {"flow": g.to_dict(), "step_order": list(g.steps.keys())}
    fn2 = api.new_function("r", kind="serving")
    rt = fn2.set_topology("router", "mlrun.serving.ModelRouter", url_prefix="/x")
    rt.add_route("a", class_name="ModelTestingClass", model_path=".", multiplier=1)
    out["router"] = rt.to_dict()
    spec = json.loads(fn._get_serving_spec())
    out["serving_spec_keys"] = sorted(spec.keys())
    out["serving_spec_graph_is_to_dict"] = spec["graph"] == json.loads(json.dumps(out["flow"], default=str))
    return json.loads(json.dumps(out, default=str))


def router_and_model_paths(api):
    """serving/routers.py:86-211 + serving/v2_serving.py:228-342 + serving/states.py:564-599 -- input / result paths on a
    router and on a model route, and a router in the middle of a sync flow addressed through the flow (model urls, the
    default route, the model list, a foreign prefix)"""
    ns = make_namespace(api)

    def call(server, path, body, **kw):
        resp = server.test(path, body, silent=True, **kw)
        if hasattr(resp, "status_code"):
            text = resp.body if isinstance(resp.body, str) else (resp.body or b"").decode()
            return [resp.status_code, _first_line(text)]
        return _clean(resp)

    out = {}
    fn = api.new_function("r", kind="serving")
    graph = fn.set_topology("router", api.ModelRouter(input_path="req", result_path="res"))
    graph.add_route("m1", class_name="ModelTestingClass", model_path=".", multiplier=2)
    server = fn.to_mock_server(namespace=ns)
    out["router_paths"] = call(server, "/v2/models/m1/infer", {"req": {"inputs": [5]}, "keep": 1})
    out["router_paths_missing"] = call(server, "/v2/models/m1/infer", {"inputs": [5]})
    fn = api.new_function("r", kind="serving")
    graph = fn.set_topology("router")
    graph.add_route("m1", class_name="ModelTestingClass", model_path=".", multiplier=2, input_path="a.b", result_path="out")
    server = fn.to_mock_server(namespace=ns)
    out["model_paths"] = call(server, "/v2/models/m1/infer", {"a": {"b": {"inputs": [4]}}, "z": 0})
    out["model_meta"] = call(server, "/v2/models/m1", None, method="GET")
    fn = api.new_function("f", kind="serving")
    flow = fn.set_topology("flow", engine="sync")
    router = flow.to("Echo", "pre").to("*", "router")
    router.add_route("m1", class_name="ModelTestingClass", model_path=".", multiplier=3)
    router.add_route("m2", class_name="ModelTestingClass", model_path=".", multiplier=5)
    router.to("Echo", "post").respond()
    server = fn.to_mock_server(namespace=ns)
    out["flow_router_m2"] = call(server, "/v2/models/m2/infer", {"inputs": [2]})
    out["flow_router_default"] = call(server, "/", {"inputs": [2]})
    out["flow_router_list"] = call(server, "/v2/models/", None, method="GET")
    out["flow_router_bad_prefix"] = call(server, "/other/m1/infer", {"inputs": [2]})
    return out


router_and_model_paths.EXPECT = {("router_paths", "res", "outputs"): 10, ("flow_router_default", "outputs"): 6}


def merger_logic(api):
    """serving/merger.py:36-156 -- the join itself, driven directly: post_init, then a sequence of arrivals through
    `_merge_events` (full events joined on event.id with a window of 3 keys; bodies joined on a key expression)"""
    import types

    log, errors = [], []
    logger = types.SimpleNamespace(warning=lambda m, **k: log.append(("warning", str(m).split("<")[0])),
                                   info=lambda m, **k: log.append(("info", str(m))))
    ctx = types.SimpleNamespace(verbose=False, logger=logger,
                                push_error=lambda event, message, source=None, **k: errors.append([event.id, message, source]))
    out = {}

    def ev(i, body):
        return types.SimpleNamespace(id=i, body=body)

    m = api.Merge(context=ctx, name="Merge", max_behind=3, expected_num_events=2)
    m.post_init()
    trace = []
    for i, body in [("a", 1), ("b", 2), ("a", 3), ("c", 4), ("d", 5), ("e", 6), ("b", 7), ("c", 8), ("f", 9), ("d", 10),
                    ("e", 11), ("f", 12), ("a", 13), ("a", 14)]:
        res = m._merge_events(ev(i, body))
        trace.append(None if res is None else [res.id, res.body])
    out["by_id"] = {"trace": trace, "errors": errors, "log": [list(x) for x in log]}

    log.clear()
    m3 = api.Merge(context=ctx, name="M3", key_path="event['key']", expected_num_events=3)
    m3.post_init()
    trace = []
    for body in [{"key": 7, "x": 1}, {"key": 8, "x": 2}, {"key": 7, "x": 3}, {"key": 7, "x": 4}, {"key": 8, "x": 5}, {"key": 8, "x": 6}]:
        trace.append(m3._merge_events(body))
    out["by_key"] = {"trace": trace, "full_event": bool(m3._full_event)}

    single = api.Merge(context=ctx, name="one", expected_num_events=1)
    single.post_init()
    e = ev("z", 5)
    out["single_uplink_passes"] = single._merge_events(e) is e
    try:
        m3._merge_events({"x": 1})
        out["missing_key"] = None
    except Exception as exc:  # noqa: BLE001
        out["missing_key"] = type(exc).__name__
    return out


merger_logic.EXPECT = {
    ("by_id", "trace", 2): ["a", [1, 3]],
    ("by_id", "trace", 6): None,       # "b" was given up when "e" opened (window of 3): its second part is ignored
    ("by_id", "trace", 7): ["c", [4, 8]],
    ("by_key", "trace", 3): [{"key": 7, "x": 1}, {"key": 7, "x": 3}, {"key": 7, "x": 4}],
    ("missing_key",): "KeyError",
}


def online_service_logic(api):
    """feature_store/feature_vector.py:903-1067 -- OnlineVectorService.initialize / get driven directly.  The online-store
    read (a storey QueryByKey graph over the NoSQL target) is the one thing replaced: a dict lookup that answers an entity
    row with the row itself joined with what the table holds for its key (`api.online_service` wires it under each API)."""
    import pandas as pd

    feats = ["f0", "f1", "f2", "f3", "y"]
    table = {
        ("GOOG",): {"f0": 1.5, "f1": 2.0, "f2": -0.25, "f3": 8.0, "y": 1.0},
        ("MSFT",): {"f0": float("nan"), "f1": float("inf"), "f2": None, "f3": float("-inf"), "y": 0.0},
        ("ZERO",): {"f0": 0.0, "f1": 0.0, "f2": 0.0, "f3": 0.0, "y": 0.0},
        ("PART",): {"f0": 4.0, "f3": 0.5},
        ("INTS",): {"f0": 3, "f1": 0, "f2": float("nan"), "f3": 7, "y": 2},
    }
    stats = pd.DataFrame({"mean": [2.0, 0.75, -1.5, 3.25, 0.5], "min": [0.0, -2.0, -4.0, 0.5, 0.0], "max": [4.0, 2.0, 0.0, 8.0, 2.0],
                          "std": [1.0, 0.5, 0.25, 2.0, 0.125], "count": [5.0, 5.0, 5.0, 5.0, 5.0]}, index=feats)

    def norm(v):
        if isinstance(v, float) and v != v:
            return "nan"
        if isinstance(v, float) and v in (float("inf"), float("-inf")):
            return repr(v)
        if isinstance(v, dict):
            return {k: norm(x) for k, x in v.items()}
        if isinstance(v, list):
            return [norm(x) for x in v]
        return v.item() if hasattr(v, "item") else v

    def attempt(fn):
        try:
            return norm(fn())
        except Exception as exc:  # noqa: BLE001
            return {"raised": type(exc).__name__, "message": _first_line(str(exc))}

    asks = [["GOOG"], ["MSFT"], ["nobody"], ["ZERO"], ["PART"], ["INTS"]]
    out = {}
    for tag, policy in (("none", None), ("mean", {"*": "$mean"}), ("mixed", {"*": 0.5, "f1": "$max", "f2": -3}),
                        ("one", {"f3": "$min"}), ("zero", {"*": 0})):
        svc = api.online_service(feats, ["ticker"], table, stats, "y", False, policy)
        out[tag] = {"impute_values": norm(dict(svc._impute_values)),
                    "lists": attempt(lambda: svc.get(asks, as_list=True)),
                    "dicts": attempt(lambda: svc.get([{"ticker": k[0]} for k in asks])),
                    "one_dict": attempt(lambda: svc.get({"ticker": "GOOG"})),
                    "extra_column": attempt(lambda: svc.get([{"ticker": "nobody", "note": "x"}, {"ticker": "PART", "note": 7}]))}
    with_idx = api.online_service(feats, ["ticker"], table, stats, "y", True, {"*": "$mean"})
    out["with_indexes"] = {"dicts": attempt(lambda: with_idx.get([{"ticker": "MSFT"}, {"ticker": "ZERO"}, {"ticker": "nobody"}])),
                           "lists": attempt(lambda: with_idx.get([["PART"]], as_list=True))}
    no_label = api.online_service(feats[:4], ["ticker"], table, stats, None, False, {"*": "$max"})
    out["no_label"] = attempt(lambda: no_label.get([["GOOG"], ["PART"]], as_list=True))
    two = api.online_service(feats, ["a", "b"], {("x", 1): {"f0": 1.0, "f1": 2.0, "f2": 3.0, "f3": 4.0, "y": 1.0}}, stats, "y", False, None)
    out["composite"] = {"hit_and_miss": attempt(lambda: two.get([["x", 1], ["x", 2]], as_list=True)),
                        "short_row": attempt(lambda: two.get([["x"]]))}
    plain = api.online_service(feats, ["ticker"], table, stats, "y", False, None)
    out["bad_input"] = {"empty": attempt(lambda: plain.get([])), "scalars": attempt(lambda: plain.get(["GOOG"])),
                        "string": attempt(lambda: plain.get("GOOG")), "tuple_rows": attempt(lambda: plain.get([("GOOG",)]))}
    out["bad_policy"] = {"unknown_feature": attempt(lambda: api.online_service(feats, ["ticker"], table, stats, "y", False, {"zz": 1})),
                         "label_in_policy": attempt(lambda: api.online_service(feats, ["ticker"], table, stats, "y", False, {"y": 1}))}
    return out


online_service_logic.EXPECT = {
    ("mean", "lists", 2): None,  # unknown entity: the graph returned only the entity columns
    ("none", "lists", 3): None,  # QUIRK: a row whose values are all falsy (0.0) is reported as missing
}


def enrichment_routers(api):
    """serving/routers.py:1118-1196, 1199-1342 -- EnrichmentModelRouter / EnrichmentVotingEnsemble served: entity keys in
    `inputs` become feature vectors (OnlineVectorService.get(as_list=True)) before the child models see them.  The
    online-store read is the stub of `online_service_logic` (`api.register_online_vector`)."""
    import pandas as pd
    from sklearn.linear_model import LinearRegression

    feats = ["f0", "f1", "f2", "f3"]
    table = {("GOOG",): {"f0": 1.5, "f1": 2.0, "f2": -0.25, "f3": 8.0}, ("MSFT",): {"f0": float("nan"), "f1": 0.5, "f2": None, "f3": 1.0},
             ("AMZN",): {"f0": 3.0, "f1": float("inf"), "f2": 2.0, "f3": -1.0}}
    stats = pd.DataFrame({"mean": [2.0, 0.75, -1.5, 3.25], "max": [4.0, 2.0, 0.0, 8.0]}, index=feats)
    api.register_online_vector("store://vectors/quotes", feats, ["ticker"], table, stats, None, False)
    coefs = [[1.0, 2.0, 0.5, -1.0], [0.25, 0.0, 4.0, 1.0], [-2.0, 1.0, 1.0, 0.125]]
    ns = {"SKLearnModelServer": api.SKLearnModelServer}

    def served(router):
        fn = api.new_function("enrich", kind="serving")
        graph = fn.set_topology("router", router)
        for i, c in enumerate(coefs):
            m = LinearRegression()
            m.coef_, m.intercept_, m.n_features_in_ = np.asarray(c, dtype=np.float64), 0.5 * i, 4
            graph.add_route(f"m{i}", class_name="SKLearnModelServer", model=m, model_path="")
        return fn.to_mock_server(namespace=ns)

    def call(server, path, body):
        resp = server.test(path, body=body, silent=True)
        if hasattr(resp, "status_code"):
            return {"status": resp.status_code, "error": _first_line(resp.body if isinstance(resp.body, str) else resp.body.decode()).split(":")[0]}
        return _clean(resp)

    out = {}
    ens = served(api.EnrichmentVotingEnsemble(feature_vector_uri="store://vectors/quotes", impute_policy={"*": "$mean", "f1": "$max"},
                                              vote_type="regression", executor_type="array"))
    out["ensemble"] = {
        "lists": call(ens, "/v2/models/infer", {"inputs": [["GOOG"], ["MSFT"], ["AMZN"]]}),
        "dict_rows": call(ens, "/v2/models/infer", {"inputs": [{"ticker": "AMZN"}]}),
        "text_body": call(ens, "/v2/models/infer", json.dumps({"inputs": [["GOOG"]]})),
        "one_model": call(ens, "/v2/models/m1/infer", {"inputs": [["GOOG"], ["AMZN"]]}),
        "unknown_entity": call(ens, "/v2/models/infer", {"inputs": [["GOOG"], ["nobody"]]}),
    }
    raw = served(api.EnrichmentModelRouter(feature_vector_uri="store://vectors/quotes"))
    out["router_no_policy"] = {
        "clean_row": call(raw, "/v2/models/m0/infer", {"inputs": [["GOOG"]]}),
        "nan_row": call(raw, "/v2/models/m0/infer", {"inputs": [["MSFT"]]}),
        "default_route": call(raw, "/v2/models/infer", {"inputs": [["GOOG"]]}),
    }
    fixed = served(api.EnrichmentModelRouter(feature_vector_uri="store://vectors/quotes", impute_policy={"*": 0.5}))
    out["router_constant_policy"] = call(fixed, "/v2/models/m2/infer", {"inputs": [["MSFT"], ["AMZN"], ["GOOG"]]})
    return out


enrichment_routers.EXPECT = {
    # GOOG = [1.5, 2.0, -0.25, 8.0]: m0 = 1.5 + 4 - 0.125 - 8 = -2.625; m1 = 0.375 - 1 + 8 + 0.5 = 7.875; m2 = -3 + 2 - 0.25 + 1 + 1 = 0.75
    ("router_no_policy", "clean_row", "outputs"): [-2.625],
    ("ensemble", "one_model", "outputs", 0): 7.875,
}


def flow_add_model(api):
    """tests/serving/test_flow.py:189-213 -- fn.add_model inside a flow: needs a router; the only router is found by itself,
    one of several is named with router_step"""
    make_namespace(api)
    out = {}

    def attempt(fn, **kw):
        try:
            fn.add_model("m1", class_name="ModelTestingClass", model_path=".", **kw)
            return None
        except Exception as exc:  # noqa: BLE001
            return type(exc).__name__

    fn = api.new_function("tests", kind="serving")
    fn.set_topology("flow", engine="sync").to("Echo", "e1").to("Echo", "e2")
    out["no_router"] = attempt(fn)
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to("Echo", "e1").to("*", "router").to("Echo", "e2")
    out["one_router"] = [attempt(fn), sorted(graph["router"].routes)]
    fn = api.new_function("tests", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to("Echo", "e1").to("*", "r1").to("Echo", "e2").to("*", "r2")
    out["named_router"] = [attempt(fn, router_step="r2"), sorted(graph["r1"].routes), sorted(graph["r2"].routes)]
    out["two_routers_unnamed"] = attempt(fn)
    out["unknown_router"] = attempt(fn, router_step="r9")
    return out


flow_add_model.EXPECT = {("one_router",): [None, ["m1"]], ("named_router",): [None, [], ["m1"]]}


def module_load(api):
    """tests/serving/test_flow.py:330-350 -- classes and handlers of the function's own code file (`command=`) are found
    without a namespace; the handler gets the graph context"""
    import os
    import tempfile

    code = (
        "class MyCls:\n"
        "    def __init__(self, context=None, name=None, **kwargs):\n"
        "        self.context, self.name = context, name\n"
        "    def do(self, event):\n"
        "        return event * 2\n"
        "\n"
        "def myhand(x, context=None):\n"
        "    assert context is not None and hasattr(context, 'logger'), 'did not get a context'\n"
        "    return x * 2\n"
    )
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "graft_myfunc.py")
        with open(path, "w") as fp:
            fp.write(code)
        fn = api.new_function("test2", command=path, kind="serving")
        graph = fn.set_topology("flow", engine="sync")
        graph.to(name="s1", class_name="MyCls").to(name="s2", handler="myhand")
        out["from_command"] = fn.to_mock_server().test(body=5)
        rel = api.new_function("test3", command="graft_myfunc.py", kind="serving")
        rel.set_topology("flow", engine="sync").to(name="s1", class_name="MyCls").to(name="s2", handler="myhand")
        out["relative_to_workdir"] = rel.to_mock_server(workdir=tmp).test(body=7)
    plain = api.new_function("test4", kind="serving")
    plain.set_topology("flow", engine="sync").to(name="s1", class_name="MyCls")
    try:
        plain.to_mock_server()
        out["without_code"] = None
    except Exception as exc:  # noqa: BLE001
        out["without_code"] = type(exc).__name__
    return out


module_load.EXPECT = {("from_command",): 20, ("relative_to_workdir",): 28}


def infer_dict_ops(api):
    """tests/serving/test_serving.py:543-563 + serving/v2_serving.py:391-426 -- the *_dict operations reorder dict inputs by
    the model artifact's input schema.  The artifact store is out of scope: `load()` sets `model_spec` directly."""
    import types

    fields = ["sepal_length_cm", "sepal_width_cm", "petal_length_cm", "petal_width_cm"]

    class SpecModel(api.V2ModelServer):
        def load(self):
            if self.get_param("with_spec", True):
                self.model_spec = types.SimpleNamespace(inputs=[types.SimpleNamespace(name=f) for f in fields])

        def predict(self, request):
            return [sum(w * v for w, v in zip((1, 10, 100, 1000), row)) for row in request["inputs"]]

    rows = [[5.1, 3.5, 1.4, 0.2], [7.7, 3.8, 6.7, 2.2]]
    shuffled = [{"petal_width_cm": r[3], "sepal_length_cm": r[0], "petal_length_cm": r[2], "sepal_width_cm": r[1]} for r in rows]
    fn = api.new_function("tst", kind="serving")
    fn.set_topology("router")
    fn.add_model("m1", ".", class_name="SpecModel")
    fn.add_model("bare", ".", class_name="SpecModel", with_spec=False)
    server = fn.to_mock_server(namespace={"SpecModel": SpecModel})

    def call(path, body):
        resp = server.test(path, body, silent=True)
        if hasattr(resp, "status_code"):
            text = resp.body if isinstance(resp.body, str) else resp.body.decode()
            return {"status": resp.status_code, "error": _first_line(text)}
        return _clean(resp)

    return {
        "infer": call("/v2/models/m1/infer", {"inputs": rows}),
        "predict": call("/v2/models/m1/predict", {"inputs": rows}),
        "infer_dict": call("/v2/models/m1/infer_dict", {"inputs": shuffled}),
        "predict_dict": call("/v2/models/m1/predict_dict", {"inputs": shuffled}),
        "one_dict": call("/v2/models/m1/infer_dict", {"inputs": {k: [d[k] for d in shuffled] for k in fields}}),
        "lists_to_dict_op": call("/v2/models/m1/infer_dict", {"inputs": rows}),
        "missing_key": call("/v2/models/m1/infer_dict", {"inputs": [{k: 1.0 for k in fields[:3]}]}),
        "no_spec": call("/v2/models/bare/infer_dict", {"inputs": shuffled}),
        "no_spec_plain_infer": call("/v2/models/bare/infer", {"inputs": rows}),
    }


infer_dict_ops.EXPECT = {
    ("infer", "outputs"): [5.1 + 35.0 + 140.0 + 200.0, 7.7 + 38.0 + 670.0 + 2200.0],
}


def no_merger(api):
    """tests/serving/test_merger.py:51-84 -- split into two branches that meet again WITHOUT a Merge step: the common step
    sees every event once per branch (directly, and behind a queue)"""
    ns = dict(make_namespace(api))

    class Gather:
        def __init__(self, context):
            self.context = context
            context.mylist = []

        def do(self, event):
            self.context.mylist.append(event)
            return event

    ns["Gather"] = Gather
    out = {}
    for with_queue in (False, True):
        fn = api.new_function("x", kind="serving")
        graph = fn.set_topology("flow", exist_ok=True)
        dbl = graph.to(name="double", handler="double")
        dbl.to(name="add3", class_name="Adder", add=3)
        dbl.to(name="add2", class_name="Adder", add=2)
        if with_queue:
            graph.add_step("$queue", "q1", path="").after_step("add2", "add3").to("Gather", function="some_function")
        else:
            graph.add_step("Gather").after_step("add2", "add3")
        server = fn.to_mock_server(namespace=ns)
        for data in [5, 10, 15]:
            server.test("", body=data)
        server.wait_for_completion()
        out["queue" if with_queue else "direct"] = sorted(server.context.mylist)
    return out


no_merger.EXPECT = {("direct",): [12, 13, 22, 23, 32, 33], ("queue",): [12, 13, 22, 23, 32, 33]}
no_merger.ASYNC = True


def merge_flows(api):
    """tests/serving/test_merger.py:87-128 -- split and merge through a served async graph (join on event.id, join on a
    body key, a missing key surfacing as the event's error)"""
    ns = make_namespace(api)
    out = {}
    fn = api.new_function("x", kind="serving")
    graph = fn.set_topology("flow", engine="async", exist_ok=True)
    dbl = graph.to(name="double", handler="double")
    dbl.to(name="add3", class_name="Adder", add=3)
    dbl.to(name="add2", class_name="Adder", add=2)
    graph.add_step(api.Merge(name="Merge")).respond().after_step("add2", "add3")
    server = fn.to_mock_server(namespace=ns)
    out["simple"] = [sorted(server.test("", body=5)), sorted(server.test("", body=6))]
    server.wait_for_completion()

    fn = api.new_function("y", kind="serving")
    graph = fn.set_topology("flow", engine="async", exist_ok=True)
    dbl = graph.to(name="double", handler="double", input_path="x", result_path="x")
    dbl.to(name="add3", class_name="Adder", add=3, input_path="x", result_path="x")
    dbl.to(name="add2", class_name="Adder", add=2, input_path="x", result_path="x")
    graph.add_step(api.Merge(name="Merge", key_path="event['key']"), after=["add2", "add3"]).respond()
    server = fn.to_mock_server(namespace=ns)
    try:
        server.test("", body={"x": 4})
        out["missing_key"] = None
    except RuntimeError as exc:
        out["missing_key"] = "KeyError" in str(exc)
    resp = server.test("", body={"x": 4, "key": 77})
    out["custom_key"] = sorted(item["x"] for item in resp)
    server.wait_for_completion()
    return out


merge_flows.EXPECT = {("simple",): [[12, 13], [14, 15]], ("custom_key",): [10, 11], ("missing_key",): True}
merge_flows.ASYNC = True  # a served async graph needs storey: the reference's own test literals pin it instead


# =========================================================================== numeric scenarios (seeded)
def vote_math(api):
    """serving/routers.py:708-741, 746-787 -- the vote kernels' reference arithmetic on seeded arrays"""
    rng = np.random.default_rng(7)
    ens = api.VotingEnsemble(vote_type="classification")
    out = {}
    preds = rng.integers(0, 5, size=(64, 4))
    for name, w in [("equal", np.full(4, 0.25)), ("skew", np.array([0.1, 0.2, 0.3, 0.4])),
                    ("ones", np.ones(4)), ("tie", np.array([0.5, 0.5, 0.0, 0.0]))]:
        out[f"majority_{name}"] = ens._majority_vote(preds.tolist(), w)
    vals = rng.normal(size=(64, 4))
    out["mean_inputs"] = vals.tolist()
    out["majority_inputs"] = preds.tolist()
    for name, w in [("equal", np.full(4, 0.25)), ("skew", np.array([0.1, 0.2, 0.3, 0.4])), ("ones", np.ones(4))]:
        out[f"mean_{name}"] = ens._mean_vote(vals.tolist(), w)
    return out


def flow3_linear_events(api):
    """BASELINE.json configs[1] shape at test size: Imputer -> OneHotEncoder -> linear model, per-event
    dict events through the sync flow (storey-engine step semantics).  16-feat version of §8(d) cfg 2."""
    from mlrun_b200.synthetic import flow3_workload

    wl = flow3_workload(n_rows=48, n_num=12, n_cat=4, seed=2, n_models=1)
    server = wl.build_server(api, engine="sync")
    outs = []
    for row in wl.rows_as_dicts():
        outs.append(_clean(server.test(body=row))["outputs"])
    return {"outputs": outs}


def flow3_ensemble_events(api):
    """metric workload at test size: Imputer -> OneHotEncoder -> VotingEnsemble(4 linear models)"""
    from mlrun_b200.synthetic import flow3_workload

    wl = flow3_workload(n_rows=48, n_num=12, n_cat=4, seed=3, n_models=4)
    server = wl.build_server(api, engine="sync")
    outs = []
    for row in wl.rows_as_dicts():
        outs.append(_clean(server.test(path="/v2/models/infer", body=row))["outputs"])
    return {"outputs": outs}


def tree_ensemble_batch(api):
    """BASELINE.json configs[2] shape at test size: VotingEnsemble router of 4 sklearn GBT regressors,
    one event carrying the whole batch in `inputs` (reference-batched mode)"""
    from mlrun_b200.synthetic import tree_workload

    wl = tree_workload(n_rows=64, n_feat=16, n_models=4, n_trees=8, depth=4, seed=3, kind="regression")
    server = wl.build_server(api)
    reg = _clean(server.test("/v2/models/infer", body={"inputs": wl.X.astype(np.float64).tolist()}))["outputs"]
    single = _clean(server.test("/v2/models/m1/infer", body={"inputs": wl.X[:8].astype(np.float64).tolist()}))["outputs"]
    wlc = tree_workload(n_rows=64, n_feat=16, n_models=4, n_trees=6, depth=3, seed=4, kind="classification")
    server = wlc.build_server(api)
    cls = _clean(server.test("/v2/models/infer", body={"inputs": wlc.X.astype(np.float64).tolist()}))["outputs"]
    return {"regression": reg, "single_model": single, "classification": cls}


SCENARIOS = [
    router_protocol, router_raised_error, ensemble_regression, ensemble_classification, ensemble_weights,
    ensemble_metadata, ensemble_weight_sum_below_one, ensemble_vote_type_inference, router_mock_direct,
    echo_plumbing, tracking, parallel_run, flow_basic_sync, flow_handlers_sync, flow_on_error_sync,
    flow_content_type, flow_model_no_router, flow_multi_function_sync, flow_path_control_sync, step_to_dict,
    route_cap, flow_add_model, module_load, infer_dict_ops, pickle_model_from_path, model_async_load, class_args_protocol, model_hooks, custom_router, server_run_details, tracking_sampling_batching, graph_validation_errors, event_envelope, set_tracking_params, add_model_args, parallel_run_details, ensemble_odd_requests, flow_odd_cases, flow_error_handler_paths, error_text_shapes, steps_odd_values, steps_random_events, model_numpy_outputs, vote_odd_predictions, graph_serialisation, router_and_model_paths, flow_async_basic, flow_async_misc, merger_logic, online_service_logic, enrichment_routers, no_merger, merge_flows, steps_dict_events, steps_pandas_engine, steps_validate_args, validator_events, validator_pandas, set_event_metadata_logic, vote_math,
    flow3_linear_events, flow3_ensemble_events, tree_ensemble_batch,
]

NUMERIC = {"vote_math", "flow3_linear_events", "flow3_ensemble_events", "tree_ensemble_batch"}


def dig(obj, path):
    for key in path:
        obj = obj[key]
    return obj
