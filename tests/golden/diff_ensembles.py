"""Differential check of VotingEnsemble / ParallelRun behind `server.test` against the REAL reference (build container only):
seeded random ensembles -- 1-5 routes whose models return ints, floats, integral floats or mixed values, one of them possibly
raising; vote_type None / classification / regression; optional weights (complete, partial, with an unknown route);
prediction_col_name / format_response_with_col_name_flag; array executor (the order-deterministic one) -- and ParallelRun with
random extend_event and route outputs (dicts with overlapping keys, a route returning a non-dict).  Requests: router-level
infer / predict / explain, one model, a versioned model, batches of 1-4 rows, GET metadata.  Responses and exceptions compared.

    python -m tests.golden.diff_ensembles
"""
import copy
import json
import os
import random
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import api_b200 as mine  # noqa: E402
from tests.golden import api_reference as ref  # noqa: E402
from tests.scenarios import _first_line, _resp  # noqa: E402


def namespace(api):
    class Table(api.V2ModelServer):
        """predict answers row i with table[i % len(table)]; `boom` raises instead"""

        def load(self):
            pass

        def predict(self, request):
            if self.get_param("boom", False):
                raise RuntimeError("route failed")
            table = self.get_param("table")
            return [table[i % len(table)] for i in range(len(request["inputs"]))]

        def explain(self, request):
            return {"rows": len(request["inputs"])}

    class Part:
        def __init__(self, out=None, **kw):
            self.out = out

        def do(self, x):
            return copy.deepcopy(self.out) if self.out != "echo" else x

    return {"Table": Table, "Part": Part}


def norm(o):
    return tuple(re.sub(r"[0-9a-f]{32}", "<id>", x) if isinstance(x, str) else x for x in o)


def call(server, path, body, method):
    try:
        r = server.test(path, copy.deepcopy(body), method=method, silent=True)
        return norm(("ok", json.dumps(_resp(r), sort_keys=True, default=str)))
    except Exception as exc:  # noqa: BLE001
        return norm(("exc", type(exc).__name__, _first_line(exc)))


def value(rnd, kind):
    if kind == "int":
        return rnd.randint(0, 3)
    if kind == "float":
        return round(rnd.uniform(-2, 2), 3)
    if kind == "intfloat":
        return float(rnd.randint(0, 3))
    return rnd.choice([rnd.randint(0, 3), round(rnd.uniform(0, 3), 2), float(rnd.randint(0, 2))])


def ensemble_case(rnd):
    kind = rnd.choice(["int", "float", "intfloat", "mixed"])
    routes = {}
    for i in range(rnd.randint(1, 5)):
        key = f"m{i}" if rnd.random() < 0.8 else f"m{i}:v{rnd.randint(1, 2)}"
        routes[key] = {"table": [value(rnd, kind) for _ in range(rnd.randint(1, 3))], "boom": rnd.random() < 0.08}
    args = {"executor_type": "array"}
    if rnd.random() < 0.6:
        args["vote_type"] = rnd.choice(["classification", "regression"])
    if rnd.random() < 0.4:
        names = list(routes)
        chosen = [n for n in names if rnd.random() < 0.8] + (["ghost"] if rnd.random() < 0.2 else [])
        args["weights"] = {n: rnd.choice([0.5, 1.0, 2.0, 0.0, 0.25]) for n in chosen}
    if rnd.random() < 0.3:
        args["prediction_col_name"] = "p"
        args["format_response_with_col_name_flag"] = rnd.random() < 0.7
    return routes, args


def build_ensemble(api, routes, args):
    fn = api.new_function("f", kind="serving")
    fn.set_topology("router", "mlrun.serving.routers.VotingEnsemble", name="ens", **args)
    for key, cfg in routes.items():
        fn.add_model(key, ".", class_name="Table", **cfg)
    return fn.to_mock_server(namespace=namespace(api))


def build_parallel(api, parts, extend):
    fn = api.new_function("f", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    router = graph.to("*mlrun.serving.routers.ParallelRun", name="par", extend_event=extend, executor_type="array")
    for name, out in parts.items():
        router.add_route(name, class_name="Part", out=out)
    router.respond()
    return fn.to_mock_server(namespace=namespace(api))


def main():
    rnd = random.Random(5)
    n = 0
    for _case in range(400):
        routes, args = ensemble_case(rnd)
        servers = []
        for api in (ref, mine):
            try:
                servers.append(("ok", build_ensemble(api, routes, args)))
            except Exception as exc:  # noqa: BLE001
                servers.append(("exc", type(exc).__name__, _first_line(exc)))
        if [s[0] for s in servers] != ["ok", "ok"]:
            if servers[0][0] != servers[1][0] or servers[0][1:] != servers[1][1:]:
                print("BUILD DIFF", routes, args, servers)
                return 1
            n += 1
            continue
        first = next(iter(routes))
        base = first.split(":")[0]
        requests = [("/v2/models/infer", {"inputs": [[1]] * rnd.randint(1, 4)}, "POST"), ("/v2/models/ens/predict", {"inputs": [[1], [2]]}, "POST"),
                    ("/v2/models/explain", {"inputs": [[1]]}, "POST"), (f"/v2/models/{base}/infer", {"inputs": [[1], [2], [3]]}, "POST"),
                    ("/v2/models/", None, "GET"), ("/v2/models/infer", {"inputs": []}, "POST"),
                    (f"/v2/models/{base}/versions/v1/infer", {"inputs": [[1]]}, "POST"), ("/v2/models/infer", {"inputs": [[1], [2]]}, "POST")]
        for path, body, method in requests:
            a, b = call(servers[0][1], path, body, method), call(servers[1][1], path, body, method)
            n += 1
            if a != b:
                print("DIFF ensemble", routes, args, path, body, method)
                print("  ref :", a)
                print("  mine:", b)
                return 1
    for _case in range(150):
        parts = {}
        for i in range(rnd.randint(1, 4)):
            parts[f"p{i}"] = rnd.choice([{"a": i}, {"a": i, "b": [i]}, {f"k{i}": {"deep": i}}, "echo", 5, None, [1, 2]])
        extend = rnd.choice([True, False, None])
        servers = []
        for api in (ref, mine):
            try:
                servers.append(("ok", build_parallel(api, parts, extend)))
            except Exception as exc:  # noqa: BLE001
                servers.append(("exc", type(exc).__name__, _first_line(exc)))
        if [s[0] for s in servers] != ["ok", "ok"]:
            if servers[0][0] != servers[1][0] or servers[0][1:] != servers[1][1:]:
                print("BUILD DIFF parallel", parts, extend, servers)
                return 1
            continue
        for body in ({"x": 1}, {"a": "orig", "z": 0}, 7, None, [3]):
            a, b = call(servers[0][1], "/", body, "POST"), call(servers[1][1], "/", body, "POST")
            n += 1
            if a != b:
                print("DIFF parallel", parts, extend, body)
                print("  ref :", a)
                print("  mine:", b)
                return 1
    print("identical on", n, "requests")
    return 0


if __name__ == "__main__":
    sys.exit(main())
