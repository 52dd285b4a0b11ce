"""Differential check of routers INSIDE sync flows against the REAL reference (build container only): step -> router (ModelRouter
| VotingEnsemble | ParallelRun) -> step chains with random input_path / result_path on the router step, 1-3 models (one may
raise), requests with model / operation URLs, body-level `model` / `operation` overrides, nested `inputs`, GET metadata, bad
prefixes.  Responses and exceptions compared.

    python -m tests.golden.diff_flow_with_routers
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
    class Times(api.V2ModelServer):
        def load(self):
            pass

        def predict(self, request):
            if self.get_param("boom", False):
                raise RuntimeError("model failed")
            k = self.get_param("k", 1)
            return [k * (sum(r) if isinstance(r, list) else r) for r in request["inputs"]]

    class Wrap:
        def __init__(self, key="w", **kw):
            self.key = key

        def do(self, x):
            return {self.key: x} if not isinstance(x, dict) else {**x, self.key: True}

    class Part:
        def __init__(self, out=None, **kw):
            self.out = out

        def do(self, x):
            return copy.deepcopy(self.out)

    return {"Times": Times, "Wrap": Wrap, "Part": Part}


def build(api, g):
    fn = api.new_function("f", kind="serving")
    flow = fn.set_topology("flow", engine="sync")
    cur = flow
    if g["pre"]:
        cur = cur.to("Wrap", name="pre", key="pre")
    kw = {k: g[k] for k in ("input_path", "result_path") if g.get(k)}
    if g["router"] == "parallel":
        cur = cur.to("*mlrun.serving.routers.ParallelRun", name="r", executor_type="array", extend_event=g["extend"], **kw)
        for i, out in enumerate(g["parts"]):
            cur.add_route(f"p{i}", class_name="Part", out=out)
    else:
        cls = "*mlrun.serving.ModelRouter" if g["router"] == "model" else "*mlrun.serving.routers.VotingEnsemble"
        extra = {} if g["router"] == "model" else {"executor_type": "array", "vote_type": "regression"}
        cur = cur.to(cls, name="r", **extra, **kw)
        for i, m in enumerate(g["models"]):
            cur.add_route(f"m{i}", class_name="Times", model_path=".", **m)
    if g["post"]:
        cur = cur.to("Wrap", name="post", key="post")
    cur.respond()
    return fn.to_mock_server(namespace=namespace(api))


def main():
    rnd = random.Random(37)
    n = 0
    paths = ["/", "/v2/models/m0/infer", "/v2/models/m1/predict", "/v2/models/infer", "/v2/models/r/infer", "/v2/models/m9/infer", "/v2/models/",
             "/v2/models/m0", "/other/m0/infer", "/v2/models/m0/explain", "/v2/models/m0/ready"]
    for case in range(350):
        g = {"pre": rnd.random() < 0.3, "post": rnd.random() < 0.4, "router": rnd.choice(["model", "ensemble", "ensemble", "parallel"]),
             "input_path": rnd.choice([None, None, "x", "x.y"]), "result_path": rnd.choice([None, None, "res", "a.b"]),
             "models": [{"k": rnd.randint(1, 4), **({"boom": True} if rnd.random() < 0.1 else {})} for _ in range(rnd.randint(1, 3))],
             "parts": [rnd.choice([{"a": i}, {"b": [i]}, 5, None]) for i in range(rnd.randint(1, 3))], "extend": rnd.choice([True, False, None])}
        servers = []
        for api in (ref, mine):
            try:
                servers.append(("ok", build(api, g)))
            except Exception as exc:  # noqa: BLE001
                servers.append(("exc", type(exc).__name__, _first_line(exc)))
        if [s[0] for s in servers] != ["ok", "ok"]:
            assert servers[0][0] == servers[1][0] and servers[0][1:] == servers[1][1:], (g, servers)
            n += 1
            continue
        for _ in range(8):
            inner = rnd.choice([{"inputs": [[1, 2], [3]]}, {"inputs": [5]}, {"inputs": []}, {"model": "m1", "inputs": [2]}, {"operation": "predict", "inputs": [1]},
                                {"inputs": "bad"}, {}, 7])
            wrap = rnd.choice(["plain", "x", "x.y"])
            body = inner if wrap == "plain" else ({"x": inner} if wrap == "x" else {"x": {"y": inner}, "keep": 1})
            path, method = rnd.choice(paths), rnd.choice(["POST", "POST", "GET"])
            out = []
            for _state, server in servers:
                try:
                    r = server.test(path, copy.deepcopy(body), method=method, silent=True)
                    out.append(("ok", json.dumps(_resp(r), sort_keys=True, default=str)))
                except Exception as exc:  # noqa: BLE001
                    out.append(("exc", type(exc).__name__, _first_line(exc)))
            out = [tuple(re.sub(r"<[\w.]*MockEvent object at 0x[0-9a-f]+>", "<MockEvent>", re.sub(r"[0-9a-f]{32}", "<id>", x)) if isinstance(x, str) else x
                         for x in o) for o in out]
            n += 1
            if out[0] != out[1]:
                print("DIFF", json.dumps(g), path, body, method)
                print("  ref :", out[0][:800])
                print("  mine:", out[1][:800])
                return 1
    print("identical on", n, "requests")
    return 0


if __name__ == "__main__":
    sys.exit(main())
