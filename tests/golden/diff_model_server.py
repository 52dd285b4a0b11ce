"""Differential check of the V2 model-server protocol against the REAL V2ModelServer (serving/v2_serving.py:32-426; build
container only): a model with every optional hook (preprocess / postprocess / validate / explain / logged_results / op_*),
ready states (sync and async load, a load that fails), predict returning lists / dicts / scalars / numpy, every operation
(infer, predict, explain, ready, metrics, infer_dict, predict_dict, custom, unknown), bodies with and without `inputs`, ids,
extra keys; stand-alone in a flow and behind a router.  Responses and exceptions compared.

    python -m tests.golden.diff_model_server
"""
import copy
import json
import os
import random
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from tests import api_b200 as mine  # noqa: E402
from tests.golden import api_reference as ref  # noqa: E402
from tests.scenarios import _first_line, _resp  # noqa: E402


def namespace(api):
    class Hooked(api.V2ModelServer):
        def load(self):
            if self.get_param("fail_load", False):
                raise RuntimeError("cannot load")
            self.loaded_with = self.get_param("k", 1)

        def preprocess(self, request, operation):
            if self.get_param("pre", False) and isinstance(request, dict):
                request = {**request, "pre": operation}
            return request

        def postprocess(self, request):
            if self.get_param("post", False):
                request["post"] = True
            return request

        def validate(self, request, operation):
            if self.get_param("strict", False):
                return super().validate(request, operation)
            return request

        def predict(self, request):
            kind = self.get_param("ret", "list")
            x = request.get("inputs", [])
            if kind == "list":
                return [self.loaded_with * (sum(r) if isinstance(r, list) else r) for r in x]
            if kind == "dict":
                return {"n": len(x), "k": self.loaded_with}
            if kind == "scalar":
                return 3.5
            if kind == "numpy":
                return (np.asarray(x, dtype=np.float64) * self.loaded_with).tolist()
            raise ValueError("predict failed")

        def explain(self, request):
            return [len(request.get("inputs", []))]

        def logged_results(self, request, response, op):
            return None, None

        def op_sum(self, event):
            return {"sum": sum(event.body.get("inputs", [])) if isinstance(event.body, dict) else None}

    return {"Hooked": Hooked}


def call(server, path, body, method):
    try:
        r = server.test(path, copy.deepcopy(body), method=method, silent=True)
        return re.sub(r"[0-9a-f]{32}", "<id>", json.dumps(("ok", _resp(r)), sort_keys=True, default=str))
    except Exception as exc:  # noqa: BLE001
        return re.sub(r"[0-9a-f]{32}", "<id>", json.dumps(("exc", type(exc).__name__, _first_line(exc))))


def main():
    rnd = random.Random(17)
    n = 0
    ops = ["infer", "predict", "explain", "ready", "metrics", "infer_dict", "predict_dict", "sum", "nope", "", "explain_dict"]
    bodies = [{"inputs": [[1, 2], [3, 4]]}, {"inputs": [5, 6]}, {"inputs": []}, {}, None, {"inputs": 5}, {"id": "given", "inputs": [1]},
              {"inputs": [[1]], "extra": {"a": 1}}, {"inputs": {"a": [1, 2]}}, '{"inputs": [2]}', {"instances": [1]}, {"inputs": [[1, "x"]]}]
    for _case in range(220):
        args = {"k": rnd.choice([1, 2]), "ret": rnd.choice(["list", "list", "dict", "scalar", "numpy", "raise"]),
                "pre": rnd.random() < 0.3, "post": rnd.random() < 0.3, "strict": rnd.random() < 0.7}
        if rnd.random() < 0.08:
            args["fail_load"] = True
        if rnd.random() < 0.2:
            args["load_mode"] = rnd.choice(["sync", "async"])
        topo = rnd.choice(["router", "flow"])
        servers = []
        for api in (ref, mine):
            try:
                fn = api.new_function("f", kind="serving")
                if topo == "router":
                    fn.set_topology("router")
                    fn.add_model("m", ".", class_name="Hooked", **args)
                else:
                    flow = fn.set_topology("flow", engine="sync")
                    flow.to("Hooked", name="m", model_path=".", **args).respond()
                servers.append(("ok", fn.to_mock_server(namespace=namespace(api))))
            except Exception as exc:  # noqa: BLE001
                servers.append(("exc", type(exc).__name__, _first_line(exc)))
        if [s[0] for s in servers] != ["ok", "ok"]:
            if servers[0] != servers[1] and (servers[0][0] != servers[1][0] or servers[0][1:] != servers[1][1:]):
                print("BUILD DIFF", topo, args, servers)
                return 1
            n += 1
            continue
        if args.get("load_mode") == "async":
            import time

            time.sleep(0.05)  # the loader thread (v2_serving.py:137-140)
        for _ in range(14):
            op = rnd.choice(ops)
            path = (f"/v2/models/m/{op}" if op else "/v2/models/m") if topo == "router" else (f"/{op}" if rnd.random() < 0.5 else f"/v2/models/m/{op}")
            body, method = rnd.choice(bodies), rnd.choice(["POST", "POST", "GET"])
            a, b = call(servers[0][1], path, body, method), call(servers[1][1], path, body, method)
            n += 1
            if a != b:
                print("DIFF", topo, args, path, body, method)
                print("  ref :", a[:700])
                print("  mine:", b[:700])
                return 1
    print("identical on", n, "requests")
    return 0


if __name__ == "__main__":
    sys.exit(main())
