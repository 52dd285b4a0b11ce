"""Differential check of the product's host layer against the REAL reference (needs /root/reference: build container only,
like gen_golden.py): a ModelRouter and a VotingEnsemble (array executor) over three V2 model servers, `server.test` with 28
URL paths x 21 bodies (dicts, JSON text, bytes, lists, malformed JSON, odd `inputs`) x GET / POST / PUT, a seeded 60 % sample:
responses (event ids normalised) and exceptions (type + first line) compared.  Last run: identical on 2 147 requests.

With the default THREAD executor the reference's mean vote depends on the order in which the routes' futures complete
(`as_completed`, serving/routers.py:414-455, 789-810): three routes predicting 7 / 14 / 21 with weights 1/3 gave 14.0 in one run
and 13.999999999999998 in 300 others -- in the reference and in the product alike.  The fused device vote adds in model order
(the array executor's order), so it is deterministic.

    python -m tests.golden.diff_server_requests
"""
import sys, random, json, math
sys.path.insert(0, '/root/repo')
import numpy as np
from tests.golden import api_reference as ref
from tests import api_b200 as mine
from tests.scenarios import _resp, _first_line

def build(api, topology):
    class M(api.V2ModelServer):
        def load(self): pass
        def predict(self, request):
            return [float(np.sum(r)) * self.get_param("k", 1) if isinstance(r, list) else r for r in request["inputs"]]
        def explain(self, request):
            return {"explained": len(request["inputs"])}
        def op_custom(self, event):
            return {"custom": event.path}
    fn = api.new_function("f", kind="serving")
    if topology == "router":
        fn.set_topology("router")
    else:
        fn.set_topology("router", "mlrun.serving.routers.VotingEnsemble", name="ens", vote_type="regression", executor_type="array")
    fn.add_model("m1", ".", class_name="M", k=1)
    fn.add_model("m2", ".", class_name="M", k=2)
    fn.add_model("m3:v2", ".", class_name="M", k=3)
    return fn.to_mock_server(namespace={"M": M})

rnd = random.Random(3)
paths = ["/", "", "/v2/models", "/v2/models/", "/v2/models/m1", "/v2/models/m1/infer", "/v2/models/m1/predict", "/v2/models/m2/explain",
         "/v2/models/m1/ready", "/v2/models/m1/custom", "/v2/models/m1/metrics", "/v2/models/m9/infer", "/v2/models/m3/versions/v2/infer",
         "/v2/models/m3/versions/v9/infer", "/v2/models/infer", "/v2/models/predict", "/v2/models/explain", "/v2/models/ens/infer",
         "/v2/models/ens", "/v2/health", "/v1/models/m1/infer", "/v2/models/m1/infer/", "/v2/models/m1/nope", "/v2/models/m1/infer_dict",
         "/v2/models/m2/predict_dict", "/bad", "/v2/models/m1/versions", "/v2/models//infer"]
bodies = [None, {}, {"inputs": [[1, 2], [3, 4]]}, {"inputs": [5]}, {"inputs": []}, {"inputs": 5}, {"inputs": "x"}, {"model": "m2", "inputs": [[1]]},
          {"operation": "explain", "inputs": [[1, 2]]}, {"operation": "predict", "model": "m1", "inputs": [1.5]}, {"id": "abc", "inputs": [[1]]},
          '{"inputs": [[1, 2]]}', '{"inputs": [[1, 2]', b'{"inputs": [3]}', "plain text", [1, 2], {"inputs": [{"a": 1}]}, {"inputs": [[1, None]]},
          {"model": "nope", "inputs": [1]}, {"operation": "ready"}, {"operation": None, "inputs": [1]}]
methods = ["POST", "GET", "PUT"]
n = 0
for topology in ("router", "ensemble"):
    sa, sb = build(ref, topology), build(mine, topology)
    for p in paths:
        for b in bodies:
            for m in methods:
                if rnd.random() > 0.6: continue
                out = []
                for s in (sa, sb):
                    try:
                        import copy
                        r = s.test(p, copy.deepcopy(b), method=m, silent=True)
                        out.append(("ok", json.dumps(_resp(r), sort_keys=True, default=str)))
                    except Exception as e:
                        out.append(("exc", type(e).__name__, _first_line(e)))
                n += 1
                import re
                out = [tuple(re.sub(r"[0-9a-f]{32}", "<id>", x) if isinstance(x, str) else x for x in o) for o in out]
                a, c = out
                # drop volatile ids/timestamps already removed by _resp/_clean
                if a != c:
                    print("DIFF", topology, repr(p), b, m); print("  ref :", a); print("  mine:", c)
                    sys.exit(1)
print("identical on", n, "requests")
