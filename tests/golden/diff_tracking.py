"""Differential check of model tracking (`_ModelLogPusher`, serving/v2_serving.py:429-504) against the REAL reference (build
container only): random `log_stream_sample` / `log_stream_batch` parameters, random request sequences (good bodies of 1-3
rows, a body that makes the model raise, explain / ready / predict operations, ids given and generated), on a router and on a
voting ensemble.  What is compared: the records pushed to the output stream -- count, order, keys, models, operations,
requests / responses (ids and timestamps by type), batching of the kept requests.

    python -m tests.golden.diff_tracking
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
from tests.scenarios import _clean, make_namespace  # noqa: E402


def shape(rec):
    keep = {k: rec[k] for k in ("model", "op", "class", "function_uri", "version", "labels", "endpoint_id") if k in rec}
    for key in ("request", "resp", "requests", "error"):
        if key in rec:
            keep[key] = _clean(rec[key])
    if "values" in rec:
        keep["values"] = [[_clean(v[0]), v[1], _clean(v[2]), type(v[3]).__name__, type(v[4]).__name__, v[5]] for v in rec["values"]]
    for key in ("when", "microsec", "metrics"):
        if key in rec:
            keep[f"{key}_kind"] = type(rec[key]).__name__
    if isinstance(rec.get("microsec"), list):
        keep["n_timings"] = len(rec["microsec"])
    keep["keys"] = sorted(rec)
    return keep


def main():
    rnd = random.Random(29)
    n = 0
    for case in range(160):
        params = {}
        if rnd.random() < 0.6:
            params["log_stream_sample"] = rnd.randint(1, 4)
        if rnd.random() < 0.6:
            params["log_stream_batch"] = rnd.randint(1, 4)
        topo = rnd.choice(["router", "ensemble"])
        calls = []
        for i in range(rnd.randint(3, 12)):
            op = rnd.choice(["infer", "infer", "infer", "predict", "explain", "ready"])
            body = rnd.choice([{"inputs": [i]}, {"inputs": [i, i + 1, i + 2]}, {"inputs": "not-a-list"}, {"inputs": [[i]]}, {"id": f"given{i}", "inputs": [i]}])
            target = rnd.choice(["my", "other"]) if topo == "router" else rnd.choice(["my", "other", None])
            path = f"/v2/models/{target}/{op}" if target else f"/v2/models/{op}"
            calls.append((path, body, f"e{i}" if rnd.random() < 0.7 else None))
        out = []
        for api in (ref, mine):
            ns = make_namespace(api)
            fn = api.new_function("trk", kind="serving")
            if topo == "router":
                fn.set_topology("router")
            else:
                fn.set_topology("router", "mlrun.serving.routers.VotingEnsemble", name="ens", executor_type="array", vote_type="regression")
            fn.add_model("my", ".", class_name=ns["ModelTestingClass"](multiplier=10))
            fn.add_model("other", ".", class_name=ns["ModelTestingClass"](multiplier=3))
            fn.set_tracking("dummy://")
            fn.spec.parameters.update(params)
            server = fn.to_mock_server(namespace=ns)
            for path, body, eid in calls:
                kw = {"event_id": eid} if eid else {}
                try:
                    server.test(path, copy.deepcopy(body), silent=True, **kw)
                except Exception:  # noqa: BLE001
                    pass
            recs = [shape(r) for r in server.context.stream.output_stream.event_list]
            out.append(re.sub(r"[0-9a-f]{32}", "<id>", json.dumps(recs, sort_keys=True, default=str)))
        n += 1
        if out[0] != out[1]:
            print("DIFF", case, topo, params, calls)
            a, b = json.loads(out[0]), json.loads(out[1])
            print("  records ref / mine:", len(a), len(b))
            for i, (x, y) in enumerate(zip(a, b)):
                if x != y:
                    print("  first differing record", i)
                    print("  ref :", json.dumps(x)[:900])
                    print("  mine:", json.dumps(y)[:900])
                    break
            return 1
    print("identical on", n, "tracked request sequences")
    return 0


if __name__ == "__main__":
    sys.exit(main())
