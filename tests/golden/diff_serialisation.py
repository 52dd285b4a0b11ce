"""Differential check of graph SERIALISATION against the REAL reference (build container only): random function graphs --
router topologies with models (class args, versions, handlers), flows holding routers, task steps with input / result paths,
`full_event`, responders, step- and graph-level error handlers, `add_step(after=...)` -- serialised with `graph.to_dict()`
(serving/utils.py:46-109, serving/states.py:102-150), rebuilt with `from_dict` where the class offers it and serialised again.
Class paths are reduced to the class name (mlrun.* vs mlrun_b200.* differ by design).

    python -m tests.golden.diff_serialisation
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import api_b200 as mine  # noqa: E402
from tests.golden import api_reference as ref  # noqa: E402
from tests.scenarios import _first_line  # noqa: E402


def strip(d):
    if isinstance(d, dict):
        return {k: (v.rsplit(".", 1)[-1] if k == "class_name" and isinstance(v, str) else strip(v)) for k, v in d.items()}
    if isinstance(d, list):
        return [strip(v) for v in d]
    return d


def build(api, plan):
    fn = api.new_function("f", kind="serving")
    if plan["topology"] == "router":
        kw = dict(plan["router_args"])
        graph = fn.set_topology("router", plan["router_class"], **kw) if plan["router_class"] else fn.set_topology("router", **kw)
        for key, args in plan["models"]:
            fn.add_model(key, ".", class_name="SomeModel", **args)
        return fn, graph
    graph = fn.set_topology("flow", engine=plan["engine"])
    cur = graph
    for st in plan["steps"]:
        kw = {k: st[k] for k in ("input_path", "result_path", "full_event", "handler") if k in st}
        if st["how"] == "to":
            cur = cur.to(st["cls"], name=st["name"], **kw, **st["args"])
        elif st["how"] == "router":
            cur = cur.to("*" + st["cls"], name=st["name"], **st["args"])
            for key, args in st["routes"]:
                cur.add_route(key, class_name="SomeModel", **args)
        else:
            cur = graph.add_step(st["cls"], name=st["name"], after=st["after"], **kw, **st["args"])
        if st.get("respond"):
            cur.respond()
        if st.get("on_error"):
            cur.error_handler(name=st["name"] + "_err", class_name="Catcher", k=1)
    if plan.get("graph_handler"):
        graph.error_handler(name="catch_all", class_name="Catcher")
    return fn, graph


def random_plan(rnd):
    models = lambda: [(f"m{i}" + (f":v{i}" if rnd.random() < 0.3 else ""), {k: v for k, v in (("multiplier", i), ("handler", "run") if rnd.random() < 0.1 else ("x", None)) if v is not None})  # noqa: E731
                      for i in range(rnd.randint(0, 3))]
    if rnd.random() < 0.4:
        cls = rnd.choice([None, "mlrun.serving.routers.VotingEnsemble", "mlrun.serving.routers.ParallelRun"])
        args = {}
        if cls and "Voting" in cls and rnd.random() < 0.5:
            args = {"vote_type": "regression", "name": "ens"}
        return {"topology": "router", "router_class": cls, "router_args": args, "models": models()}
    steps, names = [], []
    for i in range(rnd.randint(1, 5)):
        st = {"name": f"s{i}", "cls": "Worker", "args": {"k": rnd.randint(1, 3)} if rnd.random() < 0.7 else {}, "how": "to"}
        if rnd.random() < 0.2:
            st.update(how="router", cls=rnd.choice(["mlrun.serving.ModelRouter", "mlrun.serving.routers.VotingEnsemble"]), args={}, routes=models())
        elif names and rnd.random() < 0.25:
            st.update(how="add_step", after=rnd.choice(names))
        if st["how"] != "router":
            if rnd.random() < 0.3:
                st["input_path"] = "a.b"
            if rnd.random() < 0.3:
                st["result_path"] = "out"
            if rnd.random() < 0.15:
                st["full_event"] = True
            if rnd.random() < 0.15:
                st["handler"] = "other"
            if rnd.random() < 0.15:
                st["on_error"] = True
        if rnd.random() < 0.2:
            st["respond"] = True
        steps.append(st)
        names.append(st["name"])
    return {"topology": "flow", "engine": rnd.choice(["sync", "async", None]), "steps": steps, "graph_handler": rnd.random() < 0.2}


def main():
    rnd = random.Random(4)
    n = 0
    for _ in range(1500):
        plan = random_plan(rnd)
        out = []
        for api in (ref, mine):
            try:
                fn, graph = build(api, plan)
                first = strip(graph.to_dict())
                again = None
                if hasattr(type(graph), "from_dict"):
                    again = strip(type(graph).from_dict(graph.to_dict()).to_dict())
                spec = fn.spec.graph.to_dict() if hasattr(fn, "spec") and getattr(fn.spec, "graph", None) is not None else None
                out.append(json.dumps({"first": first, "again": again, "spec": strip(spec)}, sort_keys=True, default=str))
            except Exception as exc:  # noqa: BLE001
                out.append(json.dumps(("exc", type(exc).__name__, _first_line(exc))))
        n += 1
        if out[0] != out[1]:
            print("DIFF", json.dumps(plan, default=str))
            print("  ref :", out[0][:1200])
            print("  mine:", out[1][:1200])
            return 1
    print("identical on", n, "graphs")
    return 0


if __name__ == "__main__":
    sys.exit(main())
