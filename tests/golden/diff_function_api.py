"""Differential check of the function-level builder API against the REAL reference (runtimes/nuclio/serving.py:245-445; build
container only): random `set_topology` / `add_model` / `set_tracking` / `add_child_function`-free call sequences -- router and
flow topologies, explicit router steps, versions in keys, class given as name / object / missing, `handler`,
`router_step`, `child_function`, keyword class args, exist_ok, repeated keys -- comparing what is refused (exception text) and the
serialised graph that results.

    python -m tests.golden.diff_function_api
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import api_b200 as mine  # noqa: E402
from tests.golden import api_reference as ref  # noqa: E402
from tests.golden.diff_serialisation import strip  # noqa: E402
from tests.scenarios import _first_line  # noqa: E402


def run(api, plan):
    try:
        fn = api.new_function("f", kind="serving")
        for call in plan:
            name, args, kw = call
            kw = dict(kw)
            if kw.get("class_name") == "$object":
                class Inline(api.V2ModelServer):
                    def load(self):
                        pass

                    def predict(self, request):
                        return request["inputs"]

                kw["class_name"] = Inline(name=args[0] if args else "inline", model_path=".")
            getattr(fn, name)(*args, **kw)
        graph = fn.spec.graph
        out = {"graph": strip(graph.to_dict()) if graph is not None else None}
        params = dict(getattr(fn.spec, "parameters", {}) or {})
        out["parameters"] = {k: params[k] for k in sorted(params)}
        out["track_models"] = getattr(fn.spec, "track_models", None)
        return json.dumps(out, sort_keys=True, default=str)
    except Exception as exc:  # noqa: BLE001
        return json.dumps(("exc", type(exc).__name__, _first_line(exc)))


def random_plan(rnd):
    plan = []
    topo = rnd.choice(["router", "router", "router", "flow", None])
    if topo == "router":
        kw = {}
        if rnd.random() < 0.3:
            kw["class_name"] = rnd.choice(["mlrun.serving.routers.VotingEnsemble", "mlrun.serving.ModelRouter"])
        if rnd.random() < 0.2:
            kw["exist_ok"] = True
        plan.append(("set_topology", ("router",), kw))
    elif topo == "flow":
        plan.append(("set_topology", ("flow",), {"engine": rnd.choice(["sync", "async"])}))
    if rnd.random() < 0.1:
        plan.append(("set_topology", (rnd.choice(["router", "flow", "bogus"]),), {"exist_ok": rnd.random() < 0.5}))
    for i in range(rnd.randint(0, 4)):
        key = rnd.choice([f"m{i}", f"m{i}:v1", "m0", ""])
        kw = {}
        r = rnd.random()
        if r < 0.5:
            kw["class_name"] = "SomeModel"
        elif r < 0.65:
            kw["class_name"] = "$object"
        if rnd.random() < 0.93:
            kw["model_path"] = rnd.choice([".", ".", "store://models/x", ""])
        # (`model_url` -> a `$remote` step is networking: out of scope, the product refuses it loudly -- DESIGN.md section 8)
        if rnd.random() < 0.12:
            kw["handler"] = "do_it"
        if rnd.random() < 0.15:
            kw["router_step"] = rnd.choice(["r1", "missing"])
        if rnd.random() < 0.1:
            kw["child_function"] = "child"
        if rnd.random() < 0.4:
            kw["multiplier"] = i
        plan.append(("add_model", (key,), kw))
    if rnd.random() < 0.25:
        kw = {}
        if rnd.random() < 0.5:
            kw["batch"] = rnd.randint(1, 5)
        if rnd.random() < 0.5:
            kw["sample"] = rnd.choice([1, 3])
        if rnd.random() < 0.05:
            kw["sampling_percentage"] = 50  # not a parameter of this release: a TypeError in both
        plan.append(("set_tracking", (rnd.choice(["dummy://", None, "v3io:///x"]),), kw))
    rnd.shuffle(plan) if rnd.random() < 0.1 else None
    return plan


def main():
    rnd = random.Random(61)
    n = errors = 0
    for _ in range(2500):
        plan = random_plan(rnd)
        a, b = run(ref, plan), run(mine, plan)
        n += 1
        errors += a.startswith('["exc"')
        if a != b:
            print("DIFF", plan)
            print("  ref :", a[:1000])
            print("  mine:", b[:1000])
            return 1
    print("identical on", n, "call sequences (", errors, "refused by both )")
    return 0


if __name__ == "__main__":
    sys.exit(main())
