"""Differential check of the product's flow engine (sync) against the REAL reference (needs /root/reference: build container
only): seeded random chains of 1-5 steps -- classes with `do`, classes with `do_event`, plain functions, handlers by name --
with random `input_path` / `result_path` / `full_event`, a step that raises (with and without a step-level or graph-level error
handler), a random `.respond()` position, and a handful of bodies (nested dicts, a missing path, a scalar): `server.test`
responses and exceptions (type + first line) compared.  Last run: identical on every graph.

    python -m tests.golden.diff_flow_graphs
"""
import copy
import json
import random
import re
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from tests import api_b200 as mine  # noqa: E402
from tests.golden import api_reference as ref  # noqa: E402
from tests.scenarios import _first_line, _resp  # noqa: E402


def namespace(api):
    class Scale:
        def __init__(self, k=2, **kw):
            self.k = k

        def do(self, x):
            if isinstance(x, dict):
                return {key: (v * self.k if isinstance(v, (int, float)) else v) for key, v in x.items()}
            return x * self.k

    class Tag:
        def __init__(self, context=None, name=None, tag="t", **kw):
            self.tag, self.name = tag, name

        def do_event(self, event):
            body = event.body
            event.body = {"tagged": body, "by": self.name, "tag": self.tag, "path": getattr(event, "path", None)} if not isinstance(body, dict) else {**body, "tag": self.tag}
            return event

    class Boom:
        def __init__(self, **kw):
            pass

        def do(self, x):
            raise ValueError(f"boom on {type(x).__name__}")

    class Handled:
        def __init__(self, **kw):
            pass

        def do_event(self, event):
            event.body = {"handled": str(getattr(event, "error", None))[:60], "origin": getattr(event, "origin_state", None)}
            return event

    def plus_one(x):
        if isinstance(x, dict):
            return {**x, "n": x.get("n", 0) + 1}
        return x + 1 if isinstance(x, (int, float)) else x

    class Multi:
        def __init__(self, **kw):
            pass

        def first(self, x):
            return {"first": x}

        def second(self, x):
            return [x, x]

    return {"Scale": Scale, "Tag": Tag, "Boom": Boom, "Handled": Handled, "plus_one": plus_one, "Multi": Multi}


def random_graph(rnd):
    steps = []
    for i in range(rnd.randint(1, 5)):
        kind = rnd.choice(["Scale", "Tag", "plus_one", "Multi", "Boom"] if rnd.random() < 0.25 else ["Scale", "Tag", "plus_one", "Multi"])
        spec = {"name": f"s{i}", "kind": kind, "args": {}}
        if kind == "Scale":
            spec["args"] = {"k": rnd.choice([2, 3, -1])}
        if kind == "Tag":
            spec["args"] = {"tag": rnd.choice(["a", "b"])}
        if kind == "Multi":
            spec["handler"] = rnd.choice(["first", "second"])
        if rnd.random() < 0.35:
            spec["input_path"] = rnd.choice(["x", "x.y", "q", "x.missing"])
        if rnd.random() < 0.35:
            spec["result_path"] = rnd.choice(["out", "x.res", "deep.a.b"])
        if rnd.random() < 0.15 and kind in ("Tag",):
            spec["full_event"] = True
        spec["on_error"] = kind == "Boom" and rnd.random() < 0.5
        steps.append(spec)
    return {"steps": steps, "respond_at": rnd.randrange(len(steps)) if rnd.random() < 0.8 else None,
            "graph_error_handler": rnd.random() < 0.2}


def build(api, g):
    ns = namespace(api)
    fn = api.new_function("f", kind="serving")
    flow = fn.set_topology("flow", engine="sync")
    cur = flow
    for i, sp in enumerate(g["steps"]):
        kw = dict(name=sp["name"], **{k: sp[k] for k in ("input_path", "result_path", "full_event") if k in sp}, **sp["args"])
        if sp["kind"] == "plus_one":
            cur = cur.to(name=kw.pop("name"), handler="plus_one", **{k: v for k, v in kw.items()})
        elif "handler" in sp:
            cur = cur.to(sp["kind"], handler=sp["handler"], **kw)
        else:
            cur = cur.to(sp["kind"], **kw)
        if sp["on_error"]:
            cur.error_handler(name=f"catch{i}", class_name="Handled")
        if g["respond_at"] == i:
            cur.respond()
    if g["graph_error_handler"]:
        flow.error_handler(name="catch_all", class_name="Handled")
    return fn.to_mock_server(namespace=ns)


BODIES = [{"x": {"y": 3}, "q": 2, "n": 1}, {"x": 5}, 7, {"q": [1, 2]}, "text", None, {"x": {"y": {"z": 1}}}]


def main():
    rnd = random.Random(11)
    n = 0
    for gi in range(700):
        g = random_graph(rnd)
        servers = []
        for api in (ref, mine):
            try:
                servers.append(("ok", build(api, g)))
            except Exception as exc:  # noqa: BLE001
                servers.append(("exc", type(exc).__name__, _first_line(exc)))
        if servers[0][0] != servers[1][0] or (servers[0][0] == "exc" and servers[0] != servers[1]):
            print("BUILD DIFF", json.dumps(g), servers)
            return 1
        if servers[0][0] == "exc":
            n += 1
            continue
        for body in BODIES:
            out = []
            for _state, server in servers:
                try:
                    r = server.test(body=copy.deepcopy(body), silent=True)
                    out.append(("ok", json.dumps(_resp(r), sort_keys=True, default=str)))
                except Exception as exc:  # noqa: BLE001
                    out.append(("exc", type(exc).__name__, _first_line(exc)))
            out = [tuple(re.sub(r"<[\w.]*MockEvent object at 0x[0-9a-f]+>", "<MockEvent>", re.sub(r"[0-9a-f]{32}", "<id>", x)) if isinstance(x, str) else x
                         for x in o) for o in out]
            n += 1
            if out[0] != out[1]:
                print("DIFF", json.dumps(g), "body", body)
                print("  ref :", out[0])
                print("  mine:", out[1])
                return 1
    print("identical on", n, "graph x body cases")
    return 0


if __name__ == "__main__":
    sys.exit(main())
