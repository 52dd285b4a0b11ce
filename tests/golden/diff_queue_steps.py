"""Differential check of QUEUE steps in sync flows against the REAL reference (serving/states.py:749-890; build container only):
random chains with `>>` / `$queue` steps (no path, a `dummy://` stream path, shards / retention arguments), before and after
ordinary steps, with responders in various places: what is refused at build / server start, the serialised graph, the
response of a request and what lands on the queue's stream.

    python -m tests.golden.diff_queue_steps
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
from tests.golden.diff_serialisation import strip  # noqa: E402
from tests.scenarios import _clean, _first_line, _resp  # noqa: E402


def namespace():
    class Inc:
        def __init__(self, k=1, **kw):
            self.k = k

        def do(self, x):
            return x + self.k if isinstance(x, (int, float)) else {**x, "inc": self.k} if isinstance(x, dict) else x

    return {"Inc": Inc}


def run(api, plan, engine):
    out = {}
    try:
        fn = api.new_function("f", kind="serving")
        graph = fn.set_topology("flow", engine=engine)
        cur = graph
        queues = []
        for op in plan:
            if op[0] == "step":
                cur = cur.to("Inc", name=op[1], k=op[2])
            elif op[0] == "queue":
                kw = dict(op[3])
                cur = cur.to(op[2], name=op[1], **kw)
                queues.append(op[1])
            else:
                cur.respond()
        out["spec"] = strip(graph.to_dict())
        if engine == "sync":
            server = fn.to_mock_server(namespace=namespace())
            for body in (1, {"a": 1}, None):
                try:
                    r = server.test(body=copy.deepcopy(body), silent=True)
                    out[f"resp_{body!r}"] = json.dumps(_resp(r), sort_keys=True, default=str)
                except Exception as exc:  # noqa: BLE001
                    out[f"resp_{body!r}"] = f"{type(exc).__name__}: {_first_line(exc)}"
            for q in queues:
                step = server.graph[q]
                stream = getattr(step, "_stream", None)
                out[f"queue_{q}"] = _clean(list(getattr(stream, "event_list", []))) if stream is not None else None
    except Exception as exc:  # noqa: BLE001
        out["error"] = f"{type(exc).__name__}: {_first_line(exc)}"
    return re.sub(r"[0-9a-f]{32}", "<id>", json.dumps(out, sort_keys=True, default=str))


def random_plan(rnd):
    plan = []
    for i in range(rnd.randint(1, 5)):
        if rnd.random() < 0.35:
            kw = {}
            if rnd.random() < 0.6:
                kw["path"] = rnd.choice(["dummy://", "dummy://q2", ""])
            if rnd.random() < 0.2:
                kw["shards"] = 2
            if rnd.random() < 0.1:
                kw["retention_in_hours"] = 4
            plan.append(("queue", f"q{i}", rnd.choice([">>", "$queue"]), kw))
        else:
            plan.append(("step", f"s{i}", rnd.randint(1, 3)))
        if rnd.random() < 0.25:
            plan.append(("respond",))
    return plan


def main():
    rnd = random.Random(71)
    n = errors = 0
    for _ in range(1200):
        plan = random_plan(rnd)
        engine = rnd.choice(["sync", "sync", "async"])
        a, b = run(ref, plan, engine), run(mine, plan, engine)
        n += 1
        errors += '"error"' in a
        if a != b:
            print("DIFF", engine, plan)
            print("  ref :", a[:1200])
            print("  mine:", b[:1200])
            return 1
    print("identical on", n, "flows (", errors, "refused by both )")
    return 0


if __name__ == "__main__":
    sys.exit(main())
