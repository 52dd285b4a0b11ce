"""Differential check of graph BUILDING against the REAL reference (build container only): seeded random flows made with
`to` / `add_step(after=..., before=...)` -- branches, joins, steps after a responder, unknown `after` names, duplicate names,
cycles, several start steps, queues left out -- then `to_mock_server` (which runs check_and_process_graph,
serving/states.py:1073-1184) and one request.  What is compared: whether building raises and with which exception text, the
start steps / responder the graph ends up with, the serialised spec (`to_dict`, class paths stripped) and the response.

    python -m tests.golden.diff_graph_building
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


def namespace():
    class Inc:
        def __init__(self, k=1, **kw):
            self.k = k

        def do(self, x):
            return x + self.k if isinstance(x, (int, float)) else x

    return {"Inc": Inc}


def random_ops(rnd):
    ops, names = [], []
    for i in range(rnd.randint(1, 6)):
        name = f"s{i}" if rnd.random() < 0.92 or not names else rnd.choice(names)  # sometimes a duplicate name
        r = rnd.random()
        if r < 0.5 or not names:
            ops.append(("to", name, rnd.randint(1, 3)))
        else:
            after = rnd.choice(names + ["$prev", "$start", "ghost"]) if rnd.random() < 0.9 else None
            before = rnd.choice(names + ["ghost"]) if rnd.random() < 0.15 else None
            ops.append(("add_step", name, rnd.randint(1, 3), after, before))
        names.append(name)
        if rnd.random() < 0.25:
            ops.append(("respond",))
    return ops


def strip(d):
    if isinstance(d, dict):
        return {k: (v.rsplit(".", 1)[-1] if k == "class_name" and isinstance(v, str) else strip(v)) for k, v in d.items()}
    if isinstance(d, list):
        return [strip(v) for v in d]
    return d


def run(api, ops):
    out = {}
    try:
        fn = api.new_function("f", kind="serving")
        graph = fn.set_topology("flow", engine="sync")
        cur = graph
        for op in ops:
            if op[0] == "to":
                cur = cur.to("Inc", name=op[1], k=op[2])
            elif op[0] == "add_step":
                kw = {}
                if op[3] is not None:
                    kw["after"] = op[3]
                if op[4] is not None:
                    kw["before"] = op[4]
                cur = graph.add_step("Inc", name=op[1], k=op[2], **kw)
            else:
                cur.respond()
        out["spec"] = strip(graph.to_dict())
        server = fn.to_mock_server(namespace=namespace())
        out["responder"] = getattr(server.graph, "_responder", None) and server.graph._responder.name if hasattr(server.graph, "_responder") else None
        r = server.test(body=1, silent=True)
        out["response"] = json.dumps(_resp(r), sort_keys=True, default=str)
    except Exception as exc:  # noqa: BLE001
        out["error"] = f"{type(exc).__name__}: {_first_line(exc)}"
    out.pop("responder", None)
    return out


def main():
    rnd = random.Random(21)
    n = errors = 0
    for _ in range(1500):
        ops = random_ops(rnd)
        a, b = run(ref, copy.deepcopy(ops)), run(mine, copy.deepcopy(ops))
        n += 1
        errors += "error" in a
        if a != b:
            print("DIFF", ops)
            print("  ref :", json.dumps(a, sort_keys=True)[:900])
            print("  mine:", json.dumps(b, sort_keys=True)[:900])
            return 1
    print("identical on", n, "graphs (", errors, "of them rejected by both )")
    return 0


if __name__ == "__main__":
    sys.exit(main())
