"""Evidence for SURVEY 8(a) a5 (the async engine, whose third-party executor `storey` is neither under /root/reference nor
installable -- parity unpinned by construction): for LINEAR flows without raising steps the reference's async engine and its
sync engine answer alike (its own tests run the same graphs under both, tests/serving/test_flow.py).  This script runs the
product's ASYNC engine (its executor restated from the reference's call sites, over the coalescing ring) on seeded random linear
flows -- `do` / `do_event` / function / named-handler steps, random input_path / result_path / full_event, the last step the
responder -- beside the REAL reference's SYNC engine: identical on every graph x body case.  Flows with raising steps are left
out: there the two engines of the reference differ by design (storey routes the event to a recovery step; the failing step's
result_path merge never happens), and only the reference tests' literals pin the async side.

    python -m tests.golden.diff_async_linear
"""
import copy
import json
import os
import random
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.golden import diff_flow_graphs as d
from tests import api_b200 as mine
from tests.golden import api_reference as ref
from tests.scenarios import _resp, _first_line

def build(api, g, engine):
    ns = d.namespace(api)
    fn = api.new_function("f", kind="serving")
    flow = fn.set_topology("flow", engine=engine)
    cur = flow
    for i, sp in enumerate(g["steps"]):
        kw = dict(name=sp["name"], **{k: sp[k] for k in ("input_path", "result_path", "full_event") if k in sp}, **sp["args"])
        if sp["kind"] == "plus_one":
            cur = cur.to(name=kw.pop("name"), handler="plus_one", **kw)
        elif "handler" in sp:
            cur = cur.to(sp["kind"], handler=sp["handler"], **kw)
        else:
            cur = cur.to(sp["kind"], **kw)
        if sp["on_error"]:
            cur.error_handler(name=f"catch{i}", class_name="Handled")
    cur.respond()
    return fn.to_mock_server(namespace=ns)

rnd = random.Random(11); n = 0; diffs = 0
for gi in range(500):
    g = d.random_graph(rnd)
    g["steps"] = [s for s in g["steps"] if s["kind"] != "Boom"] or g["steps"][:1]
    if any(s["kind"] == "Boom" for s in g["steps"]): continue
    try:
        a = build(ref, g, "sync"); b = build(mine, g, "async")
    except Exception as e:
        print("build", e); continue
    for body in d.BODIES:
        out = []
        for s in (a, b):
            try:
                r = s.test(body=copy.deepcopy(body), silent=True)
                out.append(("ok", json.dumps(_resp(r), sort_keys=True, default=str)))
            except Exception as exc:
                out.append(("exc", type(exc).__name__, _first_line(exc)))
        out = [tuple(re.sub(r"<[\w.]*MockEvent object at 0x[0-9a-f]+>", "<MockEvent>", re.sub(r"[0-9a-f]{32}", "<id>", x)) if isinstance(x, str) else x for x in o) for o in out]
        n += 1
        if out[0] != out[1]:
            diffs += 1
            if diffs <= 4:
                print("DIFF", json.dumps(g), body); print("  ref sync  :", out[0][:300]); print("  mine async:", out[1][:300])
    try: b.wait_for_completion()
    except Exception: pass
print("identical on" if not diffs else "DIFFS in", n, "graph x body cases (async product vs sync reference)", "" if not diffs else diffs)
sys.exit(1 if diffs else 0)
