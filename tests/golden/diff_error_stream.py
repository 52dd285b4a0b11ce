"""Differential check of what failing events leave on the ERROR STREAM (serving/server.py:542-556, 605-614; the records
`context.push_error` formats) against the REAL reference (build container only): random sync flows with raising steps -- with
and without step-level / graph-level error handlers -- and routers with failing models, bad JSON bodies, illegal paths;
`error_stream` set on the function.  The records pushed per request (source, message head, event body, args, keys) and the
responses are compared.

    python -m tests.golden.diff_error_stream
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
from tests.golden.diff_flow_graphs import BODIES, namespace, random_graph  # noqa: E402
from tests.scenarios import _clean, _first_line, _resp  # noqa: E402


def shape(rec):
    out = {"keys": sorted(rec), "source": rec.get("source"), "args": _clean(rec.get("args")),
           "message_head": _first_line(rec.get("message", "")), "event_keys": sorted(rec.get("event", {})),
           "event_body": repr(rec.get("event", {}).get("body")),  # (repr: a body may hold the event that holds it)
           "has_trace": "Traceback" in str(rec.get("message", ""))}
    return out


def build_flow(api, g):
    fn = api.new_function("f", kind="serving")
    flow = fn.set_topology("flow", engine="sync")
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
        if g["respond_at"] == i:
            cur.respond()
    if g["graph_error_handler"]:
        flow.error_handler(name="catch_all", class_name="Handled")
    server = fn.to_mock_server(namespace=namespace(api))
    server.set_error_stream("dummy://")  # (to_mock_server does not carry spec.error_stream over: server.py:139-145 is the way in)
    return server


def records(server):
    stream = server._error_stream_object
    return [shape(r) for r in getattr(stream, "event_list", [])]


def main():
    rnd = random.Random(53)
    n = total_records = 0
    for gi in range(500):
        g = random_graph(rnd)
        if not any(s["kind"] == "Boom" for s in g["steps"]) and rnd.random() < 0.7:
            g["steps"][rnd.randrange(len(g["steps"]))].update(kind="Boom", args={}, on_error=rnd.random() < 0.4)
            g["steps"] = [{k: v for k, v in s.items() if k != "handler" or s["kind"] == "Multi"} for s in g["steps"]]
        servers = []
        for api in (ref, mine):
            try:
                servers.append(("ok", build_flow(api, g)))
            except Exception as exc:  # noqa: BLE001
                servers.append(("exc", type(exc).__name__, _first_line(exc)))
        if [s[0] for s in servers] != ["ok", "ok"]:
            assert servers[0][0] == servers[1][0] and servers[0][1:] == servers[1][1:], (g, servers)
            continue
        for body in BODIES[:5] + ['{"bad json"', b"\xff"]:
            out = []
            for _state, server in servers:
                before = len(records(server))
                try:
                    kw = {"content_type": "application/json"} if isinstance(body, (str, bytes)) else {}
                    r = server.test(body=copy.deepcopy(body), silent=True, **kw)
                    resp = ("ok", json.dumps(_resp(r), sort_keys=True, default=str))
                except Exception as exc:  # noqa: BLE001
                    resp = ("exc", type(exc).__name__, _first_line(exc))
                new = records(server)[before:]
                total_records += len(new)
                text = json.dumps({"resp": resp, "records": new}, sort_keys=True, default=str)
                out.append(re.sub(r"<[\w.]*MockEvent object at 0x[0-9a-f]+>", "<MockEvent>", re.sub(r"[0-9a-f]{32}", "<id>", text)))
            n += 1
            if out[0] != out[1]:
                print("DIFF", json.dumps(g), "body", body)
                print("  ref :", out[0][:1200])
                print("  mine:", out[1][:1200])
                return 1
    assert total_records > 500, total_records
    print("identical on", n, "requests (responses and", total_records, "error-stream records)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
