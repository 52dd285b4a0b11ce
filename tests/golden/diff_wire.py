"""Differential check of the wire level -- `GraphServer.run(MockEvent(...))` and `server.test(...)` with every envelope option
(serving/server.py:196-308, 445-490) -- against the REAL reference (build container only): bodies as dict / str / bytes /
malformed JSON / empty, content types (none, json, application/json, text/plain, image/png), MLRUN-EVENT-ID / MLRUN-EVENT-PATH
headers, methods, `get_body`, `silent`, explicit event ids, a flow that responds with a dict / str / bytes / None / a number /
a list, a model that raises.  Response objects (status, content type, body) and exceptions compared.

    python -m tests.golden.diff_wire
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
from tests.scenarios import _first_line  # noqa: E402


def namespace(api):
    class Answer:
        def __init__(self, what="dict", **kw):
            self.what = what

        def do_event(self, event):
            seen = {"path": event.path, "method": event.method, "ct": event.content_type, "id_len": len(str(event.id)),
                    "body_type": type(event.body).__name__}
            event.body = {"dict": {"seen": seen, "body": event.body if not isinstance(event.body, bytes) else "bytes"}, "str": "plain answer",
                          "bytes": b"raw answer", "none": None, "number": 7.5, "list": [1, {"a": 2}], "empty": {}, "raise": None}[self.what]
            if self.what == "raise":
                raise ValueError("step failed")
            return event

    return {"Answer": Answer}


def show(r):
    if hasattr(r, "status_code"):
        body = r.body
        return ("response", r.status_code, getattr(r, "content_type", None), body.decode() if isinstance(body, bytes) else body)
    return ("object", type(r).__name__, r.decode() if isinstance(r, bytes) else r)


def norm(x):
    return re.sub(r"[0-9a-f]{32}", "<id>", json.dumps(x, sort_keys=True, default=str))


def main():
    rnd = random.Random(13)
    n = 0
    bodies = [{"a": 1}, '{"a": 1}', b'{"a": [1, 2]}', '{"a": 1', b"\xff\xfe", "", b"", None, "text", 5, [1, 2], '"just a string"', "[1, 2]"]
    cts = [None, "", "json", "application/json", "text/plain", "image/png", "application/x-other"]
    for what in ("dict", "str", "bytes", "none", "number", "list", "empty", "raise"):
        servers = []
        for api in (ref, mine):
            fn = api.new_function("f", kind="serving")
            flow = fn.set_topology("flow", engine="sync")
            flow.to("Answer", name="a", what=what).respond()
            servers.append((api, fn.to_mock_server(namespace=namespace(api))))
        for _ in range(260):
            body, ct = rnd.choice(bodies), rnd.choice(cts)
            headers = rnd.choice([None, {}, {"MLRUN-EVENT-ID": "abc"}, {"MLRUN-EVENT-PATH": "/x/y"}, {"MLRUN-EVENT-ID": "i", "MLRUN-EVENT-PATH": "/p", "other": "1"}])
            method = rnd.choice(["POST", "GET", None])
            path = rnd.choice(["/", "/q", None, ""])
            mode = rnd.choice(["run", "run_get_body", "test", "test_get_body_false", "test_silent", "test_event_id"])
            out = []
            for api, server in servers:
                try:
                    if mode.startswith("run"):
                        ev = api.MockEvent(body=copy.deepcopy(body), content_type=ct, headers=copy.deepcopy(headers), method=method, path=path)
                        r = server.run(ev, get_body=mode == "run_get_body")
                    else:
                        kw = {"content_type": ct, "headers": copy.deepcopy(headers), "method": method or "", "path": path or "/"}
                        if mode == "test_get_body_false":
                            kw["get_body"] = False
                        if mode == "test_silent":
                            kw["silent"] = True
                        if mode == "test_event_id":
                            kw["event_id"] = "given-id"
                        r = server.test(body=copy.deepcopy(body), **kw)
                    out.append(norm(show(r)))
                except Exception as exc:  # noqa: BLE001
                    out.append(norm(("exc", type(exc).__name__, _first_line(exc))))
            n += 1
            if out[0] != out[1]:
                print("DIFF", what, mode, repr(body), ct, headers, method, path)
                print("  ref :", out[0][:700])
                print("  mine:", out[1][:700])
                return 1
    print("identical on", n, "requests")
    return 0


if __name__ == "__main__":
    sys.exit(main())
