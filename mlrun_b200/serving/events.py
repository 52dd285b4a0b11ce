"""Event envelope of the serving host (plugin-API mirror of mlrun/serving/server.py:437-490)."""

import uuid


class MockTrigger:
    """stand-in for a nuclio trigger descriptor: `.kind` ("http", "kafka", ...) and `.name`"""

    def __init__(self, kind="", name=""):
        self.kind = kind
        self.name = name


class MockEvent:
    """the attributes the path reads (id, key, body, headers, method, path, content_type, trigger,
    offset) and writes (error, terminated, origin_state, stream_path); see server.py:445-475"""

    def __init__(self, body=None, content_type=None, headers=None, method=None, path=None, event_id=None,
                 trigger=None, offset=None, time=None):
        self.id = event_id if event_id else uuid.uuid4().hex
        self.key = ""
        self.body = body
        self.headers = headers if headers else {}
        self.method = method
        self.path = path if path else "/"
        self.content_type = content_type
        self.error = None
        self.trigger = trigger if trigger else MockTrigger()
        self.offset = offset if offset else 0

    def __str__(self):
        tail = f", error={self.error}" if self.error else ""
        return f"Event(id={self.id}, body={self.body}, method={self.method}, path={self.path}{tail})"


class Response:
    """HTTP-ish response object handed back for non-200 results and by the nuclio handler path"""

    def __init__(self, headers=None, body=None, content_type=None, status_code=200):
        self.headers = headers if headers else {}
        self.body = body
        self.status_code = status_code
        self.content_type = content_type if content_type else "text/plain"

    def __repr__(self):
        fields = ", ".join(f"{k}={v!r}" for k, v in vars(self).items())
        return f"{type(self).__name__}({fields})"
