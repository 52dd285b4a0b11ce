"""Tracking-stream records of the batched engine path: `push_batch` must emit exactly what the
reference's per-event `_ModelLogPusher.push` (serving/v2_serving.py:457-504, restated in oracle/model_protocol.py and
pinned by the `tracking` golden scenario) emits for the same events."""

import datetime
import types

import pytest

from mlrun_b200.serving import model_server as bms
from oracle import model_protocol as oms


class _Stream:
    def __init__(self):
        self.records = []

    def push(self, data, **kw):
        self.records.extend(data)


def _pusher(mod, sample, batch):
    stream = _Stream()
    params = {"log_stream_batch": batch, "log_stream_sample": sample}
    ctx = types.SimpleNamespace(verbose=False, worker_id=3,
                                stream=types.SimpleNamespace(hostname="h", function_uri="p/f", stream_uri="s", output_stream=stream, enabled=True),
                                get_param=lambda k, default=None: params.get(k, default))
    model = types.SimpleNamespace(name="m1", version="v2", metrics={"acc": 1}, labels={"a": "b"})
    model.__class__.__name__  # noqa: B018
    return mod._ModelLogPusher(model, ctx), stream


def _strip(records):
    out = []
    for r in records:
        r = dict(r)
        r.pop("when", None)
        r.pop("microsec", None)
        if "values" in r:
            r["values"] = [[v[0], v[1], v[2], v[5]] for v in r["values"]]
        out.append(r)
    return out


@pytest.mark.parametrize("sample,batch", [(1, 1), (3, 1), (1, 4), (5, 3), (7, 2)])
def test_push_batch_equals_per_event_pushes(sample, batch):
    start = datetime.datetime(2026, 1, 1, tzinfo=datetime.timezone.utc)
    reqs = [{"id": f"e{i}", "inputs": [[i, i + 0.5]]} for i in range(53)]
    resps = [{"id": f"e{i}", "model_name": "m1", "outputs": [i * 2.0]} for i in range(53)]
    ref, ref_stream = _pusher(oms, sample, batch)
    for rq, rs in zip(reqs, resps):
        ref.push(start, rq, rs, "infer")
    got, got_stream = _pusher(bms, sample, batch)
    built = []

    def lazy_req(i):
        built.append(i)
        return reqs[i]

    lazy_req.n = 53
    # three engine batches of uneven size: positions and micro-batch boundaries carry over between calls
    for lo, hi in [(0, 10), (10, 11), (11, 53)]:
        sub = lambda i, lo=lo: lazy_req(lo + i)  # noqa: E731
        sub.n = hi - lo
        got.push_batch(start, sub, lambda i, lo=lo: resps[lo + i], "infer", microsec=7)
    assert _strip(got_stream.records) == _strip(ref_stream.records)
    assert len(built) == len(range(sample - 1, 53, sample))  # only sampled rows were materialised
