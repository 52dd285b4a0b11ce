"""Where the wall time of one FeatureSet.ingest(DataFrame) call goes (host (dis)assembly vs PCIe vs kernel).
Run on a GPU box: python profiles/prof_ingest_host.py [n_rows]"""
import contextlib
import io
import sys
import time

import numpy as np

from mlrun_b200 import _native as nat
from mlrun_b200.feature_store import ingest as bi
from mlrun_b200.feature_store import steps as bs
from mlrun_b200.synthetic import ingest_workload

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nat.init(0)
wl = ingest_workload(n_rows=n, seed=5)
plan = bi.lower_steps(wl.build_steps(bs), wl.df)


def timed(fn, reps=8):
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(3):
            fn()
        t = time.perf_counter()
        for _ in range(reps):
            out = fn()
    return (time.perf_counter() - t) / reps * 1e3, out


total, _ = timed(lambda: plan.run(wl.df))
print("plan.run total ms", round(total, 3), "device stats", {k: round(v, 3) if isinstance(v, float) else v for k, v in plan.stats.items()})
ins_ms, (ins, _k) = timed(lambda: plan._inputs(wl.df))
print("_inputs ms", round(ins_ms, 3))
outs = {s: np.empty(n * 2, dtype=np.int32) for s in range(plan.plan.n_out)}
host_ms, _ = timed(lambda: plan.plan.run_host(ins, n, outs, with_stats=True))
print("run_host (pageable in, pageable out) ms", round(host_ms, 3))
pin_in = {}
for s, a in ins.items():
    p = nat.pinned_empty(a.shape, a.dtype)
    p[:] = a
    pin_in[s] = p
pin_ms, (_c, st) = timed(lambda: plan.plan.run_host(pin_in, n, outs, with_stats=True))
print("run_host (pinned in, pageable out) ms", round(pin_ms, 3), st)
