"""fused (scoring kernel gathers its rows) vs gather-then-score, device-resident keys, no torch.
python profiles/prof_enrich_fused.py [n_keys] [batch]"""
import sys
import time

import numpy as np
from sklearn.linear_model import LinearRegression

from mlrun_b200 import _native as nat
from mlrun_b200 import packing
from mlrun_b200.feature_store.online import DeviceTable
from mlrun_b200.lowering import ColumnProgram

n_keys = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
F = 64
nat.init(0)
rng = np.random.default_rng(1)
vals = rng.standard_normal((n_keys, F), dtype=np.float32)
vals[::50, 3] = np.nan
keys = np.arange(n_keys, dtype=np.int64) * 7919 + 13
table = DeviceTable(keys, vals, np.zeros(F, dtype=np.float32))
models = []
for i in range(4):
    m = LinearRegression()
    m.coef_, m.intercept_, m.n_features_in_ = rng.normal(size=F), 0.0, F
    models.append(packing.pack_model(m))
plan = ColumnProgram([f"f{i}" for i in range(F)]).build_plan(models, vote=(nat.VOTE_MEAN, [0.25] * 4))
ask = [keys[rng.integers(0, n_keys, size=B)] for _ in range(3)]
d_keys = [nat.DeviceBuffer(B * 8).upload(a) for a in ask]
d_rows, d_out = nat.DeviceBuffer(B * F * 4), nat.DeviceBuffer(B * 4)
sync = nat.load().b2s_device_sync


def timed(fn, n=40):
    for i in range(3):
        fn(i)
    sync()
    t = time.perf_counter()
    for i in range(n):
        fn(i)
    sync()
    return (time.perf_counter() - t) / n * 1e3


def two(i):
    table.lookup_device(d_keys[i % 3].ptr, B, d_rows.ptr, F * 4)
    plan.run_device(d_rows.ptr, B, F * 4, d_out.ptr)


fused_ok = table.enrich_device(plan, d_keys[0].ptr, B, d_out.ptr)
ms2 = timed(two)
msf = timed(lambda i: table.enrich_device(plan, d_keys[i % 3].ptr, B, d_out.ptr)) if fused_ok else float("nan")
print(f"n_keys {n_keys} batch {B}: gather+score {ms2:.4f} ms, fused {msf:.4f} ms ({plan.kernel})")
