"""A small tour of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python profiles/lab/sanitize_smoke.py
Plans are tiny (the tools slow kernels down 10-100x); results are still compared with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.smoke()  # rowthread (TMA), trees3 (prep + walk + vote), columns, table gather, codec

from mlrun_b200 import _native as nat, packing  # noqa: E402
from mlrun_b200.lowering import ColumnProgram  # noqa: E402
from mlrun_b200.synthetic import tree_workload  # noqa: E402

rng = np.random.default_rng(0)
# dense head (tcgen05 + TMEM)
from sklearn.linear_model import LinearRegression  # noqa: E402

models = []
for i in range(12):
    m = LinearRegression()
    m.coef_, m.intercept_, m.n_features_in_ = rng.normal(size=64), float(rng.normal()), 64
    models.append(m)
X = rng.normal(size=(700, 64)).astype(np.float32)
plan = ColumnProgram([f"f{i}" for i in range(64)]).build_plan([packing.pack_model(m) for m in models])
out = plan.run(X)
np.testing.assert_allclose(out, np.stack([m.predict(X.astype(np.float64)) for m in models], axis=1), rtol=1e-5, atol=1e-5)
print("dense:", plan.kernel)
# trees3 with NaN routing, ragged tiles, the coalescing ring and the zero-copy path
from sklearn.ensemble import RandomForestRegressor  # noqa: E402

Xf = rng.normal(size=(500, 32)).astype(np.float32)
rf = RandomForestRegressor(n_estimators=9, max_depth=5, random_state=0).fit(Xf, Xf[:, 0] * 2 + Xf[:, 1])
Xt = rng.normal(size=(333, 32)).astype(np.float32)
Xt[rng.random(Xt.shape) < 0.1] = np.nan
plan = ColumnProgram([f"f{i}" for i in range(32)]).build_plan([packing.pack_model(rf)])
np.testing.assert_allclose(plan.run(Xt)[:, 0], rf.predict(Xt.astype(np.float64)), rtol=1e-5, atol=1e-5)
tickets = [plan.submit(Xt[i:i + 37]) for i in range(0, 333, 37)]
plan.flush()
got = np.concatenate([plan.wait(t) for t in tickets])
np.testing.assert_allclose(got[:, 0], rf.predict(Xt.astype(np.float64)), rtol=1e-5, atol=1e-5)
print("trees3 + ring:", plan.kernel)
print("sanitize smoke ok")
