"""end-to-end latency of one serving batch from pinned host rows (DevicePlan.run: in -> kernel -> out + status), per batch
size; run with B2S_ZEROCOPY_IN_BYTES=<bytes> to move the threshold under which the kernel reads the rows straight from
pinned host memory (no H2D copy on the copy engine)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mlrun_b200 import _native as nat, api  # noqa: E402
from mlrun_b200.synthetic import flow3_workload  # noqa: E402

nat.init(0)
wl = flow3_workload(n_rows=65536, n_num=56, n_cat=8, seed=2, n_models=4)
server = wl.build_server(api, engine="sync")
plan = server.compile(wl.names).plan
ref = plan.run(wl.X[:16384])
for rows in (256, 1024, 4096, 16384):
    h = nat.pinned_empty((rows, wl.X.shape[1]), np.float32)
    h[:] = wl.X[:rows]
    for _ in range(100):
        out = plan.run(h, with_status=True)
    lat = []
    for _ in range(1000):
        t0 = time.perf_counter()
        out = plan.run(h, with_status=True)
        lat.append((time.perf_counter() - t0) * 1e6)
    ok = np.array_equal(out[0], ref[:rows])
    print(f"zc_in_bytes={os.environ.get('B2S_ZEROCOPY_IN_BYTES', 'default')} rows={rows} p50={np.percentile(lat, 50):.1f}us p99={np.percentile(lat, 99):.1f}us same_result={ok}")
