"""the `ring` part of the default bench line by itself (profiles/lab/gpu9.sh)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from mlrun_b200 import _native as nat  # noqa: E402

nat.init(0)
r = bench.ring_bench(nat)
for row in r.get("native", []):
    print(json.dumps(row))
print(json.dumps({k: v for k, v in r.items() if k != "native"}))
