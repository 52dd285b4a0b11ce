"""Where the read-only reference tree is present (the build container: /root/reference), run the differential checks that pin
the kernels' three CHECKERS -- the batched scoring oracle, the columnar ingest oracle, the enrichment oracle -- against the REAL
reference classes on seeded random workloads, live, as part of the CPU suite (each in its own process: importing the reference
installs import hooks).  Elsewhere (the GPU box has no reference tree) the tests skip; the committed goldens cover that case.
The other differential scripts (`python -m tests.golden.run_diffs`, ~3 minutes) stay out of the suite."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/mlrun/serving"), reason="the reference tree is not on this machine")


@needs_reference
@pytest.mark.parametrize("script,verdict", [("diff_hot_path", "the batched oracle equals the real reference"),
                                            ("diff_ingest", "ingest_columns equals the real reference"),
                                            ("diff_online", "identical on 500 random online services")])
def test_kernel_checkers_equal_the_real_reference(script, verdict):
    done = subprocess.run([sys.executable, "-m", f"tests.golden.{script}"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert done.returncode == 0, (done.stdout[-1500:], done.stderr[-1500:])
    assert verdict in done.stdout, done.stdout[-800:]
