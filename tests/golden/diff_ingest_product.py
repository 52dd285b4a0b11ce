"""The product's feature-set ingest (`FeatureSet.ingest(DataFrame)`: validate_steps, lowering of the six steps to ONE columnar
plan, column extraction, result block, dtypes, DataFrame assembly) against the REAL reference step classes walking the frame
one row at a time (build container only).  The columnar kernel's arithmetic is the numpy emulation of
tests/device_emulator.py (the kernel itself is compared with the oracle in `-m gpu`), so what this pins is the HOST side of the
ingest path on random config-5-shaped workloads: which column gets which op with which constants, the output schema and order,
the violation count.

    python -m tests.golden.diff_ingest_product
"""
import contextlib
import io
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from mlrun_b200.feature_store import ingest as bingest  # noqa: E402
from mlrun_b200.feature_store import steps as bsteps  # noqa: E402
from mlrun_b200.synthetic import ingest_workload  # noqa: E402
from tests import device_emulator  # noqa: E402
from tests.golden.diff_ingest import RefSteps, reference_rows  # noqa: E402


class _Patch:  # the two attributes install_columns sets, without pytest
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def main():
    device_emulator.install_columns(_Patch())
    rnd = random.Random(43)
    rows = 0
    for case in range(12):
        wl = ingest_workload(n_rows=rnd.randint(150, 400), seed=400 + case, n_f32=rnd.choice([24, 32, 48]), n_cat=rnd.choice([8, 12]),
                             n_counter=rnd.choice([2, 5]), nan_frac=rnd.choice([0.02, 0.1, 0.3]))
        want, n_printed = reference_rows(wl.build_steps(RefSteps), wl.df)
        fset = bingest.FeatureSet(f"case{case}", timestamp_key="timestamp")
        cur = fset.graph
        for st in wl.build_steps(bsteps):
            cur = cur.to(st)
        for c in wl.checked_cols:
            fset[c] = bingest.Feature(validator=bsteps.MinMaxValidator(severity="info", min=-2.5, max=2.5))
        printed = io.StringIO()
        with contextlib.redirect_stdout(printed):
            got = fset.ingest(wl.df)
        assert list(got.columns) == list(want.columns), (case, [c for c in got.columns if c not in set(want.columns)][:5],
                                                         [c for c in want.columns if c not in set(got.columns)][:5])
        for c in want.columns:
            a, b = got[c].to_numpy(), want[c].to_numpy()
            if a.dtype.kind == "f" or b.dtype.kind == "f":
                assert np.array_equal(a.astype(np.float64), b.astype(np.float64), equal_nan=True), (case, c, a[:5], b[:5])
            elif a.dtype.kind == "M" or b.dtype.kind == "M":
                assert (a.astype("datetime64[ns]") == b.astype("datetime64[ns]")).all(), (case, c)
            else:
                assert (a == b).all(), (case, c, a[:5], b[:5])
        # the engine reports violations per column and batch ("info! x16 has 3 values outside [-2.5, 2.5]"), the reference one
        # line per offending value: the totals must agree
        import re

        total = sum(int(m.group(1)) for m in re.finditer(r" has (\d+) values? outside", printed.getvalue()))
        assert total == n_printed, (case, total, n_printed)
        rows += len(wl.df)
    print("FeatureSet.ingest equals the real reference's row walk on", rows, "rows of 12 random workloads")
    return 0


if __name__ == "__main__":
    sys.exit(main())
