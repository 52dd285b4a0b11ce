"""FeatureSet.ingest / IngestPlan.run end to end on CPU: column extraction, result block, output dtypes, violation and miss
counters and the DataFrame assembly of the product, with the columnar kernel's arithmetic emulated in numpy
(tests/device_emulator.py EmulatedColumns), against the per-row and the vectorised oracle -- the very assertions
tests/test_gpu_ingest.py makes on the real kernel."""

import pytest

from tests import device_emulator


@pytest.fixture(autouse=True)
def _emulated(monkeypatch):
    device_emulator.install_columns(monkeypatch)


def _cases():
    from tests import test_gpu_ingest as g  # its own tests are `-m gpu`; here their bodies run on the emulation

    return [(g.test_config5_matches_the_per_row_reference_walk, dict(n_rows=1)),
            (g.test_config5_matches_the_per_row_reference_walk, dict(n_rows=5)),
            (g.test_config5_matches_the_per_row_reference_walk, dict(n_rows=4097)),
            (g.test_config5_matches_the_per_row_reference_walk, dict(n_rows=20000)),
            (g.test_edge_values_match_the_reference_semantics, {}), (g.test_nat_and_wide_date_range, {}),
            (g.test_feature_set_ingest_api_and_dropped_validated_column, {})]


@pytest.mark.parametrize("case,kwargs", _cases(), ids=lambda v: v.__name__ if callable(v) else "-".join(map(str, v.values())))
def test_gpu_ingest_cases_hold_on_the_emulated_columns_plan(case, kwargs):
    case(**kwargs)
