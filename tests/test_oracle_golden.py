"""The CPU oracle against (a) golden outputs generated from the REAL reference
(tests/golden/scenarios.json, made by tests/golden/gen_golden.py) and (b) the literal expectations
of the reference's own tests (scenario.EXPECT).  CPU only."""

import json
import os

import pytest

from tests import api_oracle, scenarios
from tests.compare import assert_same

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scenarios.json")))


@pytest.mark.parametrize("scenario", scenarios.SCENARIOS, ids=lambda f: f.__name__)
def test_oracle_matches_reference(scenario):
    got = json.loads(json.dumps(scenario(api_oracle), default=str))
    for path, want in getattr(scenario, "EXPECT", {}).items():
        assert scenarios.dig(got, path) == json.loads(json.dumps(want)), f"{scenario.__name__}{path}"
    if getattr(scenario, "ASYNC", False):
        return  # storey engine: pinned by the reference tests' literals only (see oracle/topology.py)
    want = GOLDEN[scenario.__name__]
    for path in getattr(scenario, "GOLDEN_SKIP", []):
        scenarios.dig(got, path[:-1]).pop(path[-1])
        scenarios.dig(want, path[:-1]).pop(path[-1], None)
    # the oracle runs the same numpy / scikit-learn arithmetic as the reference: results are bit-equal
    assert_same(got, want, scenario.__name__)


def test_golden_covers_all_sync_scenarios():
    names = {f.__name__ for f in scenarios.SCENARIOS if not getattr(f, "ASYNC", False)}
    assert names == set(GOLDEN.keys())
