"""The product's host layer (mlrun_b200.serving: graph building, per-event executors, routers, V2
protocol, feature steps on dict events) against the golden outputs of the REAL reference and the
literal expectations of its tests.  CPU only: scenarios whose arithmetic runs on the device are in
tests/test_gpu_serving.py."""

import json
import os

import pytest

from tests import api_b200, scenarios
from tests.compare import assert_same

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scenarios.json")))
DEVICE = {"flow3_linear_events", "flow3_ensemble_events", "tree_ensemble_batch", "online_service_logic",
          "enrichment_routers", "pickle_model_from_path"}  # need the GPU
PANDAS = {"steps_pandas_engine", "validator_pandas"}  # DataFrame bodies are not a host path of the engine (see transforms.py)
HOST = [s for s in scenarios.SCENARIOS if s.__name__ not in DEVICE | PANDAS]


@pytest.mark.parametrize("scenario", HOST, ids=lambda f: f.__name__)
def test_host_matches_reference(scenario):
    got = json.loads(json.dumps(scenario(api_b200), default=str))
    for path, want in getattr(scenario, "EXPECT", {}).items():
        assert scenarios.dig(got, path) == json.loads(json.dumps(want)), f"{scenario.__name__}{path}"
    if getattr(scenario, "ASYNC", False):
        return
    want = GOLDEN[scenario.__name__]
    if scenario.__name__ == "step_to_dict":
        return  # class paths differ by design (mlrun_b200.* vs mlrun.*); the literal EXPECT above pins the shape
    assert_same(got, want, scenario.__name__)
