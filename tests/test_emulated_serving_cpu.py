"""The product's whole host layer around device plans -- lowering, packing, device model servers, routers, run_batch /
run_events -- on CPU, with the plans' arithmetic emulated in numpy (tests/emulated_plan.py), against the golden outputs of
the REAL reference.  The same scenarios run on the real kernels in tests/test_gpu_serving.py."""

import json
import os

import numpy as np
import pytest

from mlrun_b200.synthetic import flow3_workload, tree_workload
from oracle import batch as obatch
from tests import api_b200, emulated_plan, scenarios
from tests.compare import assert_same

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scenarios.json")))
RTOL, ATOL = 1e-5, 1e-5


@pytest.fixture(autouse=True)
def _emulated(monkeypatch):
    emulated_plan.install(monkeypatch)


@pytest.mark.parametrize("name", ["flow3_linear_events", "flow3_ensemble_events", "tree_ensemble_batch"])
def test_device_scenarios_on_the_emulated_plan_match_reference_golden(name):
    got = json.loads(json.dumps(getattr(scenarios, name)(api_b200), default=str))
    assert_same(got, GOLDEN[name], name, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("n_models", [1, 4])
@pytest.mark.parametrize("engine", ["sync", "async"])
def test_run_batch_and_run_events_equal_the_per_event_path(n_models, engine):
    wl = flow3_workload(n_rows=256, n_num=56, n_cat=8, seed=9, n_models=n_models)
    server = wl.build_server(api_b200, engine=engine)
    out, status = server.run_batch(wl.X, names=wl.names, with_status=True)
    ref = obatch.flow3(wl)["out"]
    np.testing.assert_allclose(out[:, 0], ref, rtol=RTOL, atol=ATOL)
    assert not status.any()
    path = "/" if n_models == 1 else "/v2/models/infer"
    rows = wl.rows_as_dicts(limit=16)
    for i, row in enumerate(rows):
        got = server.test(path=path, body=dict(row))["outputs"][0]
        assert abs(got - ref[i]) <= ATOL + RTOL * abs(ref[i])
    resp = server.run_events(rows)
    np.testing.assert_allclose([r["outputs"][0] for r in resp], ref[:16], rtol=RTOL, atol=ATOL)
    assert resp[0]["model_name"] == ("linear" if n_models == 1 else "ensemble")


def test_router_of_tree_models_and_single_routes():
    wl = tree_workload(n_rows=128, n_feat=24, n_models=4, n_trees=10, depth=4, seed=6, n_fit=800)
    server = wl.build_server(api_b200)
    ref = obatch.tree_ensemble(wl)
    np.testing.assert_allclose(server.run_batch(wl.X)[:, 0], ref["out"], rtol=RTOL, atol=ATOL)
    one = server.test("/v2/models/m2/infer", body={"inputs": wl.X[:8].astype(np.float64).tolist()})
    np.testing.assert_allclose(one["outputs"], ref["per_model"][:8, 1], rtol=RTOL, atol=ATOL)
    assert one["model_name"] == "m2"


def test_enrichment_routers_served_match_the_real_reference(monkeypatch):
    """`enrichment_routers` golden (the REAL Enrichment routers over a stubbed store read) through the product's routers;
    the device table is the numpy stand-in of tests/test_online_host_cpu.py"""
    from mlrun_b200.feature_store import online as bo
    from tests import test_online_host_cpu as host

    monkeypatch.setattr(bo, "DeviceTable", host._HostTable)

    class Api:
        def __getattr__(self, name):
            return getattr(api_b200, name)

        @staticmethod
        def register_online_vector(uri, features, index_keys, table, stats, label_column, with_indexes):
            frame = host._frame(features, index_keys, table)
            api_b200.register_feature_vector(uri, bo.FeatureVector("vec", features, index_keys, frame, stats, label_column=label_column,
                                                                   with_indexes=with_indexes))

    got = json.loads(json.dumps(scenarios.enrichment_routers(Api()), default=str))
    assert_same(got, GOLDEN["enrichment_routers"], "enrichment_routers", rtol=RTOL, atol=ATOL)


def _gpu_serving_cases():
    from tests import test_gpu_serving as g  # its own tests are `-m gpu`; here their bodies run on the emulated plan

    return [g.test_run_events_reports_bad_rows_as_400, g.test_router_of_tree_models_run_batch_and_single_route,
            g.test_unlowerable_graph_is_a_hard_error, g.test_run_json_answers_like_the_reference_wire_path,
            g.test_tracked_batches_emit_the_per_event_records]


@pytest.mark.parametrize("case", _gpu_serving_cases(), ids=lambda f: f.__name__)
def test_gpu_serving_cases_hold_on_the_emulated_plan(case):
    """the host-side assertions of tests/test_gpu_serving.py (400s for flagged rows, single routes, the wire path through
    the C body codec, tracked batches) do not depend on the device: they must hold with the numpy plan too"""
    case()


def _gpu_parity_cases():
    from tests import test_gpu_parity as g

    return [
        (g.test_flow3_matches_oracle, dict(n_models=1, n_rows=129)), (g.test_flow3_matches_oracle, dict(n_models=4, n_rows=1000)),
        (g.test_flow3_per_model_outputs_and_per_event_oracle, {}),
        (g.test_flow3_matches_reference_golden, dict(name="flow3_linear_events", n_models=1)),
        (g.test_flow3_matches_reference_golden, dict(name="flow3_ensemble_events", n_models=4)),
        (g.test_onehot_edge_values_and_sparse_categories, {}), (g.test_nonfinite_input_sets_row_status, {}),
        (g.test_transform_only_plan_is_exact, {}), (g.test_map_values_and_drop, {}),
        (g.test_tree_ensemble_regression, dict(n_rows=65)), (g.test_tree_ensemble_classification_is_bit_exact, {}),
        (g.test_tree_ensemble_matches_reference_golden, {}), (g.test_mixed_linear_and_tree_ensemble_with_onehot, {}),
        (g.test_votes_match_reference_golden, {}), (g.test_majority_vote_random_against_numpy, {}),
        (g.test_logistic_ensemble_majority_vote_is_exact, {}), (g.test_range_bounds_that_are_not_float32_on_the_device, {}),
    ]


@pytest.mark.parametrize("case,kwargs", _gpu_parity_cases(), ids=lambda v: v.__name__ if callable(v) else "-".join(map(str, v.values())))
def test_gpu_parity_cases_hold_on_the_emulated_plan(case, kwargs):
    """the plan-level parity assertions of tests/test_gpu_parity.py (oracle / golden equality of transforms, linear and
    tree models, votes, status words) with the numpy plan: pins the lowering and the packing on CPU, and keeps the
    emulation honest against the very assertions the kernels pass on the GPU"""
    case(**kwargs)
