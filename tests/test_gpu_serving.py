"""The product's serving API on the GPU: the reference's graph-building calls, per-event `server.test`,
and the engine's batched `run_batch` / `run_events`, checked against the golden outputs of the REAL
reference and against the CPU oracle.  `-m gpu`."""

import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from mlrun_b200.lowering import LoweringError  # noqa: E402
from mlrun_b200.synthetic import flow3_workload, tree_workload  # noqa: E402
from oracle import batch as obatch  # noqa: E402
from tests import api_b200, api_oracle, scenarios  # noqa: E402
from tests.compare import assert_same  # noqa: E402

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scenarios.json")))
RTOL, ATOL = 1e-5, 1e-5


@pytest.mark.parametrize("name", ["flow3_linear_events", "flow3_ensemble_events", "tree_ensemble_batch", "pickle_model_from_path"])
def test_device_scenarios_match_reference_golden(name):
    """same scenario code as the oracle / reference runs, through mlrun_b200's API: every predict is a CUDA plan"""
    got = json.loads(json.dumps(getattr(scenarios, name)(api_b200), default=str))
    want = GOLDEN[name]
    assert_same(got, want, name, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("n_models", [1, 4])
@pytest.mark.parametrize("engine", ["sync", "async"])
def test_flow3_run_batch_equals_per_event_and_oracle(n_models, engine):
    wl = flow3_workload(n_rows=2048, n_num=56, n_cat=8, seed=9, n_models=n_models)
    server = wl.build_server(api_b200, engine=engine)
    out, status = server.run_batch(wl.X, names=wl.names, with_status=True)
    ref = obatch.flow3(wl)["out"]
    np.testing.assert_allclose(out[:, 0], ref, rtol=RTOL, atol=ATOL)
    assert not status.any()
    path = "/" if n_models == 1 else "/v2/models/infer"
    rows = wl.rows_as_dicts(limit=32)
    for i, row in enumerate(rows):  # per-event contract: feature steps on the host, predict on the device
        got = server.test(path=path, body=dict(row))["outputs"][0]
        assert abs(got - ref[i]) <= ATOL + RTOL * abs(ref[i])
    resp = server.run_events(rows)
    np.testing.assert_allclose([r["outputs"][0] for r in resp], ref[:32], rtol=RTOL, atol=ATOL)
    assert resp[0]["model_name"] == ("linear" if n_models == 1 else "ensemble")


def test_run_events_reports_bad_rows_as_400():
    wl = flow3_workload(n_rows=16, n_num=12, n_cat=4, seed=1, n_models=1)
    fn = api_b200.new_function("t", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(api_b200.OneHotEncoder(mapping={k: list(v) for k, v in wl.onehot_mapping.items()}), name="onehot").to(
        api_b200.FeatureRowModelServer(name="linear", model=wl.sklearn_models()[0]), name="linear")
    server = fn.to_mock_server()
    rows = wl.rows_as_dicts()
    resp = server.run_events(rows)
    has_nan = np.isnan(wl.X[:, :12]).any(axis=1)
    assert has_nan.any() and not has_nan.all()
    for bad, r in zip(has_nan, resp):
        assert (getattr(r, "status_code", 200) == 400) == bool(bad)
    # the reference gives the same event a 400 too (scikit-learn rejects the NaN inside predict)
    oserver_fn = api_oracle.new_function("t", kind="serving")
    g = oserver_fn.set_topology("flow", engine="sync")
    g.to(api_oracle.OneHotEncoder(mapping={k: list(v) for k, v in wl.onehot_mapping.items()}), name="onehot").to(
        api_oracle.FeatureRowModelServer(name="linear", model=wl.sklearn_models()[0]), name="linear")
    oserver = oserver_fn.to_mock_server()
    i = int(np.argmax(has_nan))
    assert oserver.test(body=dict(rows[i]), silent=True).status_code == 400


def test_router_of_tree_models_run_batch_and_single_route():
    wl = tree_workload(n_rows=1024, n_feat=24, n_models=4, n_trees=10, depth=4, seed=6, n_fit=800)
    server = wl.build_server(api_b200)
    ref = obatch.tree_ensemble(wl)
    out = server.run_batch(wl.X)
    np.testing.assert_allclose(out[:, 0], ref["out"], rtol=RTOL, atol=ATOL)
    one = server.test("/v2/models/m2/infer", body={"inputs": wl.X[:8].astype(np.float64).tolist()})
    np.testing.assert_allclose(one["outputs"], ref["per_model"][:8, 1], rtol=RTOL, atol=ATOL)
    assert one["model_name"] == "m2"
    ens = server.test("/v2/models/infer", body={"inputs": wl.X[:8].astype(np.float64).tolist()})
    np.testing.assert_allclose(ens["outputs"], ref["out"][:8], rtol=RTOL, atol=ATOL)
    assert ens["model_name"] == "VotingEnsemble" and ens["model_version"] == "v1"


def test_unlowerable_graph_is_a_hard_error():
    fn = api_b200.new_function("t", kind="serving")
    graph = fn.set_topology("flow", engine="sync")
    graph.to(name="py", handler="(event)").to(api_b200.Imputer(default_value=0.0), name="imp")
    server = fn.to_mock_server()
    with pytest.raises(LoweringError):
        server.run_batch(np.zeros((4, 3), dtype=np.float32), names=["a", "b", "c"])


def test_run_json_answers_like_the_reference_wire_path():
    """bytes in -> bytes out through the C body codec + one fused launch, against the oracle's GraphServer.run on the
    same JSON body (json.loads -> VotingEnsemble.do_event -> json.dumps)"""
    from mlrun_b200.serving import codec

    for kind in ("regression", "classification"):
        wl = tree_workload(n_rows=300, n_feat=24, n_models=4, n_trees=10, depth=4, seed=7, n_fit=800, kind=kind)
        body = json.dumps({"inputs": wl.X.astype(np.float64).tolist()}).encode()
        server = wl.build_server(api_b200)
        got = server.run_json(body, event_id="evt-1")
        oserver = wl.build_server(api_oracle)
        want = oserver.run(api_oracle.MockEvent(body=body, path="/v2/models/infer", event_id="evt-1", content_type="application/json"))
        assert got.status_code == want.status_code == 200 and got.content_type == want.content_type
        g, w = json.loads(got.body), json.loads(want.body)
        assert list(g) == list(w) and g["id"] == w["id"] == "evt-1" and g["model_name"] == w["model_name"]
        assert g["model_version"] == w["model_version"]
        if kind == "classification":
            assert g["outputs"] == w["outputs"] and all(isinstance(v, int) for v in g["outputs"])
        else:
            np.testing.assert_allclose(g["outputs"], w["outputs"], rtol=RTOL, atol=ATOL)
            # the text itself is what json.dumps prints for those float32 votes
            assert got.body == json.dumps({**g}).encode()
    bad = wl.X.astype(np.float64).tolist()
    bad[3][5] = float("nan")
    resp = server.run_json(json.dumps({"inputs": bad}))
    assert resp.status_code == 400 and "NaN" in resp.body
    with pytest.raises(codec.NotV2Matrix):
        server.run_json(json.dumps({"inputs": [{"f0": 1.0}]}))


def test_tracked_batches_emit_the_per_event_records():
    """model tracking on the batched path: run_events / run_json push the same stream records the
    per-event path pushes (sampling + micro-batching included)"""
    wl = tree_workload(n_rows=64, n_feat=24, n_models=4, n_trees=10, depth=4, seed=8, n_fit=600)

    def tracked_server(api, **params):
        fn = api.new_function("trk", kind="serving")
        graph = fn.set_topology("router", api.VotingEnsemble(vote_type="regression", executor_type="array"))
        for i, m in enumerate(wl.models):
            graph.add_route(f"m{i + 1}", class_name="SKLearnModelServer", model=m, model_path="")
        fn.set_tracking("dummy://")
        fn.spec.parameters.update(params)
        return fn.to_mock_server(namespace={"SKLearnModelServer": api.SKLearnModelServer})

    rows = wl.X.astype(np.float64).tolist()
    # the reference path: one event per row through the oracle server; keep the ensemble's own records
    oserver = tracked_server(api_oracle, log_stream_sample=4)
    for i, row in enumerate(rows):
        oserver.test("/v2/models/infer", body={"inputs": [row]}, event_id=f"e{i}")
    want = [r for r in oserver.context.stream.output_stream.event_list if r["model"] == "VotingEnsemble"]
    assert len(want) == 16
    # the engine: ONE fused launch for the 64 rows as a V2 body, tracked as the single request it is
    bserver = tracked_server(api_b200)
    resp = bserver.run_json(json.dumps({"inputs": rows}), event_id="batch-1")
    recs = [r for r in bserver.context.stream.output_stream.event_list if r["model"] == "VotingEnsemble"]
    assert len(recs) == 1 and recs[0]["class"] == want[0]["class"] and recs[0]["op"] == want[0]["op"]
    assert recs[0]["request"]["id"] == "batch-1" and len(recs[0]["request"]["inputs"]) == 64
    np.testing.assert_allclose(recs[0]["resp"]["outputs"], json.loads(resp.body)["outputs"])
    np.testing.assert_allclose([recs[0]["resp"]["outputs"][3 + 4 * j] for j in range(16)],
                               [r["resp"]["outputs"][0] for r in want], rtol=RTOL, atol=ATOL)
