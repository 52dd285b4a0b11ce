"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol of
include/b200serve.h; the product never imports the oracle; device work fails loudly without a GPU;
event sharding + ensemble-merge all-gather over world_size-2 gloo."""

import ast
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from mlrun_b200 import _native as nat

    header = open(os.path.join(ROOT, "include", "b200serve.h")).read()
    declared = set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", header))
    declared -= {"b2s_status", "b2s_stats", "b2s_devinfo", "b2s_plan_t", "b2s_plan_s"}
    lib = ctypes.CDLL(nat.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f"libb200serve.so does not export {missing}"
    assert declared == set(nat.SIGNATURES), (declared ^ set(nat.SIGNATURES))
    assert nat.load().b2s_version() == 100


def test_plan_building_needs_no_gpu_but_execution_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mlrun_b200 import _native as nat
    from mlrun_b200.plan import DevicePlan

    plan = DevicePlan(8)
    plan.set_impute({0: 1.5})
    plan.add_linear(np.ones((1, 8)), np.zeros(1))
    with pytest.raises(nat.NativeError, match="no CUDA device|CPU fallback|cuda"):
        plan.finalize()


def test_product_never_imports_the_oracle():
    bad = []
    for base, _dirs, files in os.walk(os.path.join(ROOT, "mlrun_b200")):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(base, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.module:
                    names = [node.module]
                if any(n == "oracle" or n.startswith("oracle.") or n.startswith("tests") for n in names):
                    bad.append((f, names))
    assert not bad, bad


def test_device_model_server_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sklearn.linear_model import LinearRegression

    from mlrun_b200 import _native as nat
    from tests import api_b200

    m = LinearRegression().fit(np.random.default_rng(0).normal(size=(20, 4)), np.arange(20.0))
    fn = api_b200.new_function("t", kind="serving")
    fn.set_topology("router")
    fn.add_model("m1", ".", class_name=api_b200.SKLearnModelServer(model=m))
    server = fn.to_mock_server()
    resp = server.test("/v2/models/m1/infer", body={"inputs": [[1.0, 2.0, 3.0, 4.0]]}, silent=True)
    assert resp.status_code == 400 and "NativeError" in resp.body


def test_shard_bounds_cover_and_balance():
    from mlrun_b200.sharding import shard_bounds

    for n in (0, 1, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from mlrun_b200.sharding import shard_bounds, gather_votes, torch_exchange
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 1001
votes = torch.arange(n, dtype=torch.float32).reshape(n, 1) * 0.5   # what the full batch would produce
lo, hi = shard_bounds(n, rank, world)
full = gather_votes(dist, votes[lo:hi].clone(), n, world)
assert full.shape == (n, 1) and torch.equal(full, votes), (rank, full.shape)
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the bench's max-over-ranks timing reduction
assert t.item() == world
# the bootstrap channel of the merge communicator (MergeComm / b2s_comm_connect): 64 bytes per rank, in rank order
blobs = torch_exchange(dist)(bytes([rank]) * 64)
assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == 64 for b in blobs)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_event_sharding_all_gather_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29611", str(script)],
        capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_large_results_come_from_the_pinned_pool_and_small_ones_do_not(monkeypatch):
    """DevicePlan._result_arrays: layout of the (outputs, status) pair over one pool block; plain arrays below 1 MiB or when
    the pool is exhausted (host logic; the pool itself needs a GPU, so a stand-in hands out ordinary memory here)"""
    import ctypes

    import numpy as np

    from mlrun_b200 import _native as nat
    from mlrun_b200.plan import DevicePlan

    taken = []

    class Pool:
        def take(self, nbytes):
            taken.append(nbytes)
            return None if nbytes > (64 << 20) else (ctypes.c_char * nbytes)()

    monkeypatch.setattr(nat, "PINNED", Pool())
    plan = DevicePlan.__new__(DevicePlan)
    plan.out_cols, plan.out_is_int = 3, False
    out, st = plan._result_arrays(1000)
    assert not taken and out.shape == (1000, 3) and out.dtype == np.float32 and st.shape == (1000,) and st.dtype == np.int32
    out, st = plan._result_arrays(100_001)
    assert taken == [(100_001 * 12 + 63) // 64 * 64 + 100_001 * 4]
    assert out.shape == (100_001, 3) and out.flags["C_CONTIGUOUS"] and st.shape == (100_001,)
    out[:] = 1.5
    st[:] = 7
    assert (out == 1.5).all() and (st == 7).all()  # the two arrays do not overlap
    plan.out_is_int = True
    labels, _ = plan._result_arrays(400_000)
    assert labels.dtype == np.int32
    big_out, big_st = plan._result_arrays(10_000_000)  # the pool says no: ordinary arrays
    assert big_out.shape == (10_000_000, 3) and big_out.base is None


def test_array_addresses_handed_to_the_library():
    """_native.ptr (the cheap address of an array's first element that plan.run / submit / wait pass to the C-ABI) equals
    arr.ctypes.data for every kind of array the host layer builds: fresh, sliced, pinned-style (ctypes-backed), read-only,
    empty and strided ones (the last three through its fallback)"""
    import ctypes as C

    import numpy as np

    from mlrun_b200 import _native as nat

    raw = C.create_string_buffer(1 << 16)
    backed = np.frombuffer((C.c_char * (1 << 16)).from_address(C.addressof(raw)), dtype=np.float32, count=4096).reshape(64, 64)
    status = np.frombuffer((C.c_char * (1 << 16)).from_address(C.addressof(raw)), dtype=np.int32, count=64, offset=32768)
    cases = [np.zeros((4096, 64), np.float32), np.zeros((10, 64), np.float32)[3:], backed, status, np.frombuffer(b"12345678", dtype=np.float32),
             np.empty((0, 64), np.float32), np.zeros((8, 128), np.float32)[:, :64], np.zeros(7, np.int32), np.float32(1.0).reshape(1, 1)]
    for a in cases:
        assert nat.ptr(a) == a.ctypes.data, (a.shape, a.flags)
    assert nat.ptr(backed) == C.addressof(raw) and nat.ptr(status) == C.addressof(raw) + 32768


def test_reference_import_paths_resolve():
    """existing code imports the serving classes by the reference's module paths; they resolve to this package's classes"""
    import mlrun_b200 as mlrun
    from mlrun_b200.serving.routers import ModelRouter, ParallelRun, VotingEnsemble
    from mlrun_b200.serving.server import GraphContext, GraphServer, MockEvent, create_graph_server
    from mlrun_b200.serving.states import ErrorStep, FlowStep, QueueStep, RouterStep, TaskStep
    from mlrun_b200.serving.v2_serving import V2ModelServer

    assert mlrun.serving.VotingEnsemble is VotingEnsemble and mlrun.serving.V2ModelServer is V2ModelServer
    assert mlrun.serving.routers.ModelRouter is ModelRouter and mlrun.serving.server.MockEvent is MockEvent
    assert mlrun.ServingRuntime is mlrun.serving.ServingRuntime and type(mlrun.new_function("f", kind="serving")) is mlrun.ServingRuntime
    assert all(isinstance(c, type) for c in (ParallelRun, GraphContext, GraphServer, ErrorStep, FlowStep, QueueStep, RouterStep, TaskStep))
    assert callable(create_graph_server)
