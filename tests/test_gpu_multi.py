"""Multi-GPU (one process per GPU): event sharding + the fused P2P ensemble-merge.  Needs >= 2 GPUs; the
single-GPU `-m gpu` run skips it (the same sharding logic is covered over gloo in test_boundary_cpu.py)."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from mlrun_b200 import _native as nat, api
from mlrun_b200.sharding import shard_bounds
from mlrun_b200.synthetic import flow3_workload
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
nat.init(rank)
wl = flow3_workload(n_rows=10000, n_num=56, n_cat=8, seed=7, n_models=4)     # the same batch on every rank
server = wl.build_server(api, engine="sync")
plan = server.compile(wl.names).plan
full_ref = plan.run(wl.X)                                                      # single-GPU answer
lo, hi = shard_bounds(len(wl.X), rank, world)
merged = nat.DeviceBuffer(len(wl.X) * plan.out_cols * 4)
handles = [None] * world
dist.all_gather_object(handles, nat.ipc_export(merged.ptr))
peers = [merged.ptr if r == rank else nat.ipc_open(handles[r]) for r in range(world)]
plan.set_merge_targets(peers, lo)
shard = nat.DeviceBuffer((hi - lo) * 256).upload(wl.X[lo:hi])
plan.run_device(shard.ptr, hi - lo, 256, None)
nat.load().b2s_device_sync()
dist.barrier()
got = merged.download(np.float32, (len(wl.X), plan.out_cols))
assert np.array_equal(got, full_ref), (rank, np.abs(got - full_ref).max())
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "merge ok")
'''


def test_fused_p2p_merge_two_gpus(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29621", str(script)],
        capture_output=True, text=True, env=dict(os.environ, REPO_ROOT=ROOT), timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("merge ok") == 2
