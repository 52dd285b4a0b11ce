"""2-GPU lab: the sharded flow3 check of tests/test_gpu_multi.py, reporting every mismatching (step, source rank) instead of
stopping at the first; run under different B2S_RT_* switches to bisect."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mlrun_b200 import _native as nat, api
from mlrun_b200.sharding import ShardedGraphServer, shard_bounds, torch_exchange
from mlrun_b200.synthetic import flow3_workload
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")
nat.init(rank)
wl = flow3_workload(n_rows=10001, n_num=56, n_cat=8, seed=7, n_models=4)
server = wl.build_server(api, engine="sync")
rng = np.random.default_rng(5)
batches = [wl.X, wl.X[rng.permutation(len(wl.X))[:7777]]]
full = [server.run_batch(X, names=wl.names) for X in batches]
again = [server.run_batch(X, names=wl.names) for X in batches]
print("rank", rank, "single-GPU repeatable:", [bool(np.array_equal(a, b)) for a, b in zip(full, again)], server.compile(wl.names).plan.kernel[:50], flush=True)
max_rows = max(shard_bounds(len(X), 0, world)[1] for X in batches)
sharded = ShardedGraphServer(server, rank, world, max_rows, torch_exchange(dist), names=wl.names)
bad = 0
for step, X in enumerate(batches * 6):
    lo, hi = shard_bounds(len(X), rank, world)
    merged = sharded.run_batch(X[lo:hi])
    for r in range(world):
        rlo, rhi = shard_bounds(len(X), r, world)
        got = sharded.rows_of(merged, r, rhi - rlo)
        want = full[step % 2][rlo:rhi]
        ne = np.flatnonzero((got != want).any(axis=1))
        if len(ne):
            bad += 1
            print(f"rank {rank} step {step} (epoch {step + 1}, slot {(step + 1) & 3}) rows of rank {r}: {len(ne)} of {rhi - rlo} differ, first {ne[:6]}, last {ne[-3:]}, got {got[ne[0]]} want {want[ne[0]]}", flush=True)
print("rank", rank, "mismatching (step, source) pairs:", bad, flush=True)
sharded.close()
dist.barrier()
dist.destroy_process_group()
