"""NVLink evidence for the fused ensemble-merge: rank 0 scores 1 Mi events per launch with the merge communicator attached
(its kernel's epilogue stores the votes into every rank's response buffer); the other ranks only hold their buffers open.
Run rank 0 under `ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,...` (profiles/lab/gpu9.sh); nothing is timed here."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mlrun_b200 import _native as nat, api  # noqa: E402
from mlrun_b200.sharding import MergeComm, torch_exchange  # noqa: E402
from mlrun_b200.synthetic import flow3_workload  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")  # the bootstrap only exchanges 64-byte handles
nat.init(rank)
B = 1 << 20
wl = flow3_workload(n_rows=65536, n_num=56, n_cat=8, seed=2, n_models=4)
server = wl.build_server(api, engine="sync")
plan = server.compile(wl.names).plan
comm = MergeComm(rank, world, B, plan.out_cols, torch_exchange(dist))
dist.barrier()
if rank == 0:
    comm.attach(plan)
    X = torch.from_numpy(np.tile(wl.X, (B // 65536, 1))).cuda()
    out = torch.empty(B * plan.out_cols, dtype=torch.float32, device="cuda")
    for _ in range(6):
        plan.run_device(X.data_ptr(), B, X.shape[1] * 4, out.data_ptr(), None, None)
    torch.cuda.synchronize()
    print("rank 0: 6 launches with P2P merge stores to", world, "ranks done")
    comm.detach(plan)
dist.barrier()
comm.close()
dist.destroy_process_group()
