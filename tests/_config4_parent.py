"""one rank of the parent side of bench.config4_leg under torchrun (tests/test_bench_config4_cpu.py)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")  # the parent group stays up while the children run theirs, as in bench.py
row = bench.config4_leg(rank, world, timeout_s=120.0)
dist.barrier()
if rank == 0:
    print("ROW " + json.dumps(row))
else:
    assert row is None
dist.destroy_process_group()
