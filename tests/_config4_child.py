"""stand-in for the configs[3] child of bench.py (tests/test_bench_config4_cpu.py): joins a gloo group from the environment it
was given, proves the group works (all_reduce of the ranks), and prints a bench-style JSON line on rank 0"""
import json
import os
import sys
import time

if len(sys.argv) > 1 and sys.argv[1] == "sleep":
    time.sleep(60)
    sys.exit(0)
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
total = 0
if world > 1:
    import torch
    import torch.distributed as dist

    assert not any(k.startswith("TORCHELASTIC_") for k in os.environ), "the agent-store variables must not reach the child"
    dist.init_process_group("gloo")
    t = torch.tensor([rank + 1])
    dist.all_reduce(t)
    total = int(t.item())
    dist.destroy_process_group()
if rank == 0:
    print("some log line")
    print(json.dumps({"metric": "events/sec", "value": 1000.0 * world, "n_gpus": world, "steps": 10, "ms_per_step": 4.0, "scaling": "strong",
                      "config": {"workload": "stand-in", "global_batch": 65536, "batch_per_gpu": 65536 // world, "launches_per_step": 8,
                                 "merge_verified": total == world * (world + 1) // 2 if world > 1 else None, "kernel": "none",
                                 "parallelism": f"x{world}", "master_port": os.environ.get("MASTER_PORT")},
                      "roofline": {"frac": 0.5}, "clocks": None, "e2e": {"value": 7.0, "unit": "events/s", "batch": 5, "other": 1}, "p50_step_latency_us": {"p50": 1.0}}))
