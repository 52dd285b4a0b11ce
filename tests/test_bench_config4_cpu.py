"""bench.py's configs[3] leg runs as a child process per rank with its own rendezvous (bench.config4_leg): the orchestration --
environment handed to the children, port shift, summary row, timeout -- checked here with a gloo stand-in for the GPU child."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = os.path.join(ROOT, "tests", "_config4_child.py")


def _free_port_pair(shift=17):
    """a port p with p and p + shift both free (the parents' rendezvous and the children's)"""
    import socket

    for p in range(29533, 29933, 7):
        try:
            for q in (p, p + shift):
                with socket.socket() as sock:
                    sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                    sock.bind(("127.0.0.1", q))
            return p
        except OSError:
            continue
    return 29533


def test_children_of_a_torchrun_job_form_their_own_group():
    env = dict(os.environ, B2S_BENCH_CONFIG4_CMD=json.dumps([CHILD]))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(_free_port_pair()), os.path.join(ROOT, "tests", "_config4_parent.py")]
    done = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert done.returncode == 0, done.stderr[-2000:]
    row = json.loads([ln for ln in done.stdout.splitlines() if ln.startswith("ROW ")][-1][4:])
    assert row["n_gpus"] == 2 and row["events_per_s"] == 2000.0 and row["merge_verified"] is True  # the children's all_reduce ran
    assert row["global_batch"] == 65536 and row["batch_per_gpu"] == 32768 and row["scaling"] == "strong"
    assert abs(row["ms_per_launch"] - 0.5) < 1e-12 and row["launches_timed"] == 80 and row["roofline_frac"] == 0.5


def test_single_process_row_and_a_child_that_hangs(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    monkeypatch.setenv("B2S_BENCH_CONFIG4_CMD", json.dumps([CHILD]))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    row = bench.config4_leg(0, 1)
    assert row["n_gpus"] == 1 and row["events_per_s"] == 1000.0 and row["merge_verified"] is None
    assert row["e2e"]["value"] == 7.0 and "other" not in row["e2e"]  # the child's end-to-end figure travels with the row
    monkeypatch.setenv("B2S_BENCH_CONFIG4_CMD", json.dumps([CHILD, "sleep"]))
    row = bench.config4_leg(0, 1, timeout_s=2.0)
    assert "error" in row and "killed" in row["error"]
    assert bench.config4_leg(1, 2, timeout_s=2.0) is None  # other ranks report nothing
    monkeypatch.setenv("B2S_BENCH_CONFIG4_CMD", json.dumps(["-c", "import sys; sys.exit(3)"]))
    row = bench.config4_leg(0, 1)
    assert "error" in row and row["rc"] == 3
