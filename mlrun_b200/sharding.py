"""Event sharding across GPUs (one process per GPU).

Events are independent through every op on the path (no cross-row state: the vote reduces across
*models*, mlrun/serving/routers.py:797-810), so a batch is split by rows and every rank runs the same
plan on its shard; the only exchange is the ensemble-merge: an all-gather of each shard's (rows, out_cols)
votes so that every rank -- and the host that answers the request -- holds the whole response
(4 bytes / event).  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""


def shard_bounds(n_rows, rank, world):
    """contiguous, balanced row ranges: the first (n_rows % world) ranks get one extra row"""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_votes(dist, local, n_rows, world):
    """all-gather ragged shards of votes -> the full (n_rows, out_cols) tensor on every rank.
    `local` is this rank's (rows_r, out_cols) tensor; shards are padded to the largest shard so that a
    single all_gather_into_tensor serves every rank."""
    import torch

    max_rows = shard_bounds(n_rows, 0, world)[1]
    cols = local.shape[1]
    padded = torch.zeros((max_rows, cols), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    full = torch.empty((world * max_rows, cols), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, padded)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_rows, r, world)
        parts.append(full[r * max_rows: r * max_rows + (hi - lo)])
    return torch.cat(parts, dim=0)


class MergeComm:
    """The ensemble-merge communicator of the C-ABI (include/b200serve.h, b2s_comm_*): one device allocation per rank --
    completion flags + the merged response rows in four slots -- mapped by every peer over CUDA IPC.

    `exchange(blob) -> [blob of rank 0, ..., blob of rank world-1]` is the only thing the bootstrap needs from the outside
    (64 bytes per rank): `torch.distributed.all_gather_object`, MPI, a shared file ...; `torch_exchange(dist)` wraps the
    first."""

    def __init__(self, rank, world, max_rows_per_rank, out_cols, exchange):
        import ctypes as C

        from . import _native as nat

        self._nat, self._lib = nat, nat.init()
        self.rank, self.world, self.out_cols = int(rank), int(world), int(out_cols)
        self.max_rows = (int(max_rows_per_rank) + 3) // 4 * 4
        self._h = C.c_void_p()
        nat.check(self._lib.b2s_comm_create(self.rank, self.world, int(max_rows_per_rank), self.out_cols, C.byref(self._h)))
        if self.world > 1:
            mine = C.create_string_buffer(64)
            nat.check(self._lib.b2s_comm_handle(self._h, mine))
            blobs = exchange(bytes(mine.raw))
            if len(blobs) != self.world or any(len(b) != 64 for b in blobs):
                raise ValueError("exchange() must return one 64-byte handle per rank, in rank order")
            nat.check(self._lib.b2s_comm_connect(self._h, C.create_string_buffer(b"".join(blobs), 64 * self.world)))

    def set_fused_wait(self, lag):
        """lag 0 / 1: every launch ends by waiting (in its own last CTA) for its own / the previous step's flags -- `wait`
        then enqueues nothing for a covered step; None: off (a one-warp wait kernel per `wait`)"""
        self._nat.check(self._lib.b2s_comm_set_fused_wait(self._h, -1 if lag is None else int(lag)))

    def attach(self, plan):
        self._nat.check(self._lib.b2s_plan_attach_comm(plan._h, self._h))
        return plan

    def detach(self, plan):
        self._nat.check(self._lib.b2s_plan_attach_comm(plan._h, None))

    def wait(self, stream=None, lag=0):
        """enqueue the completion wait of the step launched `lag` launches ago (0: the one just launched; 1: the one before,
        so that a step's votes cross NVLink while the next step is scored) -> (device pointer of that step's merged rows,
        epoch); (None, 0) while fewer than lag + 1 steps have been launched"""
        import ctypes as C

        ptr, epoch = C.c_void_p(), C.c_uint32()
        self._nat.check(self._lib.b2s_comm_wait_lag(self._h, stream, int(lag), C.byref(ptr), C.byref(epoch)))
        return ptr.value, epoch.value

    def check(self):
        self._nat.check(self._lib.b2s_comm_check(self._h))

    def close(self):
        if self._h:
            self._lib.b2s_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def torch_exchange(dist, group=None):
    """exchange() over torch.distributed (NCCL on GPUs, gloo in the CPU tests)"""

    def exchange(blob):
        out = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, blob, group=group)
        return out

    return exchange


class ShardedGraphServer:
    """A GraphServer whose batches are event-sharded over `world` GPUs (one process per GPU): BASELINE configs[3], the
    8-replica router with the ensemble-merge.  Every rank holds the same graph (same models, replicated tables), scores its
    own shard with the fused plan, and the plan's kernels store the shard's votes into EVERY rank's response buffer over
    NVLink peer mappings; `run_batch` returns the merged response of all shards once the completion flags of all ranks
    have been observed on the device (no host barrier, no collective).  Replaces the per-event fan-out / reduce of
    ParallelRun._parallel_run and VotingEnsemble._apply_logic (mlrun/serving/routers.py:414-455, 789-810) across replicas.

        server = fn.to_mock_server(...)
        sharded = ShardedGraphServer(server, rank, world, max_rows_per_rank, torch_exchange(dist), names=feature_names)
        merged = sharded.run_batch(X[lo:hi])        # (world * max_rows_per_rank, out_cols); rank r's rows at r * max_rows

    Every rank must call run_batch the same number of times (a step is collective in the sense that its flags are awaited)."""

    def __init__(self, server, rank, world, max_rows_per_rank, exchange, names=None, fused_wait=0):
        """fused_wait: 0 (default) -- every launch waits for its own step's flags in its last CTA (`run_batch` / `run_device`
        with lag 0 then need no wait kernel); 1 -- for the previous step's (pipelined callers: `run_device(lag=1)`); None -- off"""
        from . import _native as nat

        self.server, self.rank, self.world = server, int(rank), int(world)
        self.plan = server.compile(names).plan
        self.comm = MergeComm(rank, world, max_rows_per_rank, self.plan.out_cols, exchange)
        self.max_rows = self.comm.max_rows
        self.comm.set_fused_wait(fused_wait)
        self.comm.attach(self.plan)
        self._nat = nat
        self._d_in = nat.DeviceBuffer(self.max_rows * self.plan.n_in * 4)

    def rows_of(self, merged, rank, n_rows):
        """rank's shard inside a merged response"""
        return merged[rank * self.max_rows: rank * self.max_rows + n_rows]

    def run_device(self, d_rows, n_rows, row_stride=None, stream=None, lag=0):
        """device-resident shard -> (device pointer of the merged rows, epoch); asynchronous on `stream`.
        lag=0: the merged response of THIS step.  lag=1 (pipelined serving): the merged response of the PREVIOUS step
        ((None, 0) on the first call) -- this step's votes travel while the next one is scored; `drain()` returns the last."""
        self.plan.run_device(d_rows, n_rows, row_stride or self.plan.n_in * 4, None, None, stream)
        return self.comm.wait(stream, lag)

    def drain(self, stream=None):
        """after lag=1 steps: the merged response of the last step launched"""
        return self.comm.wait(stream, 0)

    def run_batch(self, X):
        """this rank's shard (B_r, F) float32 -> the merged (world * max_rows, out_cols) response of all ranks' shards"""
        import numpy as np

        nat = self._nat
        X = np.ascontiguousarray(X, dtype=np.float32)
        if X.ndim != 2 or X.shape[1] != self.plan.n_in or len(X) > self.max_rows:
            raise ValueError(f"a shard is a float32 (<= {self.max_rows}, {self.plan.n_in}) array")
        self._d_in.upload(X)
        ptr, _epoch = self.run_device(self._d_in.ptr, len(X))
        out = np.empty((self.world * self.max_rows, self.plan.out_cols), dtype=self.plan.out_dtype)
        nat.check(nat.load().b2s_device_sync())  # the wait kernel has seen every rank's flag
        self.comm.check()
        nat.check(nat.load().b2s_memcpy_d2h(out.ctypes.data, ptr, out.nbytes))
        return out

    def close(self):
        self.comm.detach(self.plan)
        self.comm.close()
