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
