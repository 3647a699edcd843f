"""
Multi-GPU plumbing for the path: one process per GPU, the query batch split by contiguous index
range, no collective on the data path, a host-side gather of the output blocks to rank 0
(SURVEY.md section 8(e)).  The reference has nothing to mirror here -- it is a single-process NumPy
function -- so this file only fixes the partitioning rule and the gather.

`torch.distributed` is plumbing: NCCL is not needed by the path (weights are replicated, queries are
independent); the gather runs over a gloo (host) process group so that outputs go GPU -> pinned host
-> rank 0 without an extra device hop.
"""

from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Queries [start, stop) owned by `rank`: contiguous blocks of ceil(n / world_size)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of {world_size}")
    per = -(-n // world_size)
    start = min(n, rank * per)
    return start, min(n, start + per)


def shard_rows(raw_rows, n: int, rank: int, world_size: int):
    """Slice 23 raw rows (each of length n or 1) down to this rank's range; length-1 rows broadcast."""
    start, stop = shard_range(n, rank, world_size)
    return [r if np.shape(r)[0] == 1 else r[start:stop] for r in raw_rows], stop - start


def gather_to_rank0(local_block: np.ndarray, n: int, group=None):
    """
    Collect every rank's (rows, n_local) output block on rank 0 into one (rows, n) array, in query
    order.  Works on any initialised torch.distributed backend that supports CPU tensors (gloo).
    Returns the full array on rank 0 and None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-n // world)
    rows = local_block.shape[0]
    start, stop = shard_range(n, rank, world)
    if local_block.shape[1] != stop - start:
        raise ValueError(f"rank {rank} holds {local_block.shape[1]} queries, expected {stop - start}")
    padded = np.zeros((rows, per), dtype=local_block.dtype)
    padded[:, : stop - start] = local_block
    mine = torch.from_numpy(padded)
    pieces = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, pieces, dst=0, group=group)
    if rank != 0:
        return None
    full = np.empty((rows, n), dtype=local_block.dtype)
    for r, piece in enumerate(pieces):
        s, e = shard_range(n, r, world)
        full[:, s:e] = piece.numpy()[:, : e - s]
    return full
