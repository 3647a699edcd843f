"""
Multi-GPU plumbing for the path: the query batch is split by contiguous index range, the
weights are replicated, there is NO collective on the data path, and the outputs are gathered on the host.
The reference has nothing to mirror here -- it is a single-process NumPy function (main.py:198-201 makes every row of
the stacked input an independent case, which is what makes the split trivial).

Two deployments share the one partitioning rule (`shard_range` == the C ABI's `nfb_shard_range`):

  * one process, several GPUs: `Evaluator(devices="all")` -> `nfb_multi_eval_host`; every device copies its results
    device->host straight into its column range of the caller's block, so the gather is free.
  * one process per GPU (torchrun): every rank evaluates its range of ONE batch and lands it in rank 0's block:
      - `SharedHostBlock`: a POSIX shared-memory block that every rank of the box maps and page-locks
        (`nfb_host_register`).  Rank r passes its column range of that block as `out` to `evaluate_host_into`; the
        device->host copies write rank 0's memory directly -- again no gather step, only a barrier.
      - `gather_to_rank0`: the portable fallback over a gloo (host) process group, for ranks that do not share a box.
"""

from __future__ import annotations

import ctypes as C
from multiprocessing import shared_memory

import numpy as np


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """
    Queries [start, stop) owned by `rank`: contiguous blocks of ceil(n / world_size) rounded up to a multiple of 4
    queries (so that every shard's rows stay 16-byte aligned in a float32 block) -- the rule of nfb_shard_range.
    """
    if world_size < 1 or not (0 <= rank < world_size) or n < 0:
        raise ValueError(f"bad rank {rank} of {world_size}")
    per = (-(-n // world_size) + 3) & ~3
    return min(n, rank * per), min(n, (rank + 1) * per)


def shard_rows(raw_rows, n: int, rank: int, world_size: int):
    """Slice 23 raw rows (each of length n or 1) down to this rank's range; length-1 rows broadcast."""
    start, stop = shard_range(n, rank, world_size)
    return [r if np.shape(r)[0] == 1 else r[start:stop] for r in raw_rows], stop - start


class SharedHostBlock:
    """
    A (rows, ld) array in POSIX shared memory (/dev/shm), page-locked for CUDA in every process that maps it.

    Rank 0 creates it (`create=True`) and tells the others its `name` (e.g. `dist.broadcast_object_list`); they attach
    with the same shape and dtype.  `array` is the NumPy view; `shard(rank, world, n)` is the column range rank `rank`
    owns.  Without a GPU (the gloo tests) `pin=False` leaves the block pageable -- the partition and the hand-over are
    the same.  Call `close()` everywhere, then `unlink()` on rank 0.
    """

    def __init__(self, rows: int, ld: int, dtype=np.float32, name: str | None = None, create: bool = False, pin: bool = True):
        self.rows, self.ld, self.dtype = int(rows), int(ld), np.dtype(dtype)
        nbytes = max(1, self.rows * self.ld * self.dtype.itemsize)
        if create:  # a tmpfs hands out sparse files: without this check a block that does not fit dies later with SIGBUS
            import os

            try:
                st = os.statvfs("/dev/shm")
                free = st.f_bavail * st.f_frsize
            except OSError:
                free = None
            if free is not None and free < nbytes + (64 << 20):
                raise MemoryError(f"/dev/shm has {free >> 20} MiB free; a shared block of {nbytes >> 20} MiB does not fit")
        self._shm = shared_memory.SharedMemory(name=name, create=create, size=nbytes)
        self.name = self._shm.name
        self.array = np.ndarray((self.rows, self.ld), dtype=self.dtype, buffer=self._shm.buf)
        self._pinned = False
        self._lib = None
        if pin:
            from . import _lib

            self._lib = _lib.load()
            _lib.check(self._lib.nfb_host_register(C.c_void_p(self.array.ctypes.data), nbytes))
            self._pinned = True

    def shard(self, rank: int, world_size: int, n: int) -> np.ndarray:
        start, stop = shard_range(n, rank, world_size)
        return self.array[:, start:stop]

    def close(self) -> None:
        if self._pinned:
            self._lib.nfb_host_unregister(C.c_void_p(self.array.ctypes.data))
            self._pinned = False
        self.array = None
        self._shm.close()

    def unlink(self) -> None:
        self._shm.unlink()


def gather_to_rank0(local_block: np.ndarray, n: int, group=None):
    """
    Collect every rank's (rows, n_local) output block on rank 0 into one (rows, n) array, in query
    order.  Works on any initialised torch.distributed backend that supports CPU tensors (gloo).
    Returns the full array on rank 0 and None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (-(-n // world) + 3) & ~3
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
