"""
The N > 1 path on the CPU: two gloo ranks shard a batch by contiguous range, evaluate their ranges
independently (here with the oracle standing in for the per-GPU evaluator -- the partition and the
gather are what is under test), and rank 0 gathers the blocks in query order.
"""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, out_path: str):
    sys.path.insert(0, str(REPO))
    import torch.distributed as dist

    from neuralfoil_b200 import sharding, synthetic
    from oracle.neuralfoil_oracle import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        raw = synthetic.sample_training_distribution(n, seed=5)
        rows = [raw[i] for i in range(23)]
        rows[20] = np.array([9.0])  # one broadcast row, as the dict API produces for scalars
        local_rows, n_local = sharding.shard_rows(rows, n, rank, world)
        local_raw = np.stack([np.broadcast_to(r, (n_local,)) for r in local_rows])
        block = Oracle().evaluate_raw(local_raw, "xsmall") if n_local else np.zeros((198, 0))
        full = sharding.gather_to_rank0(block, n)
        # the same hand-over through a shared host block (what the torchrun bench does with page-locked memory):
        # rank 0 creates it, the others attach by name and write their own column range; a barrier completes it
        names = [None]
        shared = None
        if rank == 0:
            shared = sharding.SharedHostBlock(198, n + 3, np.float64, create=True, pin=False)
            names[0] = shared.name
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            shared = sharding.SharedHostBlock(198, n + 3, np.float64, name=names[0], pin=False)
        shared.shard(rank, world, n)[...] = block
        dist.barrier()
        if rank == 0:
            assert np.array_equal(shared.array[:, :n], full)
            np.save(out_path, full)
        else:
            assert full is None
        dist.barrier()
        shared.close()
        if rank == 0:
            shared.unlink()
    finally:
        dist.destroy_process_group()


def test_shard_rule_is_the_c_abis(library):
    """sharding.shard_range (Python, used by the ranks of a torchrun job) == nfb_shard_range (used by nfb_multi_eval_host)."""
    import ctypes as C

    from neuralfoil_b200 import sharding

    for n in (0, 1, 3, 4, 5, 101, 1000, 10_000_000, 99_999_999):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for rank in range(world):
                b, e = C.c_int64(), C.c_int64()
                assert library.nfb_shard_range(n, rank, world, C.byref(b), C.byref(e)) == 0
                assert (b.value, e.value) == sharding.shard_range(n, rank, world)
                assert b.value == covered and b.value % 4 == 0 or b.value == n
                covered = e.value
            assert covered == n
    assert library.nfb_shard_range(10, 2, 2, C.byref(b), C.byref(e)) != 0


@pytest.mark.parametrize("world,n", [(2, 101), (2, 1), (3, 10)])
def test_sharded_evaluation_equals_single_process(tmp_path, world, n):
    sys.path.insert(0, str(REPO))
    from neuralfoil_b200 import synthetic
    from oracle.neuralfoil_oracle import Oracle

    out_path = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, out_path), nprocs=world, join=True)
    raw = synthetic.sample_training_distribution(n, seed=5)
    raw[20] = 9.0
    want = Oracle().evaluate_raw(raw, "xsmall")
    # float64 BLAS blocks differently for different batch widths: equal to rounding, not bit for bit
    np.testing.assert_allclose(np.load(out_path), want, rtol=1e-11, atol=1e-300)
