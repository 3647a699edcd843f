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
        if rank == 0:
            np.save(out_path, full)
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


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
