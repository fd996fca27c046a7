"""CPU, world_size 2 over gloo: the multi-GPU path is pure run sharding + one metric gather at the end
(SURVEY.md section 8e), so the host logic can be exercised without GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, num_runs, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import auction_gym_b200 as ag
        from auction_gym_b200 import driver

        first, count = ag.shard_runs(num_runs, world, rank)
        # stand-in for the per-run metric block a rank would read from its engine: value encodes the global run index
        local = np.stack([np.full((3, 4, len(driver.MEASURES)), first + r, np.float64) for r in range(count)]) if count else \
            np.zeros((0, 3, 4, len(driver.MEASURES)))
        rev = np.arange(first, first + count, dtype=np.float64)[:, None] * np.ones((1, 3))
        allm = driver.gather_runs(local, world)
        allr = driver.gather_runs(rev, world)
        np.save(os.path.join(tmp, f"m{rank}.npy"), allm)
        np.save(os.path.join(tmp, f"r{rank}.npy"), allr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_runs", [5, 2, 1])
def test_shard_and_gather_equals_single_process(tmp_path, num_runs):
    world = 2
    port = 29500 + (os.getpid() % 2000) + num_runs
    mp.spawn(_worker, args=(world, port, num_runs, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        m = np.load(tmp_path / f"m{rank}.npy")
        r = np.load(tmp_path / f"r{rank}.npy")
        assert m.shape[0] == num_runs and r.shape == (num_runs, 3)
        assert np.array_equal(m[:, 0, 0, 0], np.arange(num_runs))  # rank order == global run order
        assert np.array_equal(r[:, 0], np.arange(num_runs))
