"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: LPT sharding and the result gather."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_tasks_is_a_balanced_partition():
    from mepp_b200 import parallel
    rng = np.random.default_rng(0)
    costs = [parallel.task_cost(int(m), int(f)) for m, f in zip(rng.integers(5, 26, 324), rng.integers(1, 3, 324))]
    for world in (1, 2, 4, 8):
        shards = parallel.shard_tasks(costs, world)
        allidx = sorted(i for s in shards for i in s)
        assert allidx == list(range(324))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)
        assert all(s == sorted(s) for s in shards)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from mepp_b200 import parallel
    import torch.distributed as dist
    r, w, _ = parallel.init_from_env(backend="gloo")
    n_total = 11
    costs = [parallel.task_cost(5 + i, 1 + i % 2) for i in range(n_total)]
    mine = parallel.shard_tasks(costs, w)[r]
    block = np.stack([np.full((3, 2), float(i)) + np.arange(6).reshape(3, 2) / 10 for i in mine])
    full = parallel.gather_blocks(block.astype(np.float32), mine, n_total)
    assert (full is None) == (r != 0)
    both = parallel.gather_blocks(block, mine, n_total, root_only=False)
    assert both.shape == (n_total, 3, 2) and both.dtype == np.float64
    # repeated calls reuse the gathered task ids (one collective less); the data are new every time
    again = parallel.gather_blocks((block * 2).astype(np.float32), mine, n_total)
    if r == 0:
        assert np.allclose(again, 2 * full)
    again_all = parallel.gather_blocks(block * 3, mine, n_total, root_only=False)
    assert np.allclose(again_all, 3 * both)
    tmax = parallel.barrier_and_max(1.0 + r)
    if r == 0:
        np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "tmax.npy"), np.array([tmax]))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_blocks_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = np.load(tmp_path / "full.npy")
    assert full.shape == (11, 3, 2) and full.dtype == np.float32
    for i in range(11):
        assert np.allclose(full[i], float(i) + np.arange(6).reshape(3, 2) / 10, atol=1e-6)
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0


def test_gather_blocks_single_process():
    from mepp_b200 import parallel
    blk = np.arange(12, dtype=float).reshape(3, 4)
    out = parallel.gather_blocks(blk, [2, 0, 1], 3)
    assert np.array_equal(out[2], blk[0]) and np.array_equal(out[0], blk[1])


def test_shard_task_units_keeps_orientation_pairs_together():
    import numpy as np
    from mepp_b200 import parallel
    rng = np.random.default_rng(0)
    motifs = [rng.random((4, m)) for m in (6, 8, 8, 12, 20, 9, 15)]
    tasks = [(pwm, o) for pwm in motifs for o in ('+', '+/-')]
    for world in (1, 2, 3):
        shards = parallel.shard_task_units(tasks, world)
        assert sorted(i for s in shards for i in s) == list(range(len(tasks)))
        for s in shards:
            assert all((i ^ 1) in s for i in s)            # task 2k and 2k+1 (same motif) on the same rank
