"""Multi-GPU task sharding: one process per GPU, tasks (motif x orientation) are independent.

The reference parallelises over the same unit with joblib processes, each a fresh
`python -m mepp.single` subprocess re-reading the staged dataset from disk
(mepp/batch.py:134-140, mepp/single.py:395-401).  Here every rank stages its own replica of
the packed sequences and the score operand in HBM, takes a static LPT share of the tasks, and
the small per-task result blocks are gathered to rank 0 with one collective
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).  No collective touches
the data path.
"""
import os

import numpy as np


def task_cost(m, n_filters):
    """Relative scan cost of a task; the GEMM / statistics cost is the same for every task."""
    return 8.0 + float(m) * float(n_filters)


def shard_tasks(costs, world_size):
    """Longest-processing-time-first static partition.  Returns a list (per rank) of task index lists,
    each sorted ascending so that a rank's results keep the global task order."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns
    (rank, world_size, local_rank); a single-process run needs no process group."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def gather_blocks(local_block, local_ids, n_total, device=None):
    """Gather per-task result blocks to every rank (rank 0 uses them).

    local_block: float64 array (T_local, ...) of this rank's tasks; local_ids: their global task indices.
    Returns a (n_total, ...) array ordered by global task index.  One all_gather of a padded tensor."""
    import torch
    import torch.distributed as dist
    local_block = np.ascontiguousarray(local_block, dtype=np.float64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.empty((n_total,) + local_block.shape[1:], dtype=np.float64)
        out[np.asarray(local_ids, dtype=np.int64)] = local_block
        return out
    world = dist.get_world_size()
    t_max = (n_total + world - 1) // world + 1
    item = int(np.prod(local_block.shape[1:])) if local_block.ndim > 1 else 1
    pad = np.zeros((t_max, item + 1), dtype=np.float64)
    pad[:, 0] = -1.0
    k = len(local_ids)
    if k > t_max:
        raise ValueError("shard larger than the padded gather block")
    pad[:k, 0] = np.asarray(local_ids, dtype=np.float64)
    pad[:k, 1:] = local_block.reshape(k, item)
    send = torch.from_numpy(pad)
    if device is not None:
        send = send.to(device)
    recv = torch.empty((world * t_max, item + 1), dtype=torch.float64, device=send.device)
    try:
        dist.all_gather_into_tensor(recv, send)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send)
        recv = torch.cat(parts, dim=0)
    rec = recv.cpu().numpy()
    ids = rec[:, 0]
    keep = ids >= 0
    out = np.empty((n_total, item), dtype=np.float64)
    out[ids[keep].astype(np.int64)] = rec[keep, 1:]
    return out.reshape((n_total,) + local_block.shape[1:])


def barrier_and_max(seconds, device=None):
    """Max over ranks of a per-rank duration (the contract's timing rule)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
