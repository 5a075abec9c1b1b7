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


def host_threads():
    """Host worker threads this process may use: MEPP_HOST_THREADS, else cores / ranks-on-this-node
    (torchrun exports LOCAL_WORLD_SIZE), so that 8 ranks do not oversubscribe the host 8-fold."""
    env = os.environ.get("MEPP_HOST_THREADS")
    if env:
        return max(1, int(env))
    cores = os.cpu_count() or 1
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, min(32, cores // max(1, local_world)))


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


def shard_task_units(tasks, world_size):
    """shard_tasks over scan units: consecutive tasks with the same motif matrix (the reference's task order is
    motif outer, orientation inner, batch.py:73-90) stay on one rank, so that the '+' orientation can ride on the
    '+/-' pass of its motif (engine.scan_units).  tasks: [(motif_matrix (4, m), orientation)]."""
    units, costs = [], []
    prev = None
    for i, (pwm, o) in enumerate(tasks):
        key = (np.shape(pwm), np.ascontiguousarray(pwm).tobytes())
        c = task_cost(np.shape(pwm)[1], 1 if o in ('+', '-') else 2)
        if key == prev:
            units[-1].append(i)
            costs[-1] = max(costs[-1], c)          # the twin rides on the dual pass
        else:
            units.append([i])
            costs.append(c)
        prev = key
    return [sorted(i for u in s for i in units[u]) for s in shard_tasks(costs, world_size)]


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


def gather_blocks(local_block, local_ids, n_total, device=None, root_only=True):
    """Gather per-task result blocks to rank 0 (every rank when root_only=False).

    local_block: array (T_local, ...) of this rank's tasks (dtype kept: float32 blocks stay float32);
    local_ids: their global task indices.  Returns a (n_total, ...) array ordered by global task index on
    rank 0 (None on the other ranks when root_only).  One collective on a padded tensor; the task ids travel
    in a second, tiny one."""
    import torch
    import torch.distributed as dist
    local_block = np.ascontiguousarray(local_block)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.empty((n_total,) + local_block.shape[1:], dtype=local_block.dtype)
        out[np.asarray(local_ids, dtype=np.int64)] = local_block
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    t_max = (n_total + world - 1) // world + 1
    k = len(local_ids)
    if k > t_max:
        raise ValueError("shard larger than the padded gather block")
    item_shape = local_block.shape[1:]
    ids = np.full(t_max, -1, dtype=np.int64)
    ids[:k] = np.asarray(local_ids, dtype=np.int64)
    # the task ids of a sharding do not change from call to call: their collective runs once per (shard, device)
    ids_key = (ids.tobytes(), n_total, world, str(device), bool(root_only))
    cached_ids = _GATHERED_IDS.get(ids_key)
    if device is not None:
        # padded send block in a cached pinned buffer: one host copy, one asynchronous upload
        skey = ('send', (t_max,) + tuple(item_shape), local_block.dtype.str)
        pin = _PINNED.get(skey)
        if pin is None:
            pin = _PINNED[skey] = torch.zeros((t_max,) + tuple(item_shape), dtype=torch.from_numpy(local_block[:0]).dtype,
                                              pin_memory=True)
        pin.numpy()[:k] = local_block
        send = pin.to(device, non_blocking=True)
        send_ids = torch.from_numpy(ids).to(device, non_blocking=True) if cached_ids is None else None
    else:
        pad = np.zeros((t_max,) + item_shape, dtype=local_block.dtype)
        pad[:k] = local_block
        send = torch.from_numpy(pad)
        send_ids = torch.from_numpy(ids) if cached_ids is None else None
    # one receive buffer (rank-major), the per-rank pieces are views of it; on the root it goes to the host in one copy
    # into a cached pinned buffer
    big = big_ids = None
    done = False
    if root_only:
        try:
            if rank == 0:
                big = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
            dist.gather(send, list(big.unbind(0)) if rank == 0 else None, dst=0)
            if cached_ids is None:
                if rank == 0:
                    big_ids = torch.empty((world, t_max), dtype=send_ids.dtype, device=send.device)
                dist.gather(send_ids, list(big_ids.unbind(0)) if rank == 0 else None, dst=0)
            done = True
        except (RuntimeError, NotImplementedError):
            done = False
    if not done:
        big = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        dist.all_gather(list(big.unbind(0)), send)
        if cached_ids is None:
            big_ids = torch.empty((world, t_max), dtype=send_ids.dtype, device=send.device)
            dist.all_gather(list(big_ids.unbind(0)), send_ids)
    if cached_ids is None:
        # (every rank remembers that the ids travelled; only a receiving rank keeps them)
        _GATHERED_IDS[ids_key] = big_ids.cpu().numpy() if big_ids is not None else True
    if root_only and rank != 0:
        return None
    rid = _GATHERED_IDS[ids_key]
    if big.is_cuda:
        key = (tuple(big.shape), big.dtype)
        host = _PINNED.get(key)
        if host is None:
            host = _PINNED[key] = torch.empty(big.shape, dtype=big.dtype, pin_memory=True)
        host.copy_(big, non_blocking=True)
        torch.cuda.current_stream(big.device).synchronize()     # the pinned copy is complete
        rec = host.numpy()
    else:
        rec = big.numpy()
    out = np.empty((n_total,) + item_shape, dtype=local_block.dtype)
    for r_ in range(world):
        k_ = int((rid[r_] >= 0).sum())                  # a rank's valid rows come first
        out[rid[r_, :k_]] = rec[r_, :k_]
    return out


_PINNED = {}
_GATHERED_IDS = {}


def barrier_and_max(seconds, device=None):
    """Max over ranks of a per-rank duration (the contract's timing rule)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
