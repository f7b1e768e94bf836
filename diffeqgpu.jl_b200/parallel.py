"""Multi-GPU sharding of an ensemble: one process per GPU, trajectories split by contiguous
index range, no data-path collective.

Replaces the reference's `pmap` over Distributed workers (src/solve.jl:131-152, one Julia
worker per GPU as in docs/src/tutorials/multigpu.md:19-31): same partitioning idea (contiguous
chunks of trajectories), but the plumbing is torch.distributed and nothing is serialised over
sockets -- results stay on their GPU.  The only collective is the optional ensemble reduction
(mean / variance of the saved states), which replaces the host-side `reduction(u, batch, I)`
of src/solve.jl:123-125, 145-146 by one all-reduce of a few hundred bytes.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_traj, rank, world_size):
    """Contiguous range [lo, hi) of trajectories owned by `rank`; sizes differ by at most 1."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(n_traj), world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK/WORLD_SIZE/MASTER_* (torchrun); no-op for 1 rank."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def allreduce_moments(partial, n_local):
    """partial: (rows, n, 2) float64 tensor of [sum u, sum u^2] over this rank's trajectories
    (the `reduce` output of degk_solve).  Returns (mean, var, n_total) over the whole ensemble.
    One all-reduce of rows*n*2 + 1 doubles; NVLink bandwidth is irrelevant at this size."""
    buf = torch.cat([partial.reshape(-1).to(torch.float64),
                     torch.tensor([float(n_local)], dtype=torch.float64, device=partial.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    n_total = buf[-1].item()
    s = buf[:-1].reshape(partial.shape)
    mean = s[..., 0] / n_total
    var = s[..., 1] / n_total - mean * mean
    return mean, var, int(round(n_total))


def max_over_ranks(value, device=None):
    """device-timed milliseconds -> max over ranks (bench contract)"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def sum_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()
