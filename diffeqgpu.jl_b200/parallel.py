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
from dataclasses import dataclass

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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs that are local to GPU `device_index` (same NUMA node / PCIe root), so that
    the pinned host buffers it allocates afterwards (first touch, local allocation policy) and the threads that fill
    them sit on the memory controllers the GPU's DMA engine reaches without crossing the socket interconnect.
    With 8 ranks on a two-socket host this decides whether the end-to-end path (13.6 GB D2H per rank and step on the
    C2 sweep) scales with the GPUs or saturates one socket.  Returns the CPU set used (empty: left unchanged)."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            local = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return set()


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


# ------------------------------------------------------------------------------------------
# ensemble moments: the `reduction` path of BASELINE config 5
# ------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class EnsembleMoments:
    """`EnsembleProblem(prob; reduction = EnsembleMoments())`: reduce the ensemble to the mean and variance of
    every saved row instead of returning N solutions.

    Replaces a host `reduction(u, batch, I)` that averages the batch (src/solve.jl:123-125, 145-146; the ensemble
    mean of test/gpu_kernel_de/gpu_sde_regression.jl:36-42): the SDE kernels accumulate sum(u) and sum(u^2) per
    saved row while they run (warp shuffle -> FP64 atomics), the ranks of a torchrun job then exchange
    rows * n * 2 + 1 doubles in ONE all-reduce (NCCL over NVLink on a multi-GPU box).  Trajectories are sharded
    over the ranks by contiguous index range; the noise stream is keyed by the global trajectory index, so the
    moments do not depend on the number of ranks or on batch_size."""


@dataclass
class MomentsSolution:
    """result of a moments-reduced ensemble: `t` (rows,), `mean` / `var` (rows, n), `n` trajectories"""
    t: np.ndarray
    mean: np.ndarray
    var: np.ndarray
    n: int
    n_ranks: int = 1
    stats: dict = None


def solve_moments(ensembleprob, alg, ensemblealg=None, *, trajectories, dt, batch_size=None, saveat=None,
                  save_everystep=False, adaptive=False, seed=None, **kwargs):
    """Solve `trajectories` members of the ensemble on the ranks of this job (1 rank without torch.distributed) and
    return the ensemble mean / variance per saved row.  Called by `solve(...)` when the problem's reduction is
    `EnsembleMoments()`; usable directly.  One collective: the final all-reduce of the moments."""
    from .algorithms import EnsembleGPUKernel
    from .lowerlevel_solve import vectorized_asolve, vectorized_solve
    from .problems import SDEProblem
    from .solve import _build_batch, _with_seed
    if ensemblealg is None:
        ensemblealg = EnsembleGPUKernel()
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    dev = torch.device(ensemblealg.dev)
    if world > 1 and dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    prob = ensembleprob.prob
    is_sde = isinstance(prob, SDEProblem)
    if is_sde and seed is not None:
        prob = _with_seed(prob, seed)
        ensembleprob = type(ensembleprob)(prob, ensembleprob.prob_func, ensembleprob.output_func,
                                          ensembleprob.reduction, ensembleprob.u_init, ensembleprob.safetycopy)
    lo, hi = shard_range(trajectories, rank, world)
    n_local = hi - lo
    if batch_size is None or batch_size <= 0:
        batch_size = max(n_local, 1)
    kw = dict(kwargs)
    kw.setdefault("fp_mode", ensemblealg.fp_mode)
    kw.setdefault("schedule", ensemblealg.schedule)
    red = None
    ts_row = None
    tot = None
    for b0 in range(lo, hi, batch_size):
        b1 = min(b0 + batch_size, hi)
        I = np.arange(b0 + 1, b1 + 1)                       # 1-based sim ids like the reference
        batch, _ = _build_batch(ensembleprob, I, dev)
        if adaptive:
            ts, us, st = vectorized_asolve(batch, prob, alg, dt=dt, saveat=saveat, save_everystep=save_everystep, stats=True, **kw)
        else:
            if is_sde:
                rows = len(saveat) if saveat is not None else (2 if not save_everystep else None)
                if rows is None:
                    raise ValueError("EnsembleMoments needs saveat or save_everystep = False")
                if red is None:
                    red = torch.zeros((rows, prob.u0.size, 2), dtype=torch.float64, device=dev)
                ts, us, st = vectorized_solve(batch, prob, alg, dt=dt, saveat=saveat, save_everystep=save_everystep,
                                              stats=True, traj_offset=b0, reduce=red, **kw)
            else:
                ts, us, st = vectorized_solve(batch, prob, alg, dt=dt, saveat=saveat, save_everystep=save_everystep, stats=True, **kw)
        if not (is_sde and not adaptive):
            # ODE ensembles: no fused reduction in the stepper kernels; the sums are formed on the device
            x = us.to(torch.float64)
            part = torch.stack([x.sum(0), (x * x).sum(0)], -1)
            red = part if red is None else red + part
        ts_row = ts[0] if ts.shape[0] > 0 else ts_row
        tot = st["totals"].clone() if tot is None else tot + st["totals"]
        del us
    if red is None:                                         # a rank without trajectories
        rows = len(saveat) if saveat is not None else 2
        red = torch.zeros((rows, prob.u0.size, 2), dtype=torch.float64, device=dev)
    if world > 1 and dist.get_backend() == "gloo":
        red = red.cpu()
    mean, var, n = allreduce_moments(red, n_local)
    t = None if ts_row is None else ts_row.cpu().numpy()
    if is_sde and not adaptive and saveat is None:
        t = np.asarray(prob.tspan, dtype=prob.dtype)      # endpoints: every trajectory ends at its own t; report the span
    return MomentsSolution(t=t, mean=mean.cpu().numpy(), var=var.cpu().numpy(), n=n, n_ranks=world,
                           stats=None if tot is None else dict(totals=tot.cpu().numpy()))
