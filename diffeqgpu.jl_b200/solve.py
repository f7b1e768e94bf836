"""solve(EnsembleProblem, alg, EnsembleGPUKernel(...); trajectories, ...) -- the batch glue.

Mirror of reference src/solve.jl: `__solve` :1-165 (batching by `batch_size`, `reduction`,
`u_init`), `batch_solve` :174-287 (prob_func loop, tspan/saveat consistency checks, solution
building with the "unwritten ts slot == t0 => Terminated" protocol :256-283) and
`batch_solve_up_kernel` :382-419 (H2D adapt, vectorized_(a)solve, D2H).

Two deviations, both forced by the engine's scope:
  * cpu_offload and pmap-over-Distributed are not reproduced (no CPU path; multi-GPU sharding
    is done one process per GPU by parallel.shard_range + torch.distributed).
  * `prob_func` may be vectorised: `prob_func.batched(prob, ids) -> dict(u0=..., p=..., tspan=...)`
    builds the whole batch as arrays, avoiding the O(N) host loop the reference itself calls
    out as the bottleneck (lowerlevel_solve.jl:125-129).
"""
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .algorithms import EnsembleGPUKernel, GPUODEAlgorithm, GPUSDEAlgorithm
from .lowerlevel_solve import vectorized_asolve, vectorized_solve
from .problems import (EnsembleContext, EnsembleProblem, ProblemBatch, SDEProblem,
                       make_prob_compatible)


@dataclass
class ODESolution:
    """SciMLBase.build_solution(prob, alg, ts[1:sol_idx], us[1:sol_idx]; retcode)"""
    prob: object
    alg: object
    t: np.ndarray
    u: np.ndarray          # (len, n)
    retcode: str = "Success"
    stats: Optional[dict] = None

    def __len__(self):
        return len(self.t)


@dataclass
class EnsembleSolution:
    """SciMLBase.EnsembleSolution(sols, elapsedTime, converged)"""
    u: object
    elapsedTime: float
    converged: bool = True
    stats: dict = field(default_factory=dict)

    def __len__(self):
        return len(self.u)

    def __getitem__(self, i):
        return self.u[i]


def _build_batch(ensembleprob, I, device):
    """host loop #1 (src/solve.jl:187-202)"""
    prob = ensembleprob.prob
    pf = ensembleprob.prob_func
    if pf is None:
        return ProblemBatch.from_arrays(prob, n_traj=len(I), device=device), None
    if hasattr(pf, "batched"):
        arrays = pf.batched(prob, np.asarray(I))
        return ProblemBatch.from_arrays(prob, n_traj=len(I), device=device, **arrays), None
    probs = [make_prob_compatible(pf(prob, EnsembleContext(int(i)))) for i in I]
    # src/solve.jl:206-245: differing tspans need saveat or endpoints-only
    return ProblemBatch.from_problems(probs, device=device), probs


def batch_solve(ensembleprob, alg, ensemblealg, I, adaptive, **kwargs):
    """reference src/solve.jl:174-287 (EnsembleGPUKernel branch) + :382-419"""
    if len(I) == 0:
        raise ValueError("empty batch")
    if not isinstance(alg, (GPUODEAlgorithm, GPUSDEAlgorithm)):
        raise ValueError("We don't have solvers implemented for this algorithm yet")
    dev = torch.device(ensemblealg.dev)
    batch, probs = _build_batch(ensembleprob, I, dev)
    if batch.tspan.ndim == 2 and kwargs.get("saveat") is None and not adaptive and \
            kwargs.get("save_everystep", True):
        raise ValueError("Using different time-spans require either turning off save_everystep or "
                         "using saveat. If using saveat, it should be of same length across the ensemble.")
    kw = dict(kwargs)
    kw.setdefault("fp_mode", ensemblealg.fp_mode)
    kw.setdefault("schedule", ensemblealg.schedule)
    kw["traj_offset"] = int(I[0]) - 1 if not adaptive else 0
    if adaptive:
        kw.pop("traj_offset")
        ts, us, st = vectorized_asolve(batch, ensembleprob.prob, alg, stats=True, **kw)
    else:
        ts, us, st = vectorized_solve(batch, ensembleprob.prob, alg, stats=True, **kw)
    solts = ts.cpu().numpy()          # Array(ts), Array(us): src/solve.jl:416-417
    solus = us.cpu().numpy()
    rcs = st["retcode"].cpu().numpy()
    nacc = st["naccept"].cpu().numpy()
    nrej = st["nreject"].cpu().numpy()
    t0s = batch.tspan.cpu().numpy()
    out = []
    of = ensembleprob.output_func
    for j in range(len(I)):
        t0 = t0s[j, 0] if t0s.ndim == 2 else t0s[0]
        tsj = solts[j]
        nz = np.nonzero(tsj != t0)[0]
        if nz.size == 0:                      # src/solve.jl:260-264
            raise RuntimeError("Batch solve failed")
        sol_idx = int(nz[-1]) + 1
        retcode = "Success" if sol_idx == len(tsj) else "Terminated"
        if rcs[j] not in (0, 1):
            retcode = _lib.RETCODES.get(int(rcs[j]), "Failure")
        sol = ODESolution(probs[j] if probs is not None else ensembleprob.prob, alg,
                          tsj[:sol_idx], solus[j, :sol_idx], retcode,
                          dict(naccept=int(nacc[j]), nreject=int(nrej[j])))
        out.append(of(sol, EnsembleContext(int(I[j])))[0] if of is not None else sol)
    return out


def solve(ensembleprob, alg, ensemblealg=None, *, trajectories, batch_size=None, adaptive=True,
          seed=None, **kwargs):
    """reference `SciMLBase.__solve`, src/solve.jl:1-165 (defaults: adaptive = true,
    batch_size = trajectories)."""
    if ensemblealg is None:
        ensemblealg = EnsembleGPUKernel()
    if not isinstance(ensemblealg, EnsembleGPUKernel):
        raise TypeError("only EnsembleGPUKernel is implemented by this engine")
    if not isinstance(ensembleprob, EnsembleProblem):
        raise TypeError("expected an EnsembleProblem")
    from .parallel import EnsembleMoments, solve_moments
    if isinstance(ensembleprob.reduction, EnsembleMoments):
        # fused ensemble reduction (+ one all-reduce over the ranks of a torchrun job): BASELINE config 5
        t_start = time.perf_counter()
        kw = dict(kwargs)
        dt = kw.pop("dt")
        sol = solve_moments(ensembleprob, alg, ensemblealg, trajectories=trajectories, dt=dt, batch_size=batch_size,
                            adaptive=adaptive and not isinstance(ensembleprob.prob, SDEProblem), seed=seed, **kw)
        return EnsembleSolution(sol, time.perf_counter() - t_start, True)
    if batch_size is None:
        batch_size = trajectories
    if isinstance(ensembleprob.prob, SDEProblem) and seed is not None:
        ensembleprob = EnsembleProblem(_with_seed(ensembleprob.prob, seed), ensembleprob.prob_func,
                                       ensembleprob.output_func, ensembleprob.reduction,
                                       ensembleprob.u_init, ensembleprob.safetycopy)
    num_batches = trajectories // batch_size
    if num_batches * batch_size != trajectories:
        num_batches += 1
    t_start = time.perf_counter()
    if num_batches == 1 and ensembleprob.reduction is None:
        sol = batch_solve(ensembleprob, alg, ensemblealg, np.arange(1, trajectories + 1), adaptive, **kwargs)
        return EnsembleSolution(sol, time.perf_counter() - t_start, True)
    u = ensembleprob.u_init if ensembleprob.u_init is not None else []
    sols: List = []
    for b in range(num_batches):
        lo = batch_size * b + 1
        hi = trajectories if b == num_batches - 1 else batch_size * (b + 1)
        I = np.arange(lo, hi + 1)
        data = batch_solve(ensembleprob, alg, ensemblealg, I, adaptive, **kwargs)
        if ensembleprob.reduction is not None:
            u, _ = ensembleprob.reduction(u, data, I)
            sols = u
        else:
            sols.extend(data)
    return EnsembleSolution(sols, time.perf_counter() - t_start, True)


def _with_seed(prob, seed):
    from dataclasses import replace
    return replace(prob, seed=int(seed))


def solve_host(prob, alg, *, u0=None, p=None, tspan=None, n_traj=None, dt, adaptive=False,
               abstol=1e-6, reltol=1e-3, saveat=None, save_everystep=True, fp_mode="strict",
               schedule="auto", layout="ref", chunk_traj=0, out=None, stats=False, device=None,
               seed=0, traj_offset=0, reduce=False, engine="auto"):
    """End-to-end solve with HOST (numpy / pinned torch) buffers through `degk_solve_host`:
    the `batch_solve_up_kernel` equivalent (src/solve.jl:382-419) with the H2D upload of the
    problems, the solve and the D2H download of (ts, us) pipelined in chunks over three streams.

    u0/p/tspan: None (broadcast the prototype's value) or arrays with a leading trajectory axis.
    out: optional dict with preallocated `us`/`ts` (e.g. pinned) arrays to fill.
    Returns (ts, us[, stats]) as numpy arrays in the reference layout (N, len[, n]).
    """
    import ctypes as C

    from .algorithms import SCHEDULES
    from .lowerlevel_solve import (_convert_saveat_adaptive, _convert_saveat_fixed, get_program)
    Tt = prob.dtype
    prog = get_program(prob, alg, fp_mode, device)
    n = prog.info.n_state

    def host(x, proto, width):
        if x is None:
            return np.ascontiguousarray(np.asarray(proto, dtype=Tt).reshape(-1)), 0
        if isinstance(x, torch.Tensor):
            x = x.numpy()
        x = np.ascontiguousarray(x, dtype=Tt)
        if x.ndim == 1:
            return x, 0
        if x.shape[1] != width:
            raise ValueError(f"expected (N, {width})")
        return x, width

    u0_h, u0_s = host(u0, prob.u0, n)
    p_h, p_s = host(p, prob.p, prob.p.size) if prob.p.size else (None, 0)
    ts_h, ts_s = host(tspan, prob.tspan, 2)
    sizes = [a.shape[0] for a, s in ((u0_h, u0_s), (p_h, p_s), (ts_h, ts_s)) if s]
    N = n_traj if n_traj is not None else sizes[0]
    dcode = _lib.F32 if Tt == np.float32 else _lib.F64
    t0, tf = prob.tspan
    if saveat is not None:
        sv = (_convert_saveat_adaptive if adaptive else _convert_saveat_fixed)(saveat, prob)
        n_rows = len(sv)
    else:
        sv = None
        n_rows = int(_lib.lib().degk_output_rows(dcode, float(t0), float(tf), float(Tt.type(dt)),
                                                 int(adaptive), int(bool(save_everystep)), 0))
    shape_us = (N, n_rows, n) if layout == "ref" else (n_rows, n, N)
    shape_ts = (N, n_rows) if layout == "ref" else (n_rows, N)
    us = out["us"] if out and "us" in out else np.empty(shape_us, dtype=Tt)
    ts = out["ts"] if out and "ts" in out else np.empty(shape_ts, dtype=Tt)
    us_np = us.numpy() if isinstance(us, torch.Tensor) else us
    ts_np = ts.numpy() if isinstance(ts, torch.Tensor) else ts
    assert us_np.shape == shape_us and ts_np.shape == shape_ts and us_np.dtype == Tt
    a = _lib.SolveArgs()
    a.n_traj = N
    a.traj_offset = traj_offset
    a.u0 = u0_h.ctypes.data; a.u0_stride = u0_s
    a.p = p_h.ctypes.data if p_h is not None else None; a.p_stride = p_s
    a.tspan = ts_h.ctypes.data; a.tspan_stride = ts_s
    a.dt = float(Tt.type(dt)); a.adaptive = int(adaptive)
    a.abstol = float(Tt.type(abstol)); a.reltol = float(Tt.type(reltol))
    a.saveat = sv.ctypes.data if sv is not None else None
    a.n_saveat = 0 if sv is None else len(sv)
    a.save_everystep = int(bool(save_everystep))
    a.n_rows = n_rows
    a.us = us_np.ctypes.data; a.ts = ts_np.ctypes.data
    a.out_layout = _lib.LAYOUT_REF if layout == "ref" else _lib.LAYOUT_SOA
    a.schedule = SCHEDULES[schedule]
    a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    a.engine = _lib.ENGINES[engine]
    st = {}
    if stats == "totals":      # only the 4 global counters: no per-trajectory arrays to download
        st = dict(totals=np.zeros(4, np.uint64))
        a.totals = st["totals"].ctypes.data
    elif stats:
        st = dict(retcode=np.zeros(N, np.int32), naccept=np.zeros(N, np.int32),
                  nreject=np.zeros(N, np.int32), totals=np.zeros(4, np.uint64))
        a.retcode = st["retcode"].ctypes.data; a.naccept = st["naccept"].ctypes.data
        a.nreject = st["nreject"].ctypes.data; a.totals = st["totals"].ctypes.data
    if reduce:
        st["reduce"] = np.zeros((n_rows, n, 2), np.float64)
        a.reduce = st["reduce"].ctypes.data
    prog.solve_host(a, int(chunk_traj))
    if stats or reduce:
        return ts, us, st
    return ts, us
