"""vectorized_solve / vectorized_asolve -- the operator boundary of the kernel path.

Mirror of reference src/ensemblegpukernel/lowerlevel_solve.jl:
    vectorized_solve(probs, prob::ODEProblem, alg; dt, saveat, save_everystep, ...)   :53-131
    vectorized_solve(probs, prob::SDEProblem, alg; dt, saveat, save_everystep, ...)   :134-199
    vectorized_asolve(probs, prob::ODEProblem, alg; dt, saveat, abstol, reltol, ...)  :253-346
    vectorized_asolve(probs, prob::SDEProblem, ...) -> error                          :348-356
Same names, argument meaning, defaults, output sizing and error behaviour; the kernel launch
is replaced by `degk_solve` (include/degk.h).  Returns `(ts, us)` still on the device:
ts (N, len) and us (N, len, n) torch tensors whose memory is exactly the reference's
column-major (len x N) arrays.
"""
import contextlib
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .julia_ranges import lin_range, step_range
from .algorithms import (FP_MODES, SCHEDULES, GPUEM, GPUSIEA, GPUODEAlgorithm, GPUSDEAlgorithm)
from .callbacks import as_callback_set
from .problems import ODEProblem, ProblemBatch, SDEProblem, adapt

MAX_SAVEAT_LENGTH = 100_000   # lowerlevel_solve.jl:286


class Range:
    """start:step:stop or range(start, stop, length=n) -- stands in for Julia's AbstractRange."""

    def __init__(self, start, stop, *, step=None, length=None):
        if (step is None) == (length is None):
            raise ValueError("give exactly one of step / length")
        self.start, self.stop = float(start), float(stop)
        if length is None:
            length = int(math.floor((self.stop - self.start) / float(step) + 1e-9)) + 1
            self.stop = self.start + (length - 1) * float(step)
        self.length = int(length)

    def collect(self, dtype):
        # Tt.(collect(range(Tt(first), Tt(last), length = n)))  (lowerlevel_solve.jl:90-91)
        return lin_range(dtype.type(self.start), dtype.type(self.stop), self.length, dtype)


def _is_number(x):
    return isinstance(x, (int, float, np.integer, np.floating))


def _convert_saveat_fixed(saveat, prob):
    """lowerlevel_solve.jl:84-109 (also the SDE method :159-178)"""
    Tt = prob.dtype
    if isinstance(saveat, Range):
        return saveat.collect(Tt)
    if _is_number(saveat):
        t0, tf = prob.tspan
        if Tt.type(saveat) == Tt.type(0.0):
            return np.array([t0, tf], dtype=Tt)
        num_points = int(math.ceil(abs(Tt.type(tf) - Tt.type(t0)) / abs(Tt.type(saveat)))) + 1
        return lin_range(Tt.type(t0), Tt.type(tf), num_points, Tt)
    return np.asarray(saveat).astype(Tt).reshape(-1)


def _convert_saveat_adaptive(saveat, prob):
    """lowerlevel_solve.jl:271-306"""
    Tt = prob.dtype
    if _is_number(saveat):
        t0, tf = prob.tspan
        if Tt.type(saveat) == Tt.type(0.0):
            return np.array([t0, tf], dtype=Tt)
        num_points = int(math.ceil(abs(Tt.type(tf) - Tt.type(t0)) / abs(Tt.type(saveat)))) + 1
        if num_points > MAX_SAVEAT_LENGTH:
            raise ValueError(f"saveat would create too many save points ({num_points}). "
                             "Consider using a larger saveat value.")
        return lin_range(Tt.type(t0), Tt.type(tf), num_points, Tt)
    if isinstance(saveat, Range):
        return saveat.collect(Tt)
    return np.asarray(saveat).astype(Tt).reshape(-1)


def fixed_dt_rows(Tt, t0, tf, dt, tstops=None):
    """rows of a fixed-dt, save-every-step run: len = length(t0:dt:tf) [+ length(tstops) - count(x -> x in tstops,
    timeseries)] (lowerlevel_solve.jl:64-78).  `timeseries` is the Julia range, whose elements are rounded once
    (julia_ranges.py), not t0 + k*dt: with tspan (0, 1), dt = 0.1 and tstops = [0.3] in Float64 that is 11 rows, not 12."""
    Tt = np.dtype(Tt)
    dcode = _lib.F32 if Tt == np.float32 else _lib.F64
    n_rows = int(_lib.lib().degk_output_rows(dcode, float(t0), float(tf), float(Tt.type(dt)), 0, 1, 0))
    if tstops is not None:
        series = step_range(Tt.type(t0), Tt.type(dt), n_rows, Tt)
        tst = [Tt.type(x) for x in tstops]
        n_rows += len(tst) - sum(1 for x in series if x in tst)
    return n_rows


def _model_key(f):
    return (f.builtin, f.rhs, f.jac, f.tgrad, f.n_state, f.n_param, f.force_jit, f.use_jac, f.mass_matrix)


def _jac_mode(f, alg):
    """nlsolve/type.jl:129-137: f.jac if the function has one, else ForwardDiff when the algorithm's
    autodiff flag is set (the default), else finite differences."""
    if not getattr(alg, "is_stiff", False):
        return 0, f.jac
    has_jac = f.use_jac and (f.jac is not None or (f.builtin is not None and f.rhs is None))
    if has_jac:
        return 0, f.jac
    return (2 if getattr(alg, "autodiff", True) else 1), None


def get_program(prob, alg, fp_mode="strict", device=None, callback=None, events=False):
    """Build (or fetch) the degk_program for (model, alg, eltype, fp mode[, callbacks]).
    `events` (tstops) or a non-empty callback set select the event-capable kernel pair."""
    cbs = as_callback_set(callback)
    events = bool(events) or len(cbs) > 0
    ctx = _lib.context(device)
    dtype = _lib.F32 if prob.dtype == np.float32 else _lib.F64
    if isinstance(prob, SDEProblem):
        if events:
            raise NotImplementedError("tstops / callbacks are lowered for ODE problems only")
        sf, f = prob.f, prob.f.f
        key = ("sde", _model_key(f), sf.g, sf.noise, sf.n_noise, alg.alg_id, dtype, fp_mode)
        if key in ctx._programs:
            return ctx._programs[key]
        kind = {"diagonal": _lib.NOISE_DIAGONAL, "general": _lib.NOISE_GENERAL}[sf.noise]
        desc = _lib.make_desc(builtin=f.builtin, rhs_src=f.rhs, noise_src=sf.g, n_state=f.n_state, n_param=f.n_param,
                              n_noise=sf.n_noise, noise_kind=kind, dtype=dtype, alg=alg.alg_id,
                              fp_mode=FP_MODES[fp_mode], force_jit=f.force_jit)
    else:
        f = prob.f
        jac_mode, jac_src = _jac_mode(f, alg)
        key = ("ode", _model_key(f), alg.alg_id, dtype, fp_mode, events, cbs.key(), cbs.ckey(), jac_mode)
        if key in ctx._programs:               # (the descriptor is only needed to build)
            return ctx._programs[key]
        desc = _lib.make_desc(builtin=f.builtin, rhs_src=f.rhs, jac_src=jac_src, tgrad_src=f.tgrad if jac_mode == 0 else None,
                              n_state=f.n_state, n_param=f.n_param, dtype=dtype, alg=alg.alg_id,
                              fp_mode=FP_MODES[fp_mode], force_jit=f.force_jit, events=events,
                              callbacks=cbs.key(), jac_mode=jac_mode, mass_src=f.mass_matrix, ccallbacks=cbs.ckey())
    return ctx.program(desc, key)


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _launch(probs, prob, alg, *, dt, adaptive, abstol, reltol, saveat, save_everystep, n_rows,
            fp_mode, schedule, layout, stats, stream, traj_offset=0, reduce=None, engine="auto",
            callback=None, tstops=None, sort_by=None, prepare=False):
    if not isinstance(probs, ProblemBatch):
        probs = adapt("cuda", probs)
    dev = probs.device
    if dev.type != "cuda":
        raise RuntimeError("probs must live on a CUDA device: this engine has no CPU path")
    prog = get_program(prob, alg, fp_mode, dev, callback=callback, events=tstops is not None)
    n = prog.info.n_state
    N = len(probs)
    tdt = torch.float32 if prob.dtype == np.float32 else torch.float64
    # allocations and the small H2D copies below are ordered on the stream the kernel is launched on (the caller's
    # `stream` when given): torch's caching allocator ties a block to the stream it was allocated on
    # (the context managers cost ~40 us of Python per call, more than a small ensemble's kernel: they are entered only
    #  when the launch stream / device differ from the current ones)
    launch_stream = torch.cuda.current_stream(dev) if stream is None else torch.cuda.ExternalStream(int(stream), device=dev)
    with contextlib.ExitStack() as ctx_stack:
        if dev.index is not None and torch.cuda.current_device() != dev.index:
            ctx_stack.enter_context(torch.cuda.device(dev))
        if stream is not None:
            ctx_stack.enter_context(torch.cuda.stream(launch_stream))
        # allocate(backend, T, (len, N)) -- lowerlevel_solve.jl:81-83 / 317-323.  ts needs no
        # fill!(ts, t0): the kernel writes t0 into every row it does not reach.
        if layout == "ref":
            ts = torch.empty((N, n_rows), dtype=tdt, device=dev)
            us = torch.empty((N, n_rows, n), dtype=tdt, device=dev)
        else:
            ts = torch.empty((n_rows, N), dtype=tdt, device=dev)
            us = torch.empty((n_rows, n, N), dtype=tdt, device=dev)
        d_saveat = None
        sv_stride = 0
        if getattr(probs, "saveat", None) is not None:
            d_saveat = probs.saveat                   # (N, nsave): every trajectory reads its own grid
            sv_stride = int(d_saveat.shape[1])
        elif saveat is not None:
            d_saveat = torch.as_tensor(saveat, dtype=tdt).to(dev)
        out = {}
        if stats:
            out["retcode"] = torch.zeros(N, dtype=torch.int32, device=dev)
            out["naccept"] = torch.zeros(N, dtype=torch.int32, device=dev)
            out["nreject"] = torch.zeros(N, dtype=torch.int32, device=dev)
            out["totals"] = torch.zeros(4, dtype=torch.int64, device=dev)
        a = _lib.SolveArgs()
        a.n_traj = N
        a.traj_offset = traj_offset
        a.u0 = _ptr(probs.u0); a.u0_stride = n if probs.u0.ndim == 2 else 0
        a.p = _ptr(probs.p); a.p_stride = probs.p.shape[1] if probs.p.ndim == 2 else 0
        a.tspan = _ptr(probs.tspan); a.tspan_stride = 2 if probs.tspan.ndim == 2 else 0
        a.dt = float(prob.dtype.type(dt))
        a.adaptive = int(adaptive)
        a.abstol = float(prob.dtype.type(abstol)); a.reltol = float(prob.dtype.type(reltol))
        a.saveat = _ptr(d_saveat)
        a.n_saveat = sv_stride if sv_stride else (0 if saveat is None else len(saveat))
        a.saveat_stride = sv_stride
        d_tstops = None
        if tstops is not None and len(tstops):
            d_tstops = torch.as_tensor(np.asarray(tstops, dtype=prob.dtype), dtype=tdt).to(dev)   # adapt(backend, tstops)
            a.tstops = d_tstops.data_ptr(); a.n_tstops = len(tstops)
        a.save_everystep = int(bool(save_everystep))
        a.n_rows = n_rows
        a.us = us.data_ptr(); a.ts = ts.data_ptr()
        a.out_layout = _lib.LAYOUT_REF if layout == "ref" else _lib.LAYOUT_SOA
        a.schedule = SCHEDULES[schedule]
        if stats:
            a.retcode = out["retcode"].data_ptr(); a.naccept = out["naccept"].data_ptr()
            a.nreject = out["nreject"].data_ptr(); a.totals = out["totals"].data_ptr()
        a.seed = int(getattr(probs, "seed", 0)) & 0xFFFFFFFFFFFFFFFF
        a.reduce = None if reduce is None else reduce.data_ptr()
        d_order = None
        if sort_by is not None and adaptive:
            # start order of the trajectories: sorted by a key that predicts the step count (north_star (5): "optional
            # sorting of trajectories by parameter").  `sort_by`: int = column of p, or a length-N tensor / array of keys
            key = probs.p[:, int(sort_by)] if isinstance(sort_by, (int, np.integer)) else torch.as_tensor(sort_by).to(dev)
            d_order = torch.argsort(key.reshape(-1)).to(torch.int32)
            a.order = d_order.data_ptr()
        a.max_iters = int(__import__("os").environ.get("DEGK_MAX_ITERS", "0"))   # 0 => library default (1e7 attempts per trajectory when adaptive, none for fixed dt)
        a.engine = _lib.ENGINES[engine]
        a.dae_init = int(bool(getattr(prob.f, "initialize", False))) if isinstance(prob, ODEProblem) else 0
        for buf in (probs.u0, probs.p, probs.tspan, reduce):     # inputs made on other streams stay alive until this one is done
            if isinstance(buf, torch.Tensor) and buf.is_cuda and buf.numel():
                buf.record_stream(launch_stream)
        # keep inputs alive until the stream has consumed them
        us._degk_keepalive = (probs, d_saveat, d_tstops, d_order)
        plan = SolvePlan(prog, a, dev, launch_stream, ts, us, out if stats else None, reduce)
        if prepare:
            return plan
        plan._launch()
    return plan.result()


class SolvePlan:
    """A prepared launch: program, marshalled `degk_solve_args` and output arrays of one `vectorized_solve` /
    `vectorized_asolve` call (`prepare=True`).  Calling the plan launches the kernel again into the same output arrays;
    the host side of a call is one ctypes call (a few microseconds instead of ~0.1 ms of argument conversion and
    allocation), which is what a small ensemble (BASELINE config 1: 10^4 trajectories, a ~30 us kernel) is bound by.
    `capture(k)` records k back-to-back launches in a CUDA graph, `replay()` launches that graph.
    The reference has no counterpart (every `solve` allocates and launches, lowerlevel_solve.jl:53-131)."""

    def __init__(self, prog, args, dev, stream, ts, us, stats, reduce):
        self.prog, self.args, self.device, self.stream = prog, args, dev, stream
        self.ts, self.us, self.stats, self.reduce = ts, us, stats, reduce
        self._graph = None

    def result(self):
        return (self.ts, self.us, self.stats) if self.stats is not None else (self.ts, self.us)

    def _launch(self):
        self.prog.solve(self.args, self.stream.cuda_stream)

    def _reset(self):
        if self.stats is not None:
            self.stats["totals"].zero_()                 # the kernels add to the totals
        if self.reduce is not None:
            self.reduce.zero_()

    def __call__(self):
        if self.stats is None and self.reduce is None:        # nothing for torch to do: one FFI call
            self._launch()
        else:
            with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                self._reset()
                self._launch()
        return self.result()

    def capture(self, launches=1):
        """Record `launches` launches in a CUDA graph (the plan must have run once: the first launch of a program
        sets kernel attributes, which stream capture does not allow)."""
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(self.stream)
        saved = self.stream
        # (thread-local capture mode: other threads of the process -- an NCCL watchdog, a data loader -- may go on
        #  calling the CUDA runtime while this one captures)
        try:
            with torch.cuda.device(self.device), torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                self.stream = torch.cuda.current_stream(self.device)
                for _ in range(int(launches)):
                    self._reset()
                    self._launch()
        finally:
            self.stream = saved                          # later direct launches go back to the plan's own stream
        self._graph = g
        return self

    def replay(self):
        if self._graph is None:
            raise RuntimeError("capture() first")
        self._graph.replay()
        return self.result()


def vectorized_solve(probs, prob, alg, *, dt, saveat=None, save_everystep=True, debug=False,
                     callback=None, tstops=None, fp_mode="strict", schedule="auto", layout="ref",
                     stats=False, stream=None, traj_offset=0, reduce=None, engine="auto", prepare=False, **kwargs):
    """Fixed-step batched solve; returns (ts, us) on the device (add `stats=True` for
    per-trajectory retcode/naccept/nreject and totals, which the reference does not have).
    `prepare=True` returns a `SolvePlan` instead of launching (re-launch with `plan()`).
    `engine`: "auto" (the lock-step kernel for launches with one tspan, every-step saves and an explicit RK
    stepper, else one thread per trajectory), "lockstep" (the same, also where "auto" would not stage the reference
    layout), "v1" (never)."""
    if not isinstance(alg, GPUODEAlgorithm):
        raise TypeError("alg must be a GPUODEAlgorithm / GPUSDEAlgorithm")
    is_sde = isinstance(prob, SDEProblem)
    if is_sde != isinstance(alg, GPUSDEAlgorithm):
        raise TypeError(f"{alg!r} cannot solve a {type(prob).__name__}")
    if is_sde and isinstance(alg, GPUSIEA) and not prob.is_diagonal_noise():
        # lowerlevel_solve.jl:185-186
        raise ValueError("The algorithm is not compatible with the chosen noise type. Please see "
                         "the documentation on the solver methods")
    Tt = prob.dtype
    dt = Tt.type(dt)
    dcode = _lib.F32 if Tt == np.float32 else _lib.F64
    t0, tf = prob.tspan
    if not isinstance(probs, ProblemBatch):
        probs = adapt("cuda", probs)
    saveat_c = None
    inner = getattr(probs, "saveat", None)
    if inner is not None:                       # saveat = _saveat === nothing ? saveat : _saveat  (kernels.jl:15-17)
        saveat_c = np.zeros(int(inner.shape[1]), dtype=Tt)
        n_rows = int(inner.shape[1])
    elif saveat is None:
        n_rows = fixed_dt_rows(Tt, t0, tf, dt, tstops) if save_everystep else 2
    else:
        saveat_c = _convert_saveat_fixed(saveat, prob)
        n_rows = len(saveat_c)
    return _launch(probs, prob, alg, dt=dt, adaptive=False, abstol=0.0, reltol=0.0, saveat=saveat_c,
                   save_everystep=save_everystep, n_rows=n_rows, fp_mode=fp_mode,
                   schedule=schedule, layout=layout, stats=stats, stream=stream,
                   traj_offset=traj_offset, reduce=reduce, callback=callback, tstops=tstops, engine=engine,
                   prepare=prepare)


def vectorized_asolve(probs, prob, alg, *, dt=np.float32(0.1), saveat=None, save_everystep=False,
                      abstol=np.float32(1e-6), reltol=np.float32(1e-3), debug=False, callback=None,
                      tstops=None, fp_mode="strict", schedule="auto", layout="ref", stats=False,
                      stream=None, engine="auto", sort_by=None, prepare=False, **kwargs):
    """Adaptive batched solve (defaults as lowerlevel_solve.jl:253-260).  `prepare=True`: see vectorized_solve."""
    if isinstance(prob, SDEProblem):
        raise RuntimeError("Adaptive time-stepping is not supported yet with GPUEM.")   # :348-356
    if not isinstance(alg, GPUODEAlgorithm) or isinstance(alg, GPUSDEAlgorithm):
        raise TypeError("alg must be a GPUODEAlgorithm")
    Tt = prob.dtype
    dt = Tt.type(dt)
    dcode = _lib.F32 if Tt == np.float32 else _lib.F64
    t0, tf = prob.tspan
    if not isinstance(probs, ProblemBatch):
        probs = adapt("cuda", probs)
    saveat_c = None if saveat is None else _convert_saveat_adaptive(saveat, prob)
    inner = getattr(probs, "saveat", None)
    if inner is not None:                       # the problems' own grids take precedence (kernels.jl:89-91)
        saveat_c = np.zeros(int(inner.shape[1]), dtype=Tt)
    if saveat_c is None:
        # len = ceil(Int, (tf - t0)/dt) + 1 when save_everystep else 2  (:311-316).  Only the
        # first row (+ nothing else) is ever written in the save_everystep case.
        n_rows = int(_lib.lib().degk_output_rows(dcode, float(t0), float(tf), float(dt), 1,
                                                 int(bool(save_everystep)), 0))
    else:
        n_rows = len(saveat_c)
    return _launch(probs, prob, alg, dt=dt, adaptive=True, abstol=abstol, reltol=reltol,
                   saveat=saveat_c, save_everystep=save_everystep, n_rows=n_rows, fp_mode=fp_mode,
                   schedule=schedule, layout=layout, stats=stats, stream=stream, engine=engine,
                   callback=callback, tstops=tstops, sort_by=sort_by, prepare=prepare)
