"""Problem types of the host-side mirror (the SciMLBase subset the kernel path touches).

Reference: `ODEProblem`/`SDEProblem`/`EnsembleProblem` come from SciMLBase (un-vendored);
`make_prob_compatible` is src/utils.jl:37-57; the batch `probs` handed to
`vectorized_solve` is `adapt(dev, adapt.((dev,), probs))` (src/solve.jl:399-400), an AoS vector
of ImmutableODEProblem.  Here the adapted batch is `ProblemBatch`: three strided device arrays
(u0, p, tspan) -- the form the C ABI takes.

The right-hand side is not a host closure: a Julia `f(u,p,t)` cannot cross a C ABI.  An
`ODEFunction` names a built-in model or carries CUDA C++ *bodies* (what Symbolics'
`build_function(target = CTarget())` emits) that NVRTC inlines into the stepper kernels.
"""
from dataclasses import dataclass, field, replace
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import _lib


@dataclass(frozen=True)
class ODEFunction:
    """f(u, p, t) with optional analytic `jac` / `tgrad` (reference: SciMLBase.ODEFunction
    with `jac =`, `tgrad =`, test/lower_level_api.jl:19-47).

    builtin : name of a model compiled into libdegk, or None
    rhs/jac/tgrad : CUDA C++ bodies writing du[i] / J[i][j] / dT[i] from u[i], p[i], t, T
    python  : optional host callable f(u, p, t) -> du used only by tests for ground truth
    """
    builtin: Optional[str] = None
    rhs: Optional[str] = None
    jac: Optional[str] = None
    tgrad: Optional[str] = None
    n_state: int = 0
    n_param: int = 0
    python: Optional[Callable] = field(default=None, compare=False)
    force_jit: bool = False
    mass_matrix: Optional[str] = None   # body assigning Mm[i][j] of a constant mass matrix (stiff solvers), None = identity
    use_jac: bool = True      # False: ignore the analytic Jacobian (the function "has no jac"), for the AD / FD paths
    initialize: bool = False  # mass-matrix DAE: solve the algebraic equations for consistent initial values before the first
                              # step (the reference does when the function carries initialization data, kernels.jl:19-25)

    def __post_init__(self):
        if self.builtin is None and self.rhs is None:
            raise ValueError("ODEFunction needs `builtin=` or an `rhs=` source body")
        if self.builtin is None and self.n_state <= 0:
            raise ValueError("n_state is required for a source-defined ODEFunction")

    @classmethod
    def from_python(cls, f, n_state, n_param=0, *, jac=False, mass_matrix=None):
        """Lower a host function f(u, p, t) -> du to CUDA C++ bodies by tracing it once (lowering.py; the role of
        Symbolics `build_function(target = CTarget())` / ModelingToolkit in docs/src/tutorials/modelingtoolkit.md).
        `jac=True` also derives the analytic Jacobian and time gradient symbolically (`ODEFunction(f; jac, tgrad)`).
        `mass_matrix`: constant n x n array (zeros are skipped), as `ODEFunction(f; mass_matrix = M)`."""
        from . import lowering
        src = lowering.lower_function(f, n_state, n_param, jac=jac)
        mm = None
        if mass_matrix is not None:
            M = np.asarray(mass_matrix, dtype=np.float64)
            if M.shape != (n_state, n_state):
                raise ValueError(f"mass_matrix must be {n_state} x {n_state}")
            mm = "".join(f"    Mm[{i}][{j}] = (T){float(M[i, j])!r};\n" for i in range(n_state) for j in range(n_state) if M[i, j] != 0)

        def host(u, p, t):
            return np.asarray(f(list(u), list(p), t), dtype=np.float64)
        return cls(rhs=src["rhs"], jac=src["jac"], tgrad=src["tgrad"], n_state=n_state, n_param=n_param, python=host,
                   mass_matrix=mm)


@dataclass(frozen=True)
class SDEFunction:
    """drift f and diffusion g.  noise: 'diagonal' (g[i]) or 'general' (G[i][j], n x m)."""
    f: ODEFunction
    g: Optional[str] = None
    noise: str = "diagonal"
    n_noise: int = 0

    @classmethod
    def from_python(cls, f, g, n_state, n_param=0, *, noise="diagonal", n_noise=0):
        """Lower drift f(u, p, t) and diffusion g(u, p, t) (a vector for diagonal noise, an n x m matrix with
        `noise="general"`, the reference's `noise_rate_prototype`) to CUDA C++ bodies (lowering.py)."""
        from . import lowering
        if noise not in ("diagonal", "general"):
            raise ValueError("noise must be 'diagonal' or 'general'")
        body = lowering.lower_noise(g, n_state, n_param, noise=noise, n_noise=n_noise)
        return cls(ODEFunction.from_python(f, n_state, n_param), g=body, noise=noise,
                   n_noise=n_noise if noise == "general" else 0)


BUILTIN_DIMS = {  # name -> (n_state, n_param, n_noise, noise_kind); mirror of degk_models.cuh
    "lorenz": (3, 3, 3, 1), "henon_heiles": (4, 0, 0, 0), "rober": (3, 3, 0, 0), "rober_dae": (3, 3, 0, 0),
    "decay": (1, 1, 0, 0), "linear15": (15, 0, 0, 0), "gbm": (3, 2, 3, 1),
    "scalar_sde": (1, 2, 1, 1), "osc_t": (2, 1, 0, 0), "gbm_nd": (2, 2, 4, 2),
}


def _dims(f: ODEFunction):
    if f.builtin is not None and f.rhs is None:
        n, npar, _, _ = BUILTIN_DIMS[f.builtin]
        return n, npar
    return f.n_state, f.n_param


def _as_vec(x, dtype, n, what):
    a = np.asarray(x, dtype=dtype).reshape(-1)
    if a.size != n:
        raise ValueError(f"{what} has {a.size} entries, model expects {n}")
    return a


@dataclass(frozen=True)
class ODEProblem:
    """ODEProblem{false}(f, u0::SVector, tspan, p::SVector; kwargs...).  The element type of
    `u0` fixes T (Float32/Float64) like `eltype(prob.tspan)` does in lowerlevel_solve.jl:265."""
    f: ODEFunction
    u0: np.ndarray
    tspan: tuple
    p: Optional[np.ndarray] = None
    kwargs: dict = field(default_factory=dict)

    def __post_init__(self):
        u0 = np.asarray(self.u0)
        dtype = u0.dtype if u0.dtype in (np.float32, np.float64) else np.dtype(np.float64)
        if callable(self.f) and not isinstance(self.f, ODEFunction):
            # `ODEProblem{false}(f, u0, tspan, p)` with a plain host function: lowered by tracing (lowering.py)
            object.__setattr__(self, "f", ODEFunction.from_python(self.f, u0.size, 0 if self.p is None else np.size(self.p)))
        n, npar = _dims(self.f)
        object.__setattr__(self, "u0", _as_vec(u0, dtype, n, "u0"))
        p = np.zeros(0, dtype) if self.p is None else np.asarray(self.p, dtype=dtype).reshape(-1)
        if p.size != npar:
            raise ValueError(f"p has {p.size} entries, model expects {npar}")
        object.__setattr__(self, "p", p)
        t0, tf = self.tspan
        object.__setattr__(self, "tspan", (dtype.type(t0), dtype.type(tf)))

    @property
    def dtype(self):
        return self.u0.dtype


@dataclass(frozen=True)
class SDEProblem:
    """SDEProblem(f, g, u0, tspan, p; noise_rate_prototype, seed)."""
    f: SDEFunction
    u0: np.ndarray
    tspan: tuple
    p: Optional[np.ndarray] = None
    seed: int = 0
    kwargs: dict = field(default_factory=dict)

    def __post_init__(self):
        u0 = np.asarray(self.u0)
        dtype = u0.dtype if u0.dtype in (np.float32, np.float64) else np.dtype(np.float64)
        n, npar = _dims(self.f.f)
        object.__setattr__(self, "u0", _as_vec(u0, dtype, n, "u0"))
        p = np.zeros(0, dtype) if self.p is None else np.asarray(self.p, dtype=dtype).reshape(-1)
        if p.size != npar:
            raise ValueError(f"p has {p.size} entries, model expects {npar}")
        object.__setattr__(self, "p", p)
        t0, tf = self.tspan
        object.__setattr__(self, "tspan", (dtype.type(t0), dtype.type(tf)))

    @property
    def dtype(self):
        return self.u0.dtype

    def is_diagonal_noise(self):
        if self.f.g is None and self.f.f.builtin is not None:
            return BUILTIN_DIMS[self.f.f.builtin][3] == 1
        return self.f.noise == "diagonal"


def remake(prob, **changes):
    """SciMLBase.remake: copy with some fields replaced."""
    return replace(prob, **changes)


def make_prob_compatible(prob):
    """reference src/utils.jl:37-57: ODEProblem -> ImmutableODEProblem.  Problems here are
    already immutable value types with static-size state; returned unchanged."""
    return prob


@dataclass
class EnsembleProblem:
    """SciMLBase.EnsembleProblem(prob; prob_func, output_func, reduction, u_init, safetycopy).
    prob_func(prob, ctx) -> prob where ctx has `.sim_id` (1-based, like the reference's
    `_make_ensemble_context(i, ...)`, src/solve.jl:187-202)."""
    prob: object
    prob_func: Optional[Callable] = None
    output_func: Optional[Callable] = None
    reduction: Optional[Callable] = None
    u_init: object = None
    safetycopy: bool = True


@dataclass(frozen=True)
class EnsembleContext:
    sim_id: int
    sim_seed: Optional[int] = None


class ProblemBatch:
    """`probs` adapted to the device: u0 (N, n), p (N, np) and tspan ((2,) shared or (N, 2))
    as torch CUDA tensors.  Reference: `adapt(dev, adapt.((dev,), probs))`, src/solve.jl:399-400.
    """

    def __init__(self, prob, u0, p, tspan, n_traj, seed=0, saveat=None):
        self.prob = prob
        self.u0, self.p, self.tspan = u0, p, tspan
        self.n_traj = int(n_traj)
        self.seed = int(seed)
        # per-problem saveat grids, (N, nsave) -- `ODEProblem(...; saveat = ...)` carried in prob.kwargs by the reference
        # (kernels.jl:15-17, 89-91); all of the same length (src/solve.jl:226-245).  None: the solver's `saveat` keyword
        self.saveat = saveat

    def __len__(self):
        return self.n_traj

    @property
    def device(self):
        return self.u0.device

    @staticmethod
    def from_arrays(prob, *, u0=None, p=None, tspan=None, n_traj=None, device="cuda", seed=None, saveat=None):
        """Build a batch directly from arrays (no per-trajectory Python objects).
        Each of u0/p/tspan may be None (use the prototype's value, broadcast) or an array with
        a leading trajectory axis."""
        dt = torch.float32 if prob.dtype == np.float32 else torch.float64
        dev = torch.device(device)

        def conv(x, proto, width):
            if x is None:
                return torch.as_tensor(np.asarray(proto), dtype=dt).reshape(-1).to(dev)
            if not isinstance(x, torch.Tensor):
                x = torch.as_tensor(np.ascontiguousarray(x))
            x = x.to(device=dev, dtype=dt).contiguous()
            if x.ndim == 1 and x.numel() == width:
                return x
            if x.ndim != 2 or x.shape[1] != width:
                raise ValueError(f"expected shape (N, {width}), got {tuple(x.shape)}")
            return x

        n = prob.u0.size
        u0_t = conv(u0, prob.u0, n)
        p_t = conv(p, prob.p, prob.p.size) if prob.p.size else torch.zeros(0, dtype=dt, device=dev)
        ts_t = conv(tspan, prob.tspan, 2)
        sizes = [t.shape[0] for t in (u0_t, p_t, ts_t) if t.ndim == 2]
        if n_traj is None:
            if not sizes:
                raise ValueError("n_traj is required when every input is broadcast")
            n_traj = sizes[0]
        if any(s != n_traj for s in sizes):
            raise ValueError("u0/p/tspan disagree on the number of trajectories")
        if seed is None:
            seed = getattr(prob, "seed", 0)
        sv_t = None
        if saveat is not None:
            sv_t = torch.as_tensor(np.ascontiguousarray(saveat)).to(device=dev, dtype=dt).contiguous()
            if sv_t.ndim != 2 or sv_t.shape[0] != n_traj:
                raise ValueError("per-problem saveat must have shape (N, nsave): grids of the same length for every trajectory")
        return ProblemBatch(prob, u0_t, p_t, ts_t, n_traj, seed, sv_t)

    @staticmethod
    def from_problems(probs: Sequence, device="cuda"):
        """AoS list of problems -> SoA-per-field batch (host loop #1 of src/solve.jl:187-202)."""
        if len(probs) == 0:
            raise ValueError("empty batch")
        proto = probs[0]
        u0 = np.stack([pr.u0 for pr in probs])
        p = np.stack([pr.p for pr in probs]) if proto.p.size else None
        ts = np.array([pr.tspan for pr in probs], dtype=proto.dtype)
        same_t = bool((ts == ts[0]).all())
        # inner saveat (prob.kwargs[:saveat]): all problems or none, and all of the same length -- src/solve.jl:226-245
        svs = [pr.kwargs.get("saveat") for pr in probs]
        saveat = None
        if any(sv is not None for sv in svs):
            if any(sv is None for sv in svs) or len({np.size(sv) for sv in svs}) != 1:
                raise ValueError("Using different saveat in EnsembleGPUKernel requires all of them to be of same length. "
                                 "Use saveats of same size only.")
            saveat = np.stack([np.asarray(sv, dtype=proto.dtype).reshape(-1) for sv in svs])
        return ProblemBatch.from_arrays(proto, u0=u0, p=p, tspan=None if same_t else ts,
                                        n_traj=len(probs), device=device, saveat=saveat)


def adapt(device, probs):
    """reference: adapt(dev, probs)"""
    if isinstance(probs, ProblemBatch):
        return probs
    return ProblemBatch.from_problems(list(probs), device=device)
