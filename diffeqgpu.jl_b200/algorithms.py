"""Algorithm and ensemble selector types (mirror of the reference's dispatch tags).

Reference: src/ensemblegpukernel/gpukernel_algorithms.jl:30-266 (GPUTsit5 ... GPUSIEA),
src/ensemblegpukernel/alg_utils.jl:5-15 (alg_order), src/algorithms.jl:192-211
(EnsembleGPUKernel(dev, cpu_offload = 0.0)).
"""
from dataclasses import dataclass

from . import _lib


class GPUODEAlgorithm:
    """reference: abstract type GPUODEAlgorithm (src/DiffEqGPU.jl:77-187)"""
    alg_id = -1
    order = None
    is_sde = False
    is_stiff = False

    def __repr__(self):
        return f"{type(self).__name__}()"

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self))


class GPUSDEAlgorithm(GPUODEAlgorithm):
    is_sde = True


class GPUODEImplicitAlgorithm(GPUODEAlgorithm):
    """reference: GPUODEImplicitAlgorithm{AD}.  Jacobian / time gradient as in nlsolve/type.jl:129-157:
    the function's own `jac` when it has one, else forward-mode duals (autodiff = True, the
    default) or finite differences (autodiff = False)."""
    is_stiff = True

    def __init__(self, autodiff=True):
        self.autodiff = autodiff


class GPUTsit5(GPUODEAlgorithm):
    alg_id, order = 0, 5


class GPUVern7(GPUODEAlgorithm):
    alg_id, order = 1, 7


class GPUVern9(GPUODEAlgorithm):
    alg_id, order = 2, 9


class GPURosenbrock23(GPUODEImplicitAlgorithm):
    alg_id, order = 3, 2


class GPURodas4(GPUODEImplicitAlgorithm):
    alg_id, order = 4, 4


class GPURodas5P(GPUODEImplicitAlgorithm):
    alg_id, order = 5, 5


class GPUKvaerno3(GPUODEImplicitAlgorithm):
    """ESDIRK + Newton (gpu_kvaerno3_perform_step.jl); saveat through the default Hermite interpolant"""
    alg_id, order = 8, 3


class GPUKvaerno5(GPUODEImplicitAlgorithm):
    alg_id, order = 9, 5


class GPUEM(GPUSDEAlgorithm):
    alg_id, order = 6, 1


class GPUSIEA(GPUSDEAlgorithm):
    alg_id, order = 7, 2


def alg_order(alg):
    """reference: alg_utils.jl:1-15"""
    if alg.order is None:
        raise ValueError("Order is not defined for this algorithm")
    return alg.order


@dataclass(frozen=True)
class EnsembleGPUKernel:
    """reference: src/algorithms.jl:192-211.  `dev` is a CUDA device (torch.device, index or
    "cuda"); the only backend is CUDA on sm_100a.  cpu_offload must stay 0.0: the product has
    no CPU path (north_star), so the reference's EnsembleThreads split (src/solve.jl:37-81) is
    rejected instead of silently ignored.

    Engine knobs that have no reference counterpart:
      fp_mode  "strict" (bit-parity with the reference's un-fused arithmetic) | "fast" (FFMA)
      schedule "auto" | "static" | "queue"  (adaptive divergence scheduling)
    """
    dev: object = "cuda"
    cpu_offload: float = 0.0
    fp_mode: str = "strict"
    schedule: str = "auto"

    def __post_init__(self):
        if self.cpu_offload != 0.0:
            raise ValueError("cpu_offload != 0 is not supported: this engine has no CPU path")
        if self.fp_mode not in ("strict", "fast"):
            raise ValueError("fp_mode must be 'strict' or 'fast'")
        if self.schedule not in ("auto", "static", "queue"):
            raise ValueError("schedule must be 'auto', 'static' or 'queue'")


FP_MODES = {"strict": _lib.FP_STRICT, "fast": _lib.FP_FAST}
SCHEDULES = {"auto": _lib.SCHED_AUTO, "static": _lib.SCHED_STATIC, "queue": _lib.SCHED_QUEUE}
