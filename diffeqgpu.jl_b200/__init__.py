"""diffeqgpu.jl_b200 -- B200-native engine behind DiffEqGPU.jl's `EnsembleGPUKernel` path.

Host-side mirror of the reference's operator interface (same names and argument meaning):

    vectorized_solve / vectorized_asolve      src/ensemblegpukernel/lowerlevel_solve.jl
    GPUTsit5 ... GPUSIEA, EnsembleGPUKernel   gpukernel_algorithms.jl, src/algorithms.jl
    make_prob_compatible                      src/utils.jl
    solve(EnsembleProblem, alg, EnsembleGPUKernel; ...)   src/solve.jl (batch glue)

All numerics run in hand-written sm_100a CUDA kernels inside `libdegk.so`, reached through the
C ABI in include/degk.h.  There is no CPU fallback: importing works anywhere, solving needs a GPU.

The directory name contains a dot, so import it through the `diffeqgpu_b200` shim module at
the repository root:  `import diffeqgpu_b200 as dg`.
"""
from . import _lib, models
from ._lib import DegkError
from .algorithms import (EnsembleGPUKernel, GPUEM, GPUKvaerno3, GPUKvaerno5, GPUODEAlgorithm, GPUODEImplicitAlgorithm,
                         GPURodas4, GPURodas5P, GPURosenbrock23, GPUSDEAlgorithm, GPUSIEA,
                         GPUTsit5, GPUVern7, GPUVern9, alg_order)
from .callbacks import (CallbackSet, ContinuousCallback, DiscreteCallback, GPUContinuousCallback,
                        GPUDiscreteCallback)
from .lowerlevel_solve import Range, SolvePlan, get_program, vectorized_asolve, vectorized_solve
from .problems import (EnsembleContext, EnsembleProblem, ODEFunction, ODEProblem, ProblemBatch,
                       SDEFunction, SDEProblem, adapt, make_prob_compatible, remake)
from .solve import EnsembleSolution, ODESolution, solve, solve_host
from .parallel import EnsembleMoments, MomentsSolution, solve_moments

__all__ = [
    "DegkError", "EnsembleGPUKernel", "GPUEM", "GPUODEAlgorithm", "GPUODEImplicitAlgorithm",
    "GPURodas4", "GPURodas5P", "GPURosenbrock23", "GPUSDEAlgorithm", "GPUSIEA", "GPUTsit5",
    "GPUVern7", "GPUVern9", "alg_order", "Range", "get_program", "vectorized_asolve",
    "vectorized_solve", "SolvePlan", "EnsembleContext", "EnsembleProblem", "ODEFunction", "ODEProblem",
    "ProblemBatch", "SDEFunction", "SDEProblem", "adapt", "make_prob_compatible", "remake",
    "EnsembleSolution", "ODESolution", "solve", "solve_host", "models", "EnsembleMoments", "MomentsSolution", "solve_moments",
    "GPUKvaerno3", "GPUKvaerno5", "CallbackSet", "ContinuousCallback", "GPUContinuousCallback", "DiscreteCallback", "GPUDiscreteCallback",
]
