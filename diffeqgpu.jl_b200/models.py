"""Ready-made ODE/SDE functions: the reference's test problems and the BASELINE.json systems.

`*_src` variants carry CUDA C++ bodies and take the NVRTC path; the plain ones name the
ahead-of-time kernels compiled into libdegk.  Python callables are for tests' ground truth.
"""
import numpy as np

from .problems import ODEFunction, SDEFunction


def _lorenz_py(u, p, t):
    return np.array([p[0] * (u[1] - u[0]), u[0] * (p[1] - u[2]) - u[1], u[0] * u[1] - p[2] * u[2]])


def _rober_py(u, p, t):
    return np.array([-p[0] * u[0] + p[2] * u[1] * u[2],
                     p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2], p[1] * u[1] ** 2])


def _hh_py(u, p, t):
    return np.array([u[2], u[3], -u[0] - 2 * u[0] * u[1], -u[1] - (u[0] ** 2 - u[1] ** 2)])


# test/gpu_kernel_de/gpu_ode_regression.jl:4-12, jac/tgrad: test/lower_level_api.jl:19-47
LORENZ_RHS = """
    du[0] = p[0] * (u[1] - u[0]);
    du[1] = u[0] * (p[1] - u[2]) - u[1];
    du[2] = u[0] * u[1] - p[2] * u[2];
"""
LORENZ_JAC = """
    J[0][0] = -p[0];        J[0][1] = p[0];
    J[1][0] = p[1] - u[2];  J[1][1] = (T)-1;  J[1][2] = -u[0];
    J[2][0] = u[1];         J[2][1] = u[0];   J[2][2] = -p[2];
"""
ROBER_RHS = """
    du[0] = -p[0] * u[0] + p[2] * u[1] * u[2];
    du[1] = p[0] * u[0] - p[1] * (u[1] * u[1]) - p[2] * u[1] * u[2];
    du[2] = p[1] * (u[1] * u[1]);
"""
ROBER_JAC = """
    J[0][0] = -p[0];  J[0][1] = p[2] * u[2];                        J[0][2] = p[2] * u[1];
    J[1][0] = p[0];   J[1][1] = (T)-2 * p[1] * u[1] - p[2] * u[2];  J[1][2] = -(p[2] * u[1]);
    J[2][1] = (T)2 * p[1] * u[1];
"""
LINEAR15_RHS = "    DEGK_UNROLL for (int i = 0; i < 15; ++i) du[i] = (T)1.01 * u[i];\n"
LINEAR15_JAC = "    DEGK_UNROLL for (int i = 0; i < 15; ++i) J[i][i] = (T)1.01;\n"

lorenz = ODEFunction(builtin="lorenz", python=_lorenz_py)
lorenz_src = ODEFunction(rhs=LORENZ_RHS, jac=LORENZ_JAC, n_state=3, n_param=3, python=_lorenz_py)
lorenz_jit = ODEFunction(builtin="lorenz", force_jit=True, python=_lorenz_py)
henon_heiles = ODEFunction(builtin="henon_heiles", python=_hh_py)
rober = ODEFunction(builtin="rober", python=_rober_py)
rober_src = ODEFunction(rhs=ROBER_RHS, jac=ROBER_JAC, n_state=3, n_param=3, python=_rober_py)
decay = ODEFunction(builtin="decay", python=lambda u, p, t: np.array([-p[0] * u[0]]))
osc_t = ODEFunction(builtin="osc_t", python=lambda u, p, t: np.array([u[1], -u[0] + p[0] * np.cos(t)]))
linear15 = ODEFunction(builtin="linear15", force_jit=True, python=lambda u, p, t: 1.01 * np.asarray(u))
linear15_src = ODEFunction(rhs=LINEAR15_RHS, jac=LINEAR15_JAC, n_state=15, n_param=0,
                           python=lambda u, p, t: 1.01 * np.asarray(u))

# Robertson as a DAE with mass matrix diag(1, 1, 0): test/gpu_kernel_de/stiff_ode/gpu_ode_mass_matrix.jl:5-31
ROBER_DAE_RHS = """
    du[0] = -p[0] * u[0] + p[2] * u[1] * u[2];
    du[1] = p[0] * u[0] - p[1] * (u[1] * u[1]) - p[2] * u[1] * u[2];
    du[2] = u[0] + u[1] + u[2] - (T)1;
"""
ROBER_DAE_JAC = """
    J[0][0] = p[0] * (T)-1;  J[0][1] = u[2] * p[2];                                J[0][2] = p[2] * u[1];
    J[1][0] = p[0];          J[1][1] = u[1] * p[1] * (T)-2 + u[2] * p[2] * (T)-1;  J[1][2] = p[2] * u[1] * (T)-1;
    J[2][0] = (T)1;          J[2][1] = (T)1;                                       J[2][2] = (T)1;
"""
rober_dae = ODEFunction(builtin="rober_dae", force_jit=True)
rober_dae_src = ODEFunction(rhs=ROBER_DAE_RHS, jac=ROBER_DAE_JAC, mass_matrix="    Mm[0][0] = (T)1; Mm[1][1] = (T)1;\n",
                            n_state=3, n_param=3)

# 2-state index-1 DAE with M = diag(1, 0): test/gpu_kernel_de/stiff_ode/gpu_ode_modelingtoolkit_dae.jl:22-44
# (the literals -0.04f0 and 1.0f4 of `dae_f` as parameters)
LIN_DAE_RHS = "    du[0] = -p[0] * u[0] + p[1] * u[1];\n    du[1] = u[0] + u[1] - (T)1;\n"
LIN_DAE_JAC = "    J[0][0] = -p[0];  J[0][1] = p[1];\n    J[1][0] = (T)1;   J[1][1] = (T)1;\n"
lin_dae_src = ODEFunction(rhs=LIN_DAE_RHS, jac=LIN_DAE_JAC, mass_matrix="    Mm[0][0] = (T)1;\n", n_state=2, n_param=2,
                          python=lambda u, p, t: np.array([-p[0] * u[0] + p[1] * u[1], u[0] + u[1] - 1]))

# bouncing ball x'' = -g: test/gpu_kernel_de/gpu_ode_continuous_callbacks.jl:6-10
BALL_RHS = "    du[0] = u[1];\n    du[1] = -p[0];\n"
ball_src = ODEFunction(rhs=BALL_RHS, n_state=2, n_param=1, python=lambda u, p, t: np.array([u[1], -p[0]]))
# with the analytic jac / tgrad of test/gpu_kernel_de/stiff_ode/gpu_ode_continuous_callbacks.jl:11-21
ball_jac_src = ODEFunction(rhs=BALL_RHS, jac="    J[0][1] = (T)1;\n", tgrad="", n_state=2, n_param=1,
                           python=lambda u, p, t: np.array([u[1], -p[0]]))

# test/gpu_kernel_de/finite_diff.jl:6-9 / forward_diff.jl: du = -p u^2, NO analytic Jacobian: the stiff
# solvers differentiate it (forward-mode duals, or finite differences with autodiff = False)
QUAD_DECAY_RHS = "    du[0] = -p[0] * u[0] * u[0];\n"
quad_decay_src = ODEFunction(rhs=QUAD_DECAY_RHS, n_state=1, n_param=1, python=lambda u, p, t: np.array([-p[0] * u[0] * u[0]]))
quad_decay_jac_src = ODEFunction(rhs=QUAD_DECAY_RHS, jac="    J[0][0] = (T)-2 * p[0] * u[0];\n", n_state=1, n_param=1)

# SDEs: test/gpu_kernel_de/gpu_sde_regression.jl:8-11 (dX = p1 X dt + p2 X dW), :46-55 (Lorenz +
# additive noise g = 3), :86-110 (non-diagonal 2x4 noise)
gbm = SDEFunction(ODEFunction(builtin="gbm"))
scalar_sde = SDEFunction(ODEFunction(builtin="scalar_sde"))
lorenz_additive = SDEFunction(ODEFunction(builtin="lorenz"))
gbm_nd = SDEFunction(ODEFunction(builtin="gbm_nd"), noise="general", n_noise=4)
gbm_src = SDEFunction(
    ODEFunction(rhs="    DEGK_UNROLL for (int i = 0; i < 3; ++i) du[i] = p[0] * u[i];\n", n_state=3, n_param=2),
    g="    DEGK_UNROLL for (int i = 0; i < 3; ++i) g[i] = p[1] * u[i];\n", noise="diagonal")
