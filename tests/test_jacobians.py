"""Stiff solvers on models without an analytic Jacobian (reference nlsolve/type.jl:129-157,
tests test/gpu_kernel_de/forward_diff.jl and finite_diff.jl): forward-mode duals
(`autodiff = true`, the default) and finite differences (`autodiff = false`)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

f32, f64 = np.float32, np.float64
STIFF = {"rosenbrock23": "GPURosenbrock23", "rodas4": "GPURodas4", "rodas5p": "GPURodas5P"}


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle
    return oracle


@pytest.mark.parametrize("alg", list(STIFF))
def test_oracle_ad_and_fd_jacobians(oracle, alg):
    """finite_diff.jl:22-39 / forward_diff.jl: du = -p u^2, u0 = 10, tspan (0, 10), dt = 0.01 adaptive,
    `norm(sol.u - osol.u) < 2e-4`; the exact solution u0 / (1 + p u0 t) stands in for Rodas5P()"""
    kw = dict(dt=0.01, adaptive=True, save_everystep=False)        # abstol 1e-6, reltol 1e-3 defaults
    exact = 10.0 / 101.0
    tol = 2e-4 if alg != "rosenbrock23" else 2e-3    # at reltol 1e-3 Rosenbrock23 carries a 1e-3 error of its own
    r0 = oracle.solve("quad_decay", alg, [10.0], [1.0], [0, 10], **kw)
    for mode in (1, 2):
        r = oracle.solve("quad_decay", alg, [10.0], [1.0], [0, 10], jac_mode=mode, **kw)
        assert r["retcode"][0] == 1 and abs(r["us"][0, 1, 0] - exact) < tol
        assert abs(r["us"][0, 1, 0] - r0["us"][0, 1, 0]) < 1e-4
    # forward-mode duals reproduce the analytic Jacobian of polynomial models exactly
    k = (np.array([0.04, 3e7, 1e4]) * (0.5 + np.random.default_rng(0).random((16, 3)))).astype(f32)
    rkw = dict(dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4, saveat=np.array([1.0, 100.0, 1e4, 1e5], f32))
    a = oracle.solve("rober", alg, [1, 0, 0], k, [0, 1e5], **rkw)
    b = oracle.solve("rober", alg, [1, 0, 0], k, [0, 1e5], jac_mode=2, **rkw)
    assert np.array_equal(a["us"], b["us"]) and np.array_equal(a["naccept"], b["naccept"])
    # non-autonomous model: the time gradient matters (osc_t: x'' = -x + p cos t)
    okw = dict(dt=0.01, adaptive=True, abstol=1e-7, reltol=1e-7, save_everystep=False)
    a = oracle.solve("osc_t", alg, [1.0, 0.0], [0.5], [0, 10], **okw)
    for mode, tol in ((2, 1e-6), (1, 5e-4)):
        b = oracle.solve("osc_t", alg, [1.0, 0.0], [0.5], [0, 10], jac_mode=mode, **okw)
        assert np.abs(a["us"] - b["us"]).max() < tol


def test_jacobian_modes_compile_with_nvrtc():
    from diffeqgpu_b200 import _lib
    for mode in (0, 1, 2):
        for alg in (3, 4, 5):
            st, nb, log = _lib.jit_compile_check(_lib.make_desc(rhs_src="du[0] = -p[0] * u[0] * u[0];", n_state=1,
                                                                n_param=1, dtype=_lib.F32, alg=alg, jac_mode=mode))
            assert st == 0 and nb > 0, log
    # math functions differentiate too (DiffRules derivatives of sin/cos/exp/log/sqrt)
    body = "du[0] = u[1]; du[1] = -u[0] + p[0] * cos(t) / (T)2 + exp(-u[0] * u[0]) + sqrt(u[1] * u[1] + (T)1) - log(u[0] * u[0] + (T)2) * sin(u[1]);"
    st, nb, log = _lib.jit_compile_check(_lib.make_desc(rhs_src=body, n_state=2, n_param=1, dtype=_lib.F64, alg=5, jac_mode=2))
    assert st == 0, log


def _gpu(dg, func, alg, u0, p, tspan, *, autodiff=True, fp_mode="strict", dtype=f32, **kw):
    import torch
    u0 = np.asarray(u0, dtype); p = np.asarray(p, dtype)
    prob = dg.ODEProblem(func, u0[0] if u0.ndim == 2 else u0, tuple(tspan), p[0] if p.ndim == 2 else p)
    n = max(u0.shape[0] if u0.ndim == 2 else 1, p.shape[0] if p.ndim == 2 else 1)
    probs = dg.ProblemBatch.from_arrays(prob, u0=u0 if u0.ndim == 2 else None, p=p if p.ndim == 2 else None, n_traj=n, device="cuda:0")
    a = getattr(dg, STIFF[alg])(autodiff=autodiff)
    ts, us, st = dg.vectorized_asolve(probs, prob, a, fp_mode=fp_mode, stats=True, **kw)
    torch.cuda.synchronize()
    return dict(ts=ts.cpu().numpy(), us=us.cpu().numpy(), naccept=st["naccept"].cpu().numpy(),
                nreject=st["nreject"].cpu().numpy(), retcode=st["retcode"].cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("alg", list(STIFF))
def test_gpu_ad_and_fd_jacobians_bit_exact(oracle, alg):
    import dataclasses
    import diffeqgpu_b200 as dg
    n = 200
    u0 = (10.0 + 0.01 * np.arange(n)[:, None]).astype(f32)
    p = (1.0 + 0.001 * np.arange(n)[:, None]).astype(f32)
    kw = dict(dt=f32(0.01), abstol=f32(1e-6), reltol=f32(1e-3), save_everystep=False)
    okw = dict(dt=0.01, adaptive=True, abstol=1e-6, reltol=1e-3, save_everystep=False)
    for autodiff, mode in ((True, 2), (False, 1)):
        g = _gpu(dg, dg.models.quad_decay_src, alg, u0, p, [0, 10], autodiff=autodiff, **kw)
        r = oracle.solve("quad_decay", alg, u0, p, [0, 10], jac_mode=mode, **okw)
        for key in ("ts", "us", "naccept", "nreject", "retcode"):
            assert np.array_equal(g[key], r[key]), (alg, mode, key)
    # a model WITH a Jacobian body uses it (has_jac branch), whatever the autodiff flag says
    g = _gpu(dg, dg.models.quad_decay_jac_src, alg, u0, p, [0, 10], autodiff=False, **kw)
    r = oracle.solve("quad_decay", alg, u0, p, [0, 10], **okw)
    assert np.array_equal(g["us"], r["us"])
    # built-in models asked to ignore their Jacobian: Robertson sweep by duals, forced oscillator (tgrad) both ways
    k = (np.array([0.04, 3e7, 1e4]) * (0.5 + np.random.default_rng(0).random((64, 3)))).astype(f32)
    sv = np.array([1.0, 100.0, 1e4, 1e5], f32)
    rober_nojac = dataclasses.replace(dg.models.rober, use_jac=False)
    g = _gpu(dg, rober_nojac, alg, [1, 0, 0], k, [0, 1e5], dt=f32(1e-4), abstol=f32(1e-8), reltol=f32(1e-4), saveat=sv)
    r = oracle.solve("rober", alg, [1, 0, 0], k, [0, 1e5], dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4, saveat=sv, jac_mode=2)
    assert np.array_equal(g["us"], r["us"]) and np.array_equal(g["naccept"], r["naccept"])
    osc_nojac = dataclasses.replace(dg.models.osc_t, use_jac=False)
    pp = np.linspace(0.1, 1.0, 50, dtype=f32)[:, None]
    for autodiff, mode in ((True, 2), (False, 1)):
        g = _gpu(dg, osc_nojac, alg, [1.0, 0.0], pp, [0, 10], autodiff=autodiff, dt=f32(0.01), abstol=f32(1e-6), reltol=f32(1e-6),
                 save_everystep=False)
        r = oracle.solve("osc_t", alg, [1.0, 0.0], pp, [0, 10], dt=0.01, adaptive=True, abstol=1e-6, reltol=1e-6,
                         save_everystep=False, jac_mode=mode)
        # cos/sin of the device (CUDA libm) and of the host differ in the last ulp, and at tol 1e-6 in
        # Float32 the error estimate of these steppers sits at round-off level, so the step sequences
        # differ (with the analytic tgrad too); the finite-difference time gradient additionally divides
        # a difference of two cosines by sqrt(eps) = 3.5e-4.  The solutions are compared.
        assert (g["retcode"] == 1).all() and np.abs(g["us"] - r["us"]).max() < (2e-4 if mode == 2 else 2e-3)


@pytest.mark.gpu
def test_gpu_high_level_solve_without_jacobian():
    """finite_diff.jl:22-39 through solve(EnsembleProblem, GPURodas5P(autodiff = false), EnsembleGPUKernel)"""
    import diffeqgpu_b200 as dg
    prob = dg.ODEProblem(dg.models.quad_decay_src, np.array([10.0], f32), (0.0, 10.0), np.array([1.0], f32))
    monteprob = dg.EnsembleProblem(prob, safetycopy=False)
    for alg in (dg.GPURosenbrock23(autodiff=False), dg.GPURodas4(autodiff=False), dg.GPURodas5P(autodiff=False),
                dg.GPURodas5P()):
        sol = dg.solve(monteprob, alg, dg.EnsembleGPUKernel(), trajectories=2, save_everystep=False, adaptive=True, dt=f32(0.01))
        tol = 2e-4 if not isinstance(alg, dg.GPURosenbrock23) else 2e-3
        assert abs(sol[0].u[-1, 0] - 10.0 / 101.0) < tol
        sol = dg.solve(monteprob, alg, dg.EnsembleGPUKernel(), trajectories=10_000, save_everystep=False, adaptive=True, dt=f32(0.01))
        assert len(sol) == 10_000
