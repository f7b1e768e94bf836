"""GPUKvaerno3 / GPUKvaerno5 (ESDIRK + Newton nlsolve): oracle against the
reference's stiff regression assertions (test/gpu_kernel_de/stiff_ode/gpu_ode_regression.jl, exact
solutions standing in for OrdinaryDiffEq) and the device kernels against the oracle."""
import math
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

f32, f64 = np.float32, np.float64
KV = {"kvaerno3": "GPUKvaerno3", "kvaerno5": "GPUKvaerno5"}


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle
    return oracle


@pytest.mark.parametrize("alg", list(KV))
def test_oracle_kvaerno_regression(oracle, alg):
    # gpu_ode_regression.jl:36-60: du = -u, u0 = 10, tspan (0, 10); fixed dt = 0.01 and adaptive dt0 = 0.01
    r = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, length=1001)
    assert r["retcode"][0] == 1 and abs(r["us"][0, -1, 0] - 10 * math.exp(-10)) < 5e-3
    assert abs(r["us"][0, 100, 0] - 10 * math.exp(-float(r["ts"][0, 100]))) < 1e-4
    r = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, adaptive=True, save_everystep=False)
    assert r["retcode"][0] == 1 and abs(r["us"][0, 1, 0] - 10 * math.exp(-10)) < 6e-3 and r["ts"][0, 1] == f32(10.0)
    # save_everystep = false, fixed dt: row 2 is the raw final state (t may overshoot tf, Q2)
    r = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, save_everystep=False)
    assert abs(r["us"][0, 1, 0] - 10 * math.exp(-float(r["ts"][0, 1]))) < 5e-3
    # f_large = 1.01 u with 15 states (general LU inside every Newton iteration), :24-28, finite_diff / AD Jacobians
    u0 = np.random.default_rng(1).random(15).astype(f32)
    for mode in (0, 1, 2):
        r = oracle.solve("linear15", alg, u0, None, [0, 1], dt=0.01, adaptive=True, abstol=1e-6, reltol=1e-6,
                         save_everystep=False, jac_mode=mode)
        assert np.allclose(r["us"][0, 1], u0 * math.exp(1.01), rtol=2e-4), mode
    # Robertson: stiff, the invariant y1 + y2 + y3 = 1 and the published end values at t = 1e5
    k = np.array([[0.04, 3e7, 1e4]], f64)
    r = oracle.solve("rober", alg, [1, 0, 0], k, [0, 1e5], dt=1e-4, adaptive=True, abstol=1e-10, reltol=1e-8,
                     save_everystep=False, dtype=f64)
    assert np.allclose(r["us"][0, 1], [1.78659e-2, 7.27475e-8, 9.82134e-1], rtol=1e-4)
    # "solve parameters", :62-118: saveat = [2, 4] and 0:0.1:10 through the default Hermite interpolant
    # (the reference bounds against OrdinaryDiffEq's Rosenbrock23 are 2e-4 / 2e-3 fixed and 2e-3 / 3e-2 adaptive)
    sv = np.array([2.0, 4.0], f32)
    r = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, saveat=sv)
    a = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, adaptive=True, abstol=1e-7, reltol=1e-7, saveat=sv)
    exact = 10 * np.exp(-sv.astype(f64))
    assert np.array_equal(r["ts"][0], sv) and np.linalg.norm(r["us"][0, :, 0] - exact) < 2e-4
    assert np.array_equal(a["ts"][0], sv) and np.linalg.norm(a["us"][0, :, 0] - exact) < 2e-3
    assert np.linalg.norm(a["us"][0, -1] - r["us"][0, -1]) < 4e-2
    sv = np.arange(0, 101, dtype=f32) * f32(0.1)
    r = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, saveat=sv)
    a = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, adaptive=True, abstol=1e-7, reltol=1e-7, saveat=sv)
    exact = 10 * np.exp(-sv.astype(f64))
    assert r["us"].shape[1] == 101 and np.linalg.norm(r["us"][0, :, 0] - exact) < 2e-3
    assert np.linalg.norm(a["us"][0, :, 0] - exact) < 3e-2
    # Float64: the interpolation error of the Hermite cubic between accepted steps stays at the tolerance level
    a = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, adaptive=True, abstol=1e-9, reltol=1e-9,
                     saveat=sv.astype(f64), dtype=f64)
    assert np.abs(a["us"][0, :, 0] - 10 * np.exp(-sv.astype(f32).astype(f64))).max() < 1e-5


def _gpu(dg, func, alg, u0, p, tspan, *, adaptive, autodiff=True, fp_mode="strict", dtype=f32, **kw):
    import torch
    u0 = np.asarray(u0, dtype); p = None if p is None else np.asarray(p, dtype)
    prob = dg.ODEProblem(func, u0[0] if u0.ndim == 2 else u0, tuple(tspan), None if p is None else (p[0] if p.ndim == 2 else p))
    n = max(u0.shape[0] if u0.ndim == 2 else 1, p.shape[0] if p is not None and p.ndim == 2 else 1)
    probs = dg.ProblemBatch.from_arrays(prob, u0=u0 if u0.ndim == 2 else None, p=p if p is not None and p.ndim == 2 else None,
                                        n_traj=n, device="cuda:0")
    a = getattr(dg, KV[alg])(autodiff=autodiff)
    fn = dg.vectorized_asolve if adaptive else dg.vectorized_solve
    ts, us, st = fn(probs, prob, a, fp_mode=fp_mode, stats=True, **kw)
    torch.cuda.synchronize()
    return dict(ts=ts.cpu().numpy(), us=us.cpu().numpy(), naccept=st["naccept"].cpu().numpy(),
                nreject=st["nreject"].cpu().numpy(), retcode=st["retcode"].cpu().numpy())


def _same(g, r, what):
    for key in ("ts", "us", "naccept", "nreject", "retcode"):
        assert np.array_equal(g[key], r[key]), f"{what}: {key} differs"


@pytest.mark.gpu
@pytest.mark.parametrize("alg", list(KV))
def test_gpu_kvaerno_bit_exact(oracle, alg):
    import dataclasses
    import diffeqgpu_b200 as dg
    n = 150
    u0 = (10.0 + 0.01 * np.arange(n)[:, None]).astype(f32)
    p = (1.0 + 0.002 * np.arange(n)[:, None]).astype(f32)
    # fixed dt, every-step saves (the last row is the Hermite value at tf when t overshoots) and endpoints
    g = _gpu(dg, dg.models.decay, alg, u0, p, [0, 2], adaptive=False, dt=f32(0.01))
    r = oracle.solve("decay", alg, u0, p, [0, 2], dt=0.01, length=g["us"].shape[1])
    _same(g, r, "fixed every-step")
    g = _gpu(dg, dg.models.decay, alg, u0, p, [0, 2], adaptive=False, dt=f32(0.01), save_everystep=False)
    r = oracle.solve("decay", alg, u0, p, [0, 2], dt=0.01, save_everystep=False)
    _same(g, r, "fixed endpoints")
    # adaptive: analytic Jacobian, duals, finite differences (du = -p u^2 has no Jacobian body)
    akw = dict(dt=0.01, adaptive=True, abstol=1e-6, reltol=1e-4, save_everystep=False)
    gkw = dict(dt=f32(0.01), abstol=f32(1e-6), reltol=f32(1e-4), save_everystep=False)
    g = _gpu(dg, dg.models.decay, alg, u0, p, [0, 10], adaptive=True, **gkw)
    _same(g, oracle.solve("decay", alg, u0, p, [0, 10], **akw), "adaptive decay")
    for autodiff, mode in ((True, 2), (False, 1)):
        g = _gpu(dg, dg.models.quad_decay_src, alg, u0, p, [0, 10], adaptive=True, autodiff=autodiff, **gkw)
        _same(g, oracle.solve("quad_decay", alg, u0, p, [0, 10], jac_mode=mode, **akw), f"quad_decay mode {mode}")
    # Robertson parameter sweep (3x3 closed-form solves inside the Newton iterations); AOT kernels
    k = (np.array([0.04, 3e7, 1e4]) * (0.5 + np.random.default_rng(2).random((256, 3)))).astype(f32)
    g = _gpu(dg, dg.models.rober, alg, [1, 0, 0], k, [0, 1e3], adaptive=True, dt=f32(1e-4), abstol=f32(1e-8), reltol=f32(1e-4),
             save_everystep=False)
    r = oracle.solve("rober", alg, [1, 0, 0], k, [0, 1e3], dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4, save_everystep=False)
    _same(g, r, "rober sweep")
    assert (g["retcode"] == 1).all() and np.abs(g["us"][:, 1].sum(axis=1) - 1).max() < 1e-4
    # the fast build (packed-free, FMA contraction) stays within tolerance of the strict one
    gf = _gpu(dg, dg.models.rober, alg, [1, 0, 0], k, [0, 1e3], adaptive=True, fp_mode="fast", dt=f32(1e-4), abstol=f32(1e-8),
              reltol=f32(1e-4), save_everystep=False)
    assert (gf["retcode"] == 1).all() and np.abs(gf["us"][:, 1] - g["us"][:, 1]).max() < 2e-3
    # 15 states: general LU (JIT-only model)
    u15 = np.random.default_rng(1).random((40, 15)).astype(f32)
    g = _gpu(dg, dg.models.linear15, alg, u15, None, [0, 1], adaptive=True, dt=f32(0.01), abstol=f32(1e-6), reltol=f32(1e-6),
             save_everystep=False)
    r = oracle.solve("linear15", alg, u15, None, [0, 1], dt=0.01, adaptive=True, abstol=1e-6, reltol=1e-6, save_everystep=False)
    _same(g, r, "linear15")
    # saveat (gpu_ode_regression.jl:62-118): default Hermite interpolant, fixed dt and adaptive (deferred-save replay)
    for sv in (np.array([2.0, 4.0], f32), np.arange(0, 101, dtype=f32) * f32(0.1)):
        g = _gpu(dg, dg.models.decay, alg, u0, p, [0, 10], adaptive=False, dt=f32(0.01), saveat=sv)
        _same(g, oracle.solve("decay", alg, u0, p, [0, 10], dt=0.01, saveat=sv), f"fixed saveat {len(sv)}")
        g = _gpu(dg, dg.models.decay, alg, u0, p, [0, 10], adaptive=True, dt=f32(0.01), abstol=f32(1e-7), reltol=f32(1e-7), saveat=sv)
        _same(g, oracle.solve("decay", alg, u0, p, [0, 10], dt=0.01, adaptive=True, abstol=1e-7, reltol=1e-7, saveat=sv),
              f"adaptive saveat {len(sv)}")
        assert np.abs(g["us"][:, :, 0] - u0 * np.exp(-p * sv[None].astype(f64))).max() < 3e-2
    rsv = np.array([1.0, 10.0, 100.0, 1e3], f32)
    g = _gpu(dg, dg.models.rober, alg, [1, 0, 0], k, [0, 1e3], adaptive=True, dt=f32(1e-4), abstol=f32(1e-8), reltol=f32(1e-4), saveat=rsv)
    r = oracle.solve("rober", alg, [1, 0, 0], k, [0, 1e3], dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4, saveat=rsv)
    _same(g, r, "rober saveat")
    gf = _gpu(dg, dg.models.rober, alg, [1, 0, 0], k, [0, 1e3], adaptive=True, fp_mode="fast", dt=f32(1e-4), abstol=f32(1e-8),
              reltol=f32(1e-4), saveat=rsv)
    assert (gf["retcode"] == 1).all() and np.abs(gf["us"] - g["us"]).max() < 2e-3
