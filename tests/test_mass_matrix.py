"""Constant mass matrix M u' = f(u, p, t) with the Rosenbrock family: Robertson in DAE
form, M = diag(1, 1, 0), as in test/gpu_kernel_de/stiff_ode/gpu_ode_mass_matrix.jl (GPURosenbrock23 there;
GPURodas4 / GPURodas5P carry the same `mass_matrix - gamma*J` W and `mass_matrix * (C-sum)` terms)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

f32, f64 = np.float32, np.float64
K0 = np.array([0.04, 3e7, 1e4])
REF_END = np.array([1.78659e-2, 7.27475e-8, 9.82134e-1])     # Robertson at t = 1e5 (Radau / literature)


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle
    return oracle


def test_oracle_robertson_dae(oracle):
    # gpu_ode_mass_matrix.jl:44-58: dt = 0.1f0, abstol = reltol = 1f-5, `norm(bench - sol) < 8e-4`
    k = K0[None].astype(f32)
    r = oracle.solve("rober_dae", "rosenbrock23", [1, 0, 0], k, [0, 1e5], dt=0.1, adaptive=True, abstol=1e-5, reltol=1e-5,
                     save_everystep=False)
    assert r["retcode"][0] == 1 and np.linalg.norm(r["us"][0, 1] - REF_END) < 8e-4
    assert abs(r["us"][0, 1].sum() - 1) < 1e-5                     # the algebraic row holds the constraint
    # the DAE form and the ODE form describe the same solution
    o = oracle.solve("rober", "rosenbrock23", [1, 0, 0], k, [0, 1e5], dt=0.1, adaptive=True, abstol=1e-5, reltol=1e-5,
                     save_everystep=False)
    assert np.linalg.norm(r["us"][0, 1] - o["us"][0, 1]) < 8e-4
    # Float64, tight tolerance, saveat through the Rosenbrock23 interpolant
    sv = np.array([1.0, 1e2, 1e4, 1e5])
    r = oracle.solve("rober_dae", "rosenbrock23", [1, 0, 0], K0[None], [0, 1e5], dt=0.1, adaptive=True, abstol=1e-9, reltol=1e-8,
                     saveat=sv, dtype=f64)
    assert np.allclose(r["us"][0, -1], REF_END, rtol=2e-4) and np.abs(r["us"][0].sum(axis=1) - 1).max() < 1e-9
    # ForwardDiff duals give the same Jacobian as the analytic body (the test itself passes no jac)
    a = oracle.solve("rober_dae", "rosenbrock23", [1, 0, 0], k, [0, 1e3], dt=0.1, adaptive=True, abstol=1e-5, reltol=1e-5, save_everystep=False)
    b = oracle.solve("rober_dae", "rosenbrock23", [1, 0, 0], k, [0, 1e3], dt=0.1, adaptive=True, abstol=1e-5, reltol=1e-5, save_everystep=False, jac_mode=2)
    assert np.array_equal(a["us"], b["us"])
    # explicit solvers: no mass matrix
    for alg in ("tsit5", "vern9"):
        with pytest.raises(RuntimeError):
            oracle.solve("rober_dae", alg, [1, 0, 0], k, [0, 1.0], dt=0.1, adaptive=True, save_everystep=False)


STIFF = ["rosenbrock23", "rodas4", "rodas5p", "kvaerno3", "kvaerno5"]


@pytest.mark.parametrize("alg", STIFF)
def test_oracle_direct_mass_matrix_dae(oracle, alg):
    """stiff_ode/gpu_ode_modelingtoolkit_dae.jl:22-112 ("Direct mass matrix DAE"): M = diag(1, 0), fixed dt = 0.001 on
    (0, 0.1) with all five stiff solvers; `!any(isnan, u_end)` and `|u1 + u2 - 1| < 0.01`.  The Kvaerno steppers
    see the mass matrix inside their Newton iteration only (nlsolve/utils.jl:10-21)."""
    r = oracle.solve("lin_dae", alg, [1.0, 0.0], [0.04, 1e4], [0, 0.1], dt=0.001, length=102)
    ts, us = r["ts"][0], r["us"][0]
    last = np.nonzero(ts != 0)[0][-1]
    assert r["retcode"][0] == 1 and ts[last] == f32(0.1) and not np.isnan(us[:last + 1]).any()
    assert abs(us[last].sum() - 1) < 0.01
    # beyond the reference's bound: the differential state relaxes to k2 / (k1 + k2) within a few steps
    assert abs(us[last, 0] - 1e4 / (1e4 + 0.04)) < 1e-6 and abs(us[last, 1] - 0.04 / (1e4 + 0.04)) < 1e-7


@pytest.mark.parametrize("alg", ["rodas4", "rodas5p"])
def test_oracle_robertson_dae_rodas(oracle, alg):
    # Float64, tight tolerance: the DAE form lands on the literature value and holds the constraint to rounding;
    # it differs from the ODE form only at rounding level (same steps), which shows the mass terms are exercised
    kw = dict(dt=1e-4, adaptive=True, abstol=1e-9, reltol=1e-8, save_everystep=False, dtype=f64)
    r = oracle.solve("rober_dae", alg, [1, 0, 0], K0[None], [0, 1e5], **kw)
    o = oracle.solve("rober", alg, [1, 0, 0], K0[None], [0, 1e5], **kw)
    assert r["retcode"][0] == 1 and np.allclose(r["us"][0, 1], REF_END, rtol=2e-5)
    assert abs(r["us"][0, 1].sum() - 1) < 1e-14
    assert np.allclose(r["us"][0, 1], o["us"][0, 1], rtol=1e-7) and not np.array_equal(r["us"][0, 1], o["us"][0, 1])
    # Float32 at the reference test's tolerances
    r = oracle.solve("rober_dae", alg, [1, 0, 0], K0[None].astype(f32), [0, 1e5], dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4,
                     save_everystep=False)
    assert r["retcode"][0] == 1 and np.linalg.norm(r["us"][0, 1] - REF_END) < 8e-4 and abs(r["us"][0, 1].sum() - 1) < 1e-5
    # analytic Jacobian and duals agree bit for bit
    b = oracle.solve("rober_dae", alg, [1, 0, 0], K0[None].astype(f32), [0, 1e5], dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4,
                     save_everystep=False, jac_mode=2)
    assert np.array_equal(r["us"], b["us"])


def test_mass_matrix_lowering_compiles():
    import diffeqgpu_b200 as dg
    from diffeqgpu_b200 import _lib
    for fp in (_lib.FP_STRICT, _lib.FP_FAST):
        for dt in (_lib.F32, _lib.F64):
            d = _lib.make_desc(rhs_src=dg.models.ROBER_DAE_RHS, mass_src="Mm[0][0] = (T)1; Mm[1][1] = (T)1;", n_state=3, n_param=3,
                               dtype=dt, alg=3, fp_mode=fp)       # no jac body: forward-mode duals
            st, nb, log = _lib.jit_compile_check(d)
            assert st == 0 and nb > 0, log
    for alg in (4, 5):                                            # GPURodas4 / GPURodas5P
        for fp in (_lib.FP_STRICT, _lib.FP_FAST):
            d = _lib.make_desc(rhs_src=dg.models.ROBER_DAE_RHS, jac_src=dg.models.ROBER_DAE_JAC, mass_src="Mm[0][0] = (T)1; Mm[1][1] = (T)1;",
                               n_state=3, n_param=3, dtype=_lib.F32, alg=alg, fp_mode=fp)
            st, nb, log = _lib.jit_compile_check(d)
            assert st == 0 and nb > 0, log
    d = _lib.make_desc(rhs_src=dg.models.LIN_DAE_RHS, jac_src=dg.models.LIN_DAE_JAC, mass_src="Mm[0][0] = (T)1;", n_state=2, n_param=2,
                       dtype=_lib.F32, alg=9)                     # GPUKvaerno5: M inside the Newton iteration
    st, nb, log = _lib.jit_compile_check(d)
    assert st == 0 and nb > 0, log
    for alg in (0, 2):                                            # GPUTsit5, GPUVern9: refused, user bodies and built-in
        st, _, log = _lib.jit_compile_check(_lib.make_desc(rhs_src=dg.models.ROBER_DAE_RHS, mass_src="Mm[0][0] = (T)1;", n_state=3,
                                                           n_param=3, dtype=_lib.F32, alg=alg))
        assert st == _lib.ERR_UNSUPPORTED and "implicit solver" in log
        st, _, log = _lib.jit_compile_check(_lib.make_desc(builtin="rober_dae", dtype=_lib.F32, alg=alg, force_jit=True))
        assert st == _lib.ERR_UNSUPPORTED and "implicit solver" in log


@pytest.mark.gpu
def test_gpu_robertson_dae_bit_exact(oracle):
    import dataclasses
    import torch
    import diffeqgpu_b200 as dg
    k = (K0 * (0.5 + np.random.default_rng(4).random((300, 3)))).astype(f32)
    sv = np.array([1.0, 1e2, 1e3, 1e4], f32)

    def gpu(func, autodiff=True, fp_mode="strict", **kw):
        prob = dg.ODEProblem(func, np.array([1, 0, 0], f32), (0.0, 1e4), k[0])
        probs = dg.ProblemBatch.from_arrays(prob, p=k, device="cuda:0")
        ts, us, st = dg.vectorized_asolve(probs, prob, dg.GPURosenbrock23(autodiff=autodiff), dt=f32(0.1), abstol=f32(1e-5),
                                          reltol=f32(1e-5), fp_mode=fp_mode, stats=True, **kw)
        torch.cuda.synchronize()
        return dict(ts=ts.cpu().numpy(), us=us.cpu().numpy(), naccept=st["naccept"].cpu().numpy(),
                    nreject=st["nreject"].cpu().numpy(), retcode=st["retcode"].cpu().numpy())

    okw = dict(dt=0.1, adaptive=True, abstol=1e-5, reltol=1e-5)
    for kw in (dict(save_everystep=False), dict(saveat=sv)):
        r = oracle.solve("rober_dae", "rosenbrock23", [1, 0, 0], k, [0, 1e4], **okw, **kw)
        for func in (dg.models.rober_dae, dg.models.rober_dae_src):          # built-in struct and user bodies
            g = gpu(func, **kw)
            for key in ("ts", "naccept", "nreject", "retcode"):
                assert np.array_equal(g[key], r[key], equal_nan=True), (key, sorted(kw))
            w = g["ts"] != 0           # rows a failed trajectory never reached are uninitialised on the device
            w[:, 0] = True if "saveat" not in kw else w[:, 0]
            assert np.array_equal(g["us"][w], r["us"][w], equal_nan=True), ("us", sorted(kw))
        # no Jacobian body (what the reference test does): duals, same bits
        g = gpu(dataclasses.replace(dg.models.rober_dae_src, jac=None), **kw)
        assert np.array_equal(g["us"][w], r["us"][w], equal_nan=True)
    # (in Float32 at tol 1e-5 the method itself loses a few of these stiff DAE trajectories -- oracle and
    #  device agree on them bit for bit -- so the invariant is checked on the bulk of the sweep)
    good = np.abs(g["us"].sum(axis=2) - 1).max(axis=1) < 1e-4
    assert np.isin(g["retcode"], (1, 2)).all() and (g["retcode"] == 1).mean() > 0.95 and good.mean() > 0.95
    # fast build (packed pairs): same solution within tolerance
    gf = gpu(dg.models.rober_dae, fp_mode="fast", saveat=sv)
    gf_good = np.abs(gf["us"].sum(axis=2) - 1).max(axis=1) < 1e-4
    both = good & gf_good
    assert both.mean() > 0.9 and np.abs(gf["us"][both] - g["us"][both]).max() < 2e-3
    assert np.isin(gf["retcode"], (1, 2, 3)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("alg", ["rodas4", "rodas5p"])
def test_gpu_robertson_dae_rodas_bit_exact(oracle, alg):
    import torch
    import diffeqgpu_b200 as dg
    k = (K0 * (0.5 + np.random.default_rng(5).random((200, 3)))).astype(f32)
    sv = np.array([1.0, 1e2, 1e3, 1e4], f32)
    Alg = dict(rodas4=dg.GPURodas4, rodas5p=dg.GPURodas5P)[alg]

    def gpu(func, fp_mode="strict", **kw):
        prob = dg.ODEProblem(func, np.array([1, 0, 0], f32), (0.0, 1e4), k[0])
        probs = dg.ProblemBatch.from_arrays(prob, p=k, device="cuda:0")
        ts, us, st = dg.vectorized_asolve(probs, prob, Alg(), dt=f32(1e-4), abstol=f32(1e-8), reltol=f32(1e-4),
                                          fp_mode=fp_mode, stats=True, **kw)
        torch.cuda.synchronize()
        return dict(ts=ts.cpu().numpy(), us=us.cpu().numpy(), naccept=st["naccept"].cpu().numpy(),
                    nreject=st["nreject"].cpu().numpy(), retcode=st["retcode"].cpu().numpy())

    okw = dict(dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4)
    for kw in (dict(save_everystep=False), dict(saveat=sv)):
        r = oracle.solve("rober_dae", alg, [1, 0, 0], k, [0, 1e4], **okw, **kw)
        for func in (dg.models.rober_dae, dg.models.rober_dae_src):          # built-in struct and user bodies
            g = gpu(func, **kw)
            for key in ("ts", "naccept", "nreject", "retcode"):
                assert np.array_equal(g[key], r[key], equal_nan=True), (key, sorted(kw))
            w = g["ts"] != 0
            w[:, 0] = True if "saveat" not in kw else w[:, 0]
            assert np.array_equal(g["us"][w], r["us"][w], equal_nan=True), ("us", sorted(kw))
    ok = g["retcode"] == 1
    assert ok.mean() > 0.95 and np.abs(g["us"][ok].sum(axis=2) - 1).max() < 1e-4
    # fast build (packed pairs): same solution within tolerance
    gf = gpu(dg.models.rober_dae, fp_mode="fast", saveat=sv)
    both = ok & (gf["retcode"] == 1)
    assert both.mean() > 0.9 and np.abs(gf["us"][both] - g["us"][both]).max() < 2e-3
    assert np.abs(gf["us"][both].sum(axis=2) - 1).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("alg", STIFF)
def test_gpu_direct_mass_matrix_dae_bit_exact(oracle, alg):
    """the reference's "Direct mass matrix DAE" cases (fixed dt, five stiff solvers) on a parameter sweep"""
    import torch
    import diffeqgpu_b200 as dg
    n = 100
    rng = np.random.default_rng(8)
    p = (np.array([0.04, 1e4]) * (0.5 + rng.random((n, 2)))).astype(f32)
    Alg = dict(rosenbrock23=dg.GPURosenbrock23, rodas4=dg.GPURodas4, rodas5p=dg.GPURodas5P, kvaerno3=dg.GPUKvaerno3,
               kvaerno5=dg.GPUKvaerno5)[alg]
    prob = dg.ODEProblem(dg.models.lin_dae_src, np.array([1, 0], f32), (0.0, 0.1), p[0])
    probs = dg.ProblemBatch.from_arrays(prob, p=p, device="cuda:0")
    for kw in (dict(), dict(save_everystep=False), dict(saveat=np.array([0.0, 0.05, 0.1], f32))):
        ts, us, st = dg.vectorized_solve(probs, prob, Alg(), dt=f32(0.001), stats=True, **kw)
        torch.cuda.synchronize()
        g_ts, g_us = ts.cpu().numpy(), us.cpu().numpy()
        okw = dict(kw) if kw else dict(length=g_us.shape[1])
        r = oracle.solve("lin_dae", alg, [1.0, 0.0], p, [0, 0.1], dt=0.001, **okw)
        assert np.array_equal(g_ts, r["ts"]) and np.array_equal(g_us, r["us"], equal_nan=True), sorted(kw)
        assert (st["retcode"].cpu().numpy() == 1).all()
    assert np.abs(g_us[:, -1].sum(axis=1) - 1).max() < 0.01 and not np.isnan(g_us).any()


# ------------------------------------------------------------------------------------------
# DAE initialisation (reference kernels.jl:19-25, 93-99 -> nlsolve/initialization.jl; `ODEFunction(..., initialize=True)`)
# ------------------------------------------------------------------------------------------
def test_dae_init_restatement_solves_the_algebraic_rows():
    from oracle.dae_init import dae_initialize
    M = np.diag([1.0, 0.0])
    f = lambda u, p, t: np.array([-u[0], u[1] * u[1] + u[0] - p[0]])
    jac = lambda u, p, t: np.array([[-1.0, 0.0], [1.0, 2 * u[1]]])
    u, ok = dae_initialize(f, jac, M, [1.0, 3.0], [2.0], 0.0, dtype=f64)
    assert ok and u[0] == 1.0 and abs(u[1] - 1.0) < 1e-6
    u, ok = dae_initialize(f, jac, M, [1.0, 3.0], [-5.0], 0.0, dtype=f64)     # u1^2 = -6: no consistent value
    assert not ok
    u, ok = dae_initialize(f, jac, np.eye(2), [1.0, 3.0], [2.0], 0.0)          # no algebraic rows: untouched
    assert ok and np.array_equal(u, np.array([1.0, 3.0], f32))


@pytest.mark.gpu
@pytest.mark.parametrize("adaptive", [False, True])
def test_gpu_dae_initialisation(adaptive):
    import torch
    import diffeqgpu_b200 as dg
    from oracle.dae_init import dae_initialize
    rhs = "    du[0] = -u[0];\n    du[1] = u[1] * u[1] + u[0] - p[0];\n"
    jac = "    J[0][0] = (T)-1;\n    J[1][0] = (T)1;  J[1][1] = (T)2 * u[1];\n"
    f = dg.ODEFunction(rhs=rhs, jac=jac, mass_matrix="    Mm[0][0] = (T)1;\n", n_state=2, n_param=1, initialize=True)
    n = 257
    ps = np.linspace(1.5, 6.0, n).astype(f32)[:, None]
    ps[5, 0] = -5.0                                        # no consistent value: u1^2 = p - u0 < 0
    prob = dg.ODEProblem(f, np.array([1.0, 3.0], f32), (0.0, 1.0), np.array([2.0], f32))
    probs = dg.ProblemBatch.from_arrays(prob, p=ps, device="cuda:0")
    sv = np.array([0.0, 0.5, 1.0], f32)
    for alg in (dg.GPURosenbrock23(), dg.GPURodas5P()):
        if adaptive:
            ts, us, st = dg.vectorized_asolve(probs, prob, alg, dt=f32(0.01), saveat=sv, abstol=f32(1e-6), reltol=f32(1e-6), stats=True)
        else:
            ts, us, st = dg.vectorized_solve(probs, prob, alg, dt=f32(0.01), saveat=sv, stats=True)
        torch.cuda.synchronize()
        us, ts, rc = us.cpu().numpy(), ts.cpu().numpy(), st["retcode"].cpu().numpy()
        ok = np.ones(n, bool); ok[5] = False
        assert (rc[ok] == 1).all() and rc[5] == 7          # InitialFailure: not integrated, row 1 = prob.u0
        assert np.array_equal(us[5, 0], np.array([1.0, 3.0], f32)) and ts[5, 0] == 0 and (ts[5, 1:] == 0).all()
        fpy = lambda u, p, t: np.array([-u[0], u[1] * u[1] + u[0] - p[0]])
        jpy = lambda u, p, t: np.array([[-1.0, 0.0], [1.0, 2 * u[1]]])
        for i in (0, 17, 100, 256):
            u_ref, ok_ref = dae_initialize(fpy, jpy, np.diag([1.0, 0.0]), [1.0, 3.0], ps[i], 0.0)
            assert ok_ref and np.abs(us[i, 0] - u_ref).max() < 2e-6          # row 1 holds the initialised state
        # the algebraic row holds along the whole solution: u1^2 + u0 = p, u0 = exp(-t)
        res = us[ok, :, 1] ** 2 + us[ok, :, 0] - ps[ok]
        assert np.abs(res).max() < 2e-4
        assert np.abs(us[ok, -1, 0] - np.exp(-1.0)).max() < (2e-3 if not adaptive else 1e-3)
    # without `initialize` the inconsistent u0 is integrated as given (row 1 = prob.u0)
    f0 = dg.ODEFunction(rhs=rhs, jac=jac, mass_matrix="    Mm[0][0] = (T)1;\n", n_state=2, n_param=1)
    prob0 = dg.ODEProblem(f0, np.array([1.0, 3.0], f32), (0.0, 1.0), np.array([2.0], f32))
    ts, us = dg.vectorized_solve(dg.ProblemBatch.from_arrays(prob0, p=ps[:4], device="cuda:0"), prob0, dg.GPURosenbrock23(), dt=f32(0.01), saveat=sv)
    torch.cuda.synchronize()
    assert np.array_equal(us.cpu().numpy()[:, 0], np.tile(np.array([1.0, 3.0], f32), (4, 1)))
