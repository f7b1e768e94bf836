"""Seeded inputs shared by the oracle tests, the golden-vector generator and the GPU parity tests."""
import numpy as np

P0_LORENZ = np.array([10.0, 28.0, 8.0 / 3.0])
K0_ROBER = np.array([0.04, 3.0e7, 1.0e4])


def lorenz_sweep(n, dtype=np.float32, seed=20240607):
    """p_i = r_i .* (10, 28, 8/3), r ~ U[0,1)^3 in Float32 (README.md:74 of the reference)"""
    r = np.random.default_rng(seed).random((n, 3), dtype=np.float32)
    return (r * P0_LORENZ.astype(np.float32)).astype(dtype)


def rober_sweep(n, dtype=np.float32, seed=7):
    r = np.random.default_rng(seed).random((n, 3))
    return (K0_ROBER * (0.5 + r)).astype(dtype)


def henon_heiles_u0(n, dtype=np.float64, seed=11):
    """E = 1/8, x = 0, y = -0.1 + 0.3 r1, py = -0.1 + 0.2 r2"""
    r = np.random.default_rng(seed).random((n, 2))
    y = -0.1 + 0.3 * r[:, 0]
    py = -0.1 + 0.2 * r[:, 1]
    px = np.sqrt(2 * 0.125 - py ** 2 - y ** 2 + (2.0 / 3.0) * y ** 3)
    return np.stack([np.zeros(n), y, px, py], axis=1).astype(dtype)


U0_LORENZ = np.array([1.0, 0.0, 0.0])

# (name, kwargs for oracle.solve) -- small cases whose oracle outputs are committed as goldens
def golden_cases():
    f32, f64 = np.float32, np.float64
    sv = np.arange(0, 11.0)
    cases = []
    for alg in ("tsit5", "vern7", "vern9"):
        cases.append((f"lorenz_{alg}_fixed_f32", dict(model="lorenz", alg=alg, u0=U0_LORENZ, p=lorenz_sweep(16), tspan=[0, 10], dt=0.1, length=101, dtype=f32)))
        cases.append((f"lorenz_{alg}_adaptive_saveat_f32", dict(model="lorenz", alg=alg, u0=U0_LORENZ, p=lorenz_sweep(16), tspan=[0, 10], dt=0.1, adaptive=True, abstol=1e-6, reltol=1e-6, saveat=sv, dtype=f32)))
        cases.append((f"lorenz_{alg}_adaptive_endpoints_f64", dict(model="lorenz", alg=alg, u0=U0_LORENZ, p=lorenz_sweep(16, f64), tspan=[0, 10], dt=0.1, adaptive=True, abstol=1e-10, reltol=1e-10, save_everystep=False, dtype=f64)))
    cases.append(("hh_vern9_adaptive_f64", dict(model="henon_heiles", alg="vern9", u0=henon_heiles_u0(16), p=None, tspan=[0, 100], dt=0.1, adaptive=True, abstol=1e-10, reltol=1e-10, save_everystep=False, dtype=f64)))
    for alg in ("rosenbrock23", "rodas4", "rodas5p"):
        cases.append((f"rober_{alg}_adaptive_f32", dict(model="rober", alg=alg, u0=U0_LORENZ, p=rober_sweep(16), tspan=[0, 1e5], dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4, saveat=[1.0, 10.0, 1e3, 1e5], dtype=f32)))
        cases.append((f"decay_{alg}_fixed_f32", dict(model="decay", alg=alg, u0=[10.0], p=[1.0], tspan=[0, 10], dt=0.01, length=1001, dtype=f32)))
        cases.append((f"lorenz_{alg}_adaptive_f64", dict(model="lorenz", alg=alg, u0=U0_LORENZ, p=lorenz_sweep(8, f64), tspan=[0, 2], dt=0.01, adaptive=True, abstol=1e-8, reltol=1e-8, saveat=[0.5, 1.0, 2.0], dtype=f64)))
    for alg in ("em", "siea"):
        cases.append((f"gbm_{alg}_f32", dict(model="gbm", alg=alg, u0=np.full((32, 3), 0.1), p=[1.5, 0.01], tspan=[0, 1], dt=1 / 64, save_everystep=False, seed=1234, dtype=f32)))
    cases.append(("lorenz_additive_em_saveat_f32", dict(model="lorenz_additive", alg="em", u0=U0_LORENZ, p=lorenz_sweep(8), tspan=[0, 1], dt=1e-3, saveat=[0.0, 0.25, 0.5, 1.0], seed=7, dtype=f32)))
    return cases


# cases frozen for the oracle only (events, mass matrices, the ESDIRK steppers): the GPU tests of these features
# compare the device with a fresh oracle run on larger sweeps (test_events / test_mass_matrix / test_kvaerno)
def golden_cases_oracle_only():
    f32 = np.float32
    kick = (("t_eq", 0, 2.4), ("u_add", 0, 10.0))
    bounce = dict(condition=("u_minus", 0, 0.0), affect=("u_scale", 1, -1.0))
    sv = np.array([2.0, 4.0])
    cases = []
    for alg in ("kvaerno3", "kvaerno5"):
        cases.append((f"decay_{alg}_fixed_saveat_f32", dict(model="decay", alg=alg, u0=[10.0], p=[1.0], tspan=[0, 10], dt=0.01, saveat=sv, dtype=f32)))
        cases.append((f"decay_{alg}_adaptive_saveat_f32", dict(model="decay", alg=alg, u0=[10.0], p=[1.0], tspan=[0, 10], dt=0.01, adaptive=True, abstol=1e-7, reltol=1e-7, saveat=sv, dtype=f32)))
        cases.append((f"decay_{alg}_fixed_tstops_kick_f32", dict(model="decay", alg=alg, u0=[10.0], p=[1.0], tspan=[0, 10], dt=0.5, length=22, tstops=[2.4], callbacks=[kick], dtype=f32)))
    for alg in ("rosenbrock23", "rodas4", "rodas5p", "kvaerno3", "kvaerno5"):
        cases.append((f"lin_dae_{alg}_fixed_f32", dict(model="lin_dae", alg=alg, u0=[1.0, 0.0], p=[0.04, 1e4], tspan=[0, 0.1], dt=0.001, length=102, dtype=f32)))
    # stiff_ode/gpu_ode_mass_matrix.jl:44-58 (GPURosenbrock23, dt = 0.1, tol 1e-5) and the Rodas steppers on a sweep
    cases.append(("rober_dae_rosenbrock23_adaptive_f32", dict(model="rober_dae", alg="rosenbrock23", u0=U0_LORENZ, p=[0.04, 3e7, 1e4], tspan=[0, 1e5], dt=0.1, adaptive=True, abstol=1e-5, reltol=1e-5, save_everystep=False, dtype=f32)))
    for alg in ("rodas4", "rodas5p"):
        cases.append((f"rober_dae_{alg}_adaptive_f32", dict(model="rober_dae", alg=alg, u0=U0_LORENZ, p=rober_sweep(8), tspan=[0, 1e4], dt=1e-4, adaptive=True, abstol=1e-8, reltol=1e-4, saveat=[1.0, 1e2, 1e3, 1e4], dtype=f32)))
    for alg in ("tsit5", "rosenbrock23", "rodas4", "kvaerno3"):
        cases.append((f"ball_{alg}_fixed_saveat_f32", dict(model="ball", alg=alg, u0=[45.0, 0.0], p=[10.0], tspan=[0, 10], dt=0.1, saveat=[0.0, 4.3, 9.1], continuous_callbacks=[bounce], dtype=f32)))
        cases.append((f"ball_{alg}_adaptive_f32", dict(model="ball", alg=alg, u0=[45.0, 0.0], p=[10.0], tspan=[0, 16.5], dt=0.1, adaptive=True, abstol=1e-6, reltol=1e-6, save_everystep=False, continuous_callbacks=[bounce], dtype=f32)))
    return cases


# ------------------------------------------------------------------------------------------
# discrete-callback specs: one description, two lowerings -- the oracle interprets the spec
# (oracle.COND_KINDS / AFFECT_KINDS), the device gets CUDA-C bodies (include/degk.h,
# degk_model_desc.cb_condition_src / cb_affect_src)
# ------------------------------------------------------------------------------------------
def _lit(v, dtype):
    return f"(T){float(np.dtype(dtype).type(v))!r}"


def callback_sources(spec, dtype=np.float32):
    """spec: ((cond_kind, idx, val), (affect_kind, idx, val)) -> (condition_src, affect_src)"""
    (ck, ci, cv), (ak, ai, av) = spec
    cond = {"t_eq": f"return t == {_lit(cv, dtype)};", "u_lt": f"return u[{ci}] < {_lit(cv, dtype)};",
            "u_gt": f"return u[{ci}] > {_lit(cv, dtype)};", "t_ge": f"return t >= {_lit(cv, dtype)};"}[ck]
    aff = {"u_add": f"u[{ai}] = u[{ai}] + {_lit(av, dtype)};", "u_set": f"u[{ai}] = {_lit(av, dtype)};",
           "u_scale": f"u[{ai}] = u[{ai}] * {_lit(av, dtype)};", "terminate": "terminate();",
           "p_set": f"p[{ai}] = {_lit(av, dtype)};"}[ak]
    return cond, aff


def continuous_callback_sources(cc, dtype=np.float32):
    """cc: oracle-style dict(condition=(kind, idx, val), affect=..., affect_neg=..., rootfind, ...) ->
    kwargs of dg.ContinuousCallback (condition / affect bodies in CUDA-C)"""
    ck, ci, cv = cc["condition"]
    cond = {"u_minus": f"return u[{ci}] - {_lit(cv, dtype)};", "t_minus": f"return t - {_lit(cv, dtype)};"}[ck]

    def aff(a):
        if a is None:
            return None
        return callback_sources((("t_eq", 0, 0.0), a), dtype)[1]
    affect = aff(cc.get("affect"))
    neg = cc.get("affect_neg", "same")
    kw = dict(affect_neg="same" if neg == "same" else aff(neg), rootfind=cc.get("rootfind", "left"))
    for k in ("abstol", "repeat_nudge", "dtrelax"):
        if k in cc:
            kw[k] = cc[k]
    return cond, affect, kw
