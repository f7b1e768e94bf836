"""The CPU oracle against (a) the committed golden vectors, (b) the reference tests' own
assertions with scipy standing in for OrdinaryDiffEq, (c) the quirks listed in SURVEY §8."""
import sys
from pathlib import Path

import numpy as np
import pytest
from scipy.integrate import solve_ivp

sys.path.insert(0, str(Path(__file__).resolve().parent))
from cases import K0_ROBER, P0_LORENZ, U0_LORENZ, golden_cases, golden_cases_oracle_only, lorenz_sweep  # noqa: E402
from oracle import oracle  # noqa: E402

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "oracle_golden.npz")


def lorenz_rhs(t, u, s, r, b):
    return [s * (u[1] - u[0]), u[0] * (r - u[2]) - u[1], u[0] * u[1] - b * u[2]]


def truth_lorenz(tf, p=P0_LORENZ, t_eval=None):
    s = solve_ivp(lorenz_rhs, [0, tf], U0_LORENZ, args=tuple(p), rtol=1e-13, atol=1e-13,
                  method="DOP853", t_eval=t_eval)
    return s.y.T


@pytest.mark.parametrize("name,kw", golden_cases() + golden_cases_oracle_only(),
                         ids=[c[0] for c in golden_cases() + golden_cases_oracle_only()])
def test_oracle_matches_golden(name, kw):
    kw = dict(kw)
    model, alg = kw.pop("model"), kw.pop("alg")
    r = oracle.solve(model, alg, kw.pop("u0"), kw.pop("p"), kw.pop("tspan"), **kw)
    for k in ("ts", "us", "naccept", "nreject", "retcode"):
        assert np.array_equal(r[k], GOLD[f"{name}/{k}"], equal_nan=True), (name, k)


# ---- reference test/gpu_kernel_de/gpu_ode_regression.jl:19-111 ----
@pytest.mark.parametrize("alg", ["tsit5", "vern7", "vern9"])
def test_lorenz_regression_like_reference(alg):
    p = P0_LORENZ.astype(np.float32)
    sol = oracle.solve("lorenz", alg, U0_LORENZ, p, [0, 10], dt=0.01, length=1001)
    asol = oracle.solve("lorenz", alg, U0_LORENZ, p, [0, 10], dt=0.01, adaptive=True, abstol=1e-7,
                        reltol=1e-7, save_everystep=False)
    tr = truth_lorenz(10.0, p.astype(np.float64))[-1]
    assert np.linalg.norm(sol["us"][0, -1] - tr) < 1e-2        # :45 (5e-3 vs Float32 Vern9)
    assert np.linalg.norm(asol["us"][0, 1] - tr) < 5e-3        # :46
    assert np.all(sol["us"][0, 0] == U0_LORENZ) and asol["ts"][0, 1] == 10.0
    # saveat vector (:50-72): values on the grid, first point = first saveat
    sv = np.array([2.0, 4.0], np.float32)
    s2 = oracle.solve("lorenz", alg, U0_LORENZ, p, [0, 10], dt=0.01, saveat=sv)
    a2 = oracle.solve("lorenz", alg, U0_LORENZ, p, [0, 10], dt=0.01, adaptive=True, abstol=1e-7, reltol=1e-7, saveat=sv)
    trs = truth_lorenz(10.0, p.astype(np.float64), t_eval=[2.0, 4.0])
    assert np.array_equal(s2["ts"][0], sv) and np.array_equal(a2["ts"][0], sv)
    assert np.linalg.norm(s2["us"][0] - trs) < 1e-2 and np.linalg.norm(a2["us"][0] - trs) < 5e-3
    # save_everystep = false -> length 2 (:113)
    e = oracle.solve("lorenz", alg, U0_LORENZ, p, [0, 10], dt=0.01, save_everystep=False)
    assert e["us"].shape == (1, 2, 3)


@pytest.mark.parametrize("alg,order", [("tsit5", 5), ("vern7", 7), ("vern9", 9)])
def test_convergence_order_f64(alg, order):
    errs = []
    tr = truth_lorenz(1.0)[-1]
    dts = [0.05, 0.025] if order < 9 else [0.1, 0.05]
    for dt in dts:
        n = int(round(1.0 / dt))
        r = oracle.solve("lorenz", alg, U0_LORENZ, P0_LORENZ, [0, 1], dt=dt, length=n + 1, dtype=np.float64)
        errs.append(np.linalg.norm(r["us"][0, n] - tr))
    obs = np.log2(errs[0] / errs[1])
    assert obs > order - 0.8, (alg, errs, obs)


@pytest.mark.parametrize("alg", ["tsit5", "vern7", "vern9"])
def test_dense_output_accuracy_f64(alg):
    sv = np.linspace(0.05, 1.95, 13)
    r = oracle.solve("lorenz", alg, U0_LORENZ, P0_LORENZ, [0, 2], dt=0.01, adaptive=True, abstol=1e-10,
                     reltol=1e-10, saveat=sv, dtype=np.float64)
    tr = truth_lorenz(2.0, t_eval=sv)
    assert np.abs(r["us"][0] - tr).max() < 2e-7


def test_quirk_q1_time_accumulates_in_float32():
    """C1: 100 steps of 0.1f0 end at 10.000002f0 > tf; the last row is then overwritten with
    the dense output at tf (kernels.jl:53-57)."""
    p = lorenz_sweep(4)
    r = oracle.solve("lorenz", "tsit5", U0_LORENZ, p, [0, 10], dt=0.1, length=101)
    assert r["naccept"].tolist() == [100] * 4
    assert r["ts"][0, -1] == np.float32(10.0) and r["ts"][0, -2] == np.float32(9.900002)
    e = oracle.solve("lorenz", "tsit5", U0_LORENZ, p, [0, 10], dt=0.1, save_everystep=False)
    assert e["ts"][0, 1] == np.float32(10.000002)      # Q2: raw integ.t, not tf
    assert not np.array_equal(e["us"][:, 1], r["us"][:, -1])


def test_quirk_q3_adaptive_save_everystep_writes_row1_only():
    r = oracle.solve("lorenz", "tsit5", U0_LORENZ, P0_LORENZ.astype(np.float32), [0, 1], dt=0.1,
                     adaptive=True, save_everystep=True, length=11)
    assert np.all(r["ts"][0] == 0) and np.all(r["us"][0, 0] == U0_LORENZ) and np.all(r["us"][0, 1:] == 0)


def test_dt_less_than_min_is_reported_not_trapped():
    # an initial step below 1e-14 makes the reference `error("dt<dtmin")`
    r = oracle.solve("lorenz", "tsit5", U0_LORENZ, P0_LORENZ, [0, 1], dt=1e-15, adaptive=True,
                     save_everystep=False, dtype=np.float64)
    assert oracle.RETCODES[int(r["retcode"][0])] == "DtLessThanMin" and r["ts"][0, 1] == 0.0


# ---- reference test/gpu_kernel_de/stiff_ode/gpu_ode_regression.jl:31-131 ----
@pytest.mark.parametrize("alg", ["rosenbrock23", "rodas4", "rodas5p"])
def test_stiff_decay_like_reference(alg):
    sol = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, length=1001)
    asol = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, adaptive=True, save_everystep=False)
    tr = 10 * np.exp(-10.0)
    assert abs(sol["us"][0, -1, 0] - tr) < 5e-3 and abs(asol["us"][0, 1, 0] - tr) < 6e-3
    sv = np.array([2.0, 4.0], np.float32)
    a2 = oracle.solve("decay", alg, [10.0], [1.0], [0, 10], dt=0.01, adaptive=True, abstol=1e-7, reltol=1e-7, saveat=sv)
    assert np.abs(a2["us"][0, :, 0] - 10 * np.exp(-sv)).max() < 7e-3


@pytest.mark.parametrize("alg,tol", [("rosenbrock23", 2e-4), ("rodas4", 1e-5), ("rodas5p", 1e-5)])
def test_robertson_vs_radau(alg, tol):
    def rob(t, y, k1, k2, k3):
        return [-k1 * y[0] + k3 * y[1] * y[2], k1 * y[0] - k2 * y[1] ** 2 - k3 * y[1] * y[2], k2 * y[1] ** 2]
    sv = [1.0, 10.0, 1e3, 1e5]
    tr = solve_ivp(rob, [0, 1e5], [1, 0, 0], args=tuple(K0_ROBER), rtol=1e-11, atol=1e-14, method="Radau", t_eval=sv).y.T
    r = oracle.solve("rober", alg, [1, 0, 0], K0_ROBER, [0, 1e5], dt=1e-4, adaptive=True, abstol=1e-8,
                     reltol=1e-4, saveat=sv, dtype=np.float64)
    assert np.abs(r["us"][0] - tr).max() < tol
    assert abs(r["us"][0, -1].sum() - 1) < 1e-12      # invariant y1+y2+y3 = 1


@pytest.mark.parametrize("alg", ["rosenbrock23", "rodas4", "rodas5p"])
def test_general_lu_path_linear15(alg):
    u0 = np.linspace(0.1, 1, 15)
    r = oracle.solve("linear15", alg, u0, None, [0, 1], dt=0.01, adaptive=True, abstol=1e-9, reltol=1e-9,
                     save_everystep=False, dtype=np.float64)
    assert np.abs(r["us"][0, 1] - u0 * np.exp(1.01)).max() < 2e-6


# ---- reference test/gpu_kernel_de/gpu_sde_regression.jl:42, gpu_sde_convergence.jl:32,46 ----
@pytest.mark.parametrize("alg,tol", [("em", 6e-2), ("siea", 6e-2)])
def test_sde_mean_like_reference(alg, tol):
    n = 4000
    u0 = np.full((n, 1), 0.5, np.float32)
    sv = np.linspace(0, 1, 11).astype(np.float32)
    r = oracle.solve("scalar_sde", alg, u0, [1.0, 1.0], [0, 1], dt=1 / 128, saveat=sv, seed=42)
    mean = r["us"][:, :, 0].mean(0)
    assert np.abs(mean[:-1] - 0.5 * np.exp(sv[:-1])).max() < tol * 1.5


def test_sde_weak_orders():
    """EM weak order ~1, SIEA ~2 on dX = a X dt + b X dW (E[X_1] = x0 e^a)"""
    n = 200000
    u0 = np.full((n, 1), 0.5)
    out = {}
    for alg in ("em", "siea"):
        errs = []
        for dt in (1 / 8, 1 / 16, 1 / 32):
            r = oracle.solve("scalar_sde", alg, u0, [1.0, 0.02], [0, 1], dt=dt, save_everystep=False, seed=3, dtype=np.float64)
            errs.append(abs(r["us"][:, 1, 0].mean() - 0.5 * np.e))
        out[alg] = np.log2(errs[0] / errs[2]) / 2
    assert 0.85 < out["em"] < 1.15, out      # reference: 1.0 +- 0.1  (gpu_sde_convergence.jl:32)
    assert 1.7 < out["siea"] < 2.5, out      # reference: 2.1 +- 0.4  (gpu_sde_convergence.jl:46)


def test_philox_known_answer():
    """Random123 kat_vectors: philox4x32-10, counter/key all zero and all ones"""
    import ctypes
    out = (ctypes.c_uint32 * 4)()
    oracle.lib().degk_oracle_philox(0, 0, 0, 0, 0, 0, out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    oracle.lib().degk_oracle_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_sde_seeds_are_independent_streams():
    """with key = seed ^ trajectory the ensembles of seeds s and s ^ d were permutations of each other.
    Seed and trajectory index now sit in different Philox words."""
    n = 64
    u0 = np.full((n, 1), 0.5, np.float32)
    kw = dict(dt=1 / 32, save_everystep=False)
    a = oracle.solve("scalar_sde", "em", u0, [1.0, 0.5], [0, 1], seed=0, **kw)["us"][:, 1, 0]
    b = oracle.solve("scalar_sde", "em", u0, [1.0, 0.5], [0, 1], seed=1, **kw)["us"][:, 1, 0]
    a2 = oracle.solve("scalar_sde", "em", u0, [1.0, 0.5], [0, 1], seed=0, **kw)["us"][:, 1, 0]
    assert np.array_equal(a, a2)
    assert len(np.intersect1d(a, b)) == 0
    # trajectory i of seed 0 is not trajectory i ^ 1 of seed 1 any more
    assert not np.array_equal(a[np.arange(n) ^ 1], b)


def test_normals_form_one_sequence_per_trajectory():
    """The stream definition (oracle normals_for_step, kernel normals_of_block): normal q of a trajectory is element
    q & 3 of the Box-Muller pairs of Philox block q >> 2 with key = seed and counter (block, trajectory), whatever the
    number of noise terms per step -- so m = 1, 2, 3, 4 read the same sequence, no word of a block is skipped, and the
    normals are Box-Muller of the raw Philox words."""
    import ctypes
    lib = oracle.lib()
    seed, traj = 0x1234567890ABCDEF, 4097
    def seq(m, steps):
        out = []
        for j in range(steps):
            z = (ctypes.c_float * m)()
            lib.degk_oracle_normals_f32(ctypes.c_uint64(seed), ctypes.c_uint64(traj), ctypes.c_uint32(j), m, z)
            out += list(z)
        return np.array(out, np.float32)
    ref = seq(4, 12)                                   # 48 normals = blocks 0..11
    for m, steps in ((1, 48), (2, 24), (3, 16)):
        assert np.array_equal(seq(m, steps), ref), m
    # block 0 from the raw words: u = ((x >> 8) + 1) 2^-24, r = sqrt(-2 ln u1), (r cos 2 pi u2, r sin 2 pi u2)
    w = (ctypes.c_uint32 * 4)()
    lib.degk_oracle_philox(0, 0, traj & 0xFFFFFFFF, traj >> 32, seed & 0xFFFFFFFF, seed >> 32, w)
    u = [np.float32(((x >> 8) + 1) * 2.0 ** -24) for x in w]
    exp = []
    for a, b in ((u[0], u[1]), (u[2], u[3])):
        r = np.sqrt(np.float32(-2) * np.log(a))
        th = np.float32(6.283185307179586) * b
        exp += [r * np.cos(th), r * np.sin(th)]
    assert np.allclose(ref[:4], np.array(exp, np.float32), rtol=2e-6, atol=1e-7)
