"""Lock-step fixed-dt kernel (csrc/device/degk_ode_lockstep.cuh; replaces reference kernels.jl:1-72 for launches with
one (t0, tf, dt), every-step saves and an explicit RK stepper): the strict build against the CPU oracle bit for bit,
against the one-thread-per-trajectory kernel, both output layouts, the overshoot row, sizes that are not multiples of
the 32 W trajectories of a warp, the JIT path, and the packed fast build within tolerance."""
import numpy as np
import pytest

from cases import U0_LORENZ, lorenz_sweep

pytestmark = pytest.mark.gpu
f32, f64 = np.float32, np.float64
ALGS = {"tsit5": "GPUTsit5", "vern7": "GPUVern7", "vern9": "GPUVern9"}


@pytest.fixture(scope="module")
def dg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import diffeqgpu_b200 as dg
    return dg


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle
    return oracle


def solve(dg, alg, p, tspan, dt, *, engine, dtype=f32, fp_mode="strict", layout="ref", func=None, u0=U0_LORENZ):
    import torch
    p = np.asarray(p, dtype=dtype)
    prob = dg.ODEProblem(func or dg.models.lorenz, np.asarray(u0, dtype=dtype), tuple(tspan), p[0])
    probs = dg.ProblemBatch.from_arrays(prob, p=p, device="cuda:0")
    ts, us, st = dg.vectorized_solve(probs, prob, getattr(dg, ALGS[alg])(), dt=dtype(dt), fp_mode=fp_mode, layout=layout,
                                     stats=True, engine=engine)
    torch.cuda.synchronize()
    return dict(ts=ts.cpu().numpy(), us=us.cpu().numpy(), naccept=st["naccept"].cpu().numpy(),
                nreject=st["nreject"].cpu().numpy(), retcode=st["retcode"].cpu().numpy(), totals=st["totals"].cpu().numpy())


def same(a, b, what):
    for k in ("ts", "us", "naccept", "nreject", "retcode"):
        assert np.array_equal(a[k], b[k], equal_nan=True), f"{what}: {k} differs"


@pytest.mark.parametrize("n", [1, 31, 33, 64, 65, 257, 1000])
def test_strict_bit_exact_against_the_oracle(dg, oracle, n):
    p = lorenz_sweep(n, seed=5 + n)
    g = solve(dg, "tsit5", p, [0, 10], 0.1, engine="lockstep")
    r = oracle.solve("lorenz", "tsit5", U0_LORENZ, p, [0, 10], dt=0.1, length=g["us"].shape[1])
    same(g, r, f"C1 lock-step n={n}")
    assert g["totals"][0] == g["naccept"].sum() and (g["naccept"] == g["naccept"][0]).all()


@pytest.mark.parametrize("alg", ["tsit5", "vern7", "vern9"])
@pytest.mark.parametrize("tspan,dt", [([0, 10], 0.3), ([0, 1], 0.25), ([0.5, 3], 0.07)])
def test_strict_steppers_and_overshoot_row(dg, oracle, alg, tspan, dt):
    """a dt that does not divide the span: the last row holds the value interpolated at tf (kernels.jl:53-57)"""
    p = lorenz_sweep(300, seed=11)
    g = solve(dg, alg, p, tspan, dt, engine="lockstep")
    r = oracle.solve("lorenz", alg, U0_LORENZ, p, tspan, dt=dt, length=g["us"].shape[1])
    same(g, r, f"{alg} {tspan} dt={dt}")
    v1 = solve(dg, alg, p, tspan, dt, engine="v1")
    same(g, v1, f"{alg} lock-step vs one thread per trajectory")


@pytest.mark.parametrize("tspan,dt", [([0, 0.3], 0.1), ([0, 0.75], 0.1), ([0, 1.65], 0.1), ([0, 3.2], 0.1), ([0, 4.05], 0.05)])
def test_short_runs_and_buffer_boundaries(dg, oracle, tspan, dt):
    """fewer rows than one flush period, exactly one period, one more, ... : first / last ragged pieces of a trajectory"""
    p = lorenz_sweep(200, seed=8)
    for fp_alg in ("tsit5", "vern7"):
        g = solve(dg, fp_alg, p, tspan, dt, engine="lockstep")
        r = oracle.solve("lorenz", fp_alg, U0_LORENZ, p, tspan, dt=dt, length=g["us"].shape[1])
        same(g, r, f"{fp_alg} {tspan}")
    a = solve(dg, "tsit5", p, tspan, dt, engine="lockstep", fp_mode="fast")
    b = solve(dg, "tsit5", p, tspan, dt, engine="lockstep", fp_mode="fast", layout="soa")
    assert np.array_equal(a["us"], b["us"].transpose(2, 0, 1)) and np.array_equal(a["ts"], b["ts"].T)


def test_four_component_state_without_parameters(dg):
    """Henon-Heiles (n = 4, no parameters): other buffer strides and sector phases"""
    from cases import henon_heiles_u0
    u0 = henon_heiles_u0(333, seed=5).astype(f32)
    import torch
    prob = dg.ODEProblem(dg.models.henon_heiles, u0[0], (0.0, 7.0), None)
    probs = dg.ProblemBatch.from_arrays(prob, u0=u0, device="cuda:0")
    out = {}
    for engine in ("lockstep", "v1"):
        for fp in ("strict", "fast"):
            ts, us, st = dg.vectorized_solve(probs, prob, dg.GPUVern7(), dt=f32(0.05), fp_mode=fp, stats=True, engine=engine)
            torch.cuda.synchronize()
            out[engine, fp] = (ts.cpu().numpy(), us.cpu().numpy(), st["naccept"].cpu().numpy())
    for k in range(3):
        assert np.array_equal(out["lockstep", "strict"][k], out["v1", "strict"][k])
    assert np.array_equal(out["lockstep", "fast"][0], out["v1", "fast"][0])
    assert np.abs(out["lockstep", "fast"][1] - out["v1", "fast"][1]).max() < 1e-3


def test_float64_strict_equals_the_per_thread_kernel(dg):
    p = lorenz_sweep(500, seed=3).astype(f64)
    for alg in ("tsit5", "vern9"):
        a = solve(dg, alg, p, [0, 5], 0.05, engine="lockstep", dtype=f64)
        b = solve(dg, alg, p, [0, 5], 0.05, engine="v1", dtype=f64)
        same(a, b, alg + " f64")


@pytest.mark.parametrize("fp", ["strict", "fast"])
def test_trajectory_major_layout(dg, fp):
    p = lorenz_sweep(777, seed=9)
    a = solve(dg, "tsit5", p, [0, 10], 0.1, engine="lockstep", fp_mode=fp)
    b = solve(dg, "tsit5", p, [0, 10], 0.1, engine="lockstep", fp_mode=fp, layout="soa")
    assert np.array_equal(a["us"], b["us"].transpose(2, 0, 1)) and np.array_equal(a["ts"], b["ts"].T)   # soa: (rows, n, N)


def test_fast_build_packed_pairs_within_tolerance(dg):
    """FMA-contracted, h-scaled stage sums, two trajectories per thread: same steps, values within rounding"""
    p = lorenz_sweep(4097, seed=21)
    s = solve(dg, "tsit5", p, [0, 1], 0.05, engine="lockstep", fp_mode="strict")
    f = solve(dg, "tsit5", p, [0, 1], 0.05, engine="lockstep", fp_mode="fast")
    assert np.array_equal(s["ts"], f["ts"]) and np.array_equal(s["naccept"], f["naccept"])
    scale = np.maximum(np.abs(s["us"]), 1.0)
    assert (np.abs(s["us"] - f["us"]) / scale).max() < 5e-4
    g = solve(dg, "tsit5", p, [0, 1], 0.05, engine="v1", fp_mode="fast")
    assert (np.abs(g["us"] - f["us"]) / scale).max() < 5e-4


@pytest.mark.parametrize("n", [1, 33, 63, 64, 65, 2049])
@pytest.mark.parametrize("layout", ["ref", "soa"])
def test_fast_build_both_widths(dg, n, layout, monkeypatch):
    """the fast Float32 build has the lock-step kernel with two trajectories per thread (packed pairs) and with one
    (degk_api.cu picks by layout and launch size; DEGK_LOCKSTEP_W1_BELOW pins it): same rows, values within rounding"""
    p = lorenz_sweep(n, seed=23)
    out = {}
    for width, below in (("one", str(1 << 40)), ("two", "0")):
        monkeypatch.setenv("DEGK_LOCKSTEP_W1_BELOW", below)
        out[width] = solve(dg, "tsit5", p, [0, 2.35], 0.05, engine="lockstep", fp_mode="fast", layout=layout)
    a, b = out["one"], out["two"]
    assert np.array_equal(a["ts"], b["ts"]) and np.array_equal(a["naccept"], b["naccept"])
    assert (np.abs(a["us"] - b["us"]) / np.maximum(np.abs(b["us"]), 1.0)).max() < 5e-4
    monkeypatch.delenv("DEGK_LOCKSTEP_W1_BELOW")
    s = solve(dg, "tsit5", p, [0, 2.35], 0.05, engine="lockstep", fp_mode="strict", layout=layout)
    assert np.array_equal(a["ts"], s["ts"])
    # (47 steps of the chaotic members amplify the rounding difference between fused and un-fused arithmetic)
    assert (np.abs(a["us"] - s["us"]) / np.maximum(np.abs(s["us"]), 1.0)).max() < 5e-3


def test_large_launch_takes_it_by_default(dg):
    """engine="auto" switches to the lock-step kernel for launches that fill the GPU: identical strict results"""
    p = lorenz_sweep(200_000, seed=2)
    a = solve(dg, "tsit5", p, [0, 2], 0.1, engine="auto")
    b = solve(dg, "tsit5", p, [0, 2], 0.1, engine="v1")
    same(a, b, "auto vs v1 at 200k")


def test_jit_model(dg):
    """a user right-hand side compiled by NVRTC gets the same kernel"""
    src = "du[0] = p[0] * (u[1] - u[0]); du[1] = u[0] * (p[1] - u[2]) - u[1]; du[2] = u[0] * u[1] - p[2] * u[2];"
    func = dg.ODEFunction(rhs=src, n_state=3, n_param=3)
    p = lorenz_sweep(130, seed=4)
    a = solve(dg, "tsit5", p, [0, 3], 0.1, engine="lockstep", func=func)
    b = solve(dg, "tsit5", p, [0, 3], 0.1, engine="lockstep")
    same(a, b, "jit vs aot strict")
    # fast build: NVRTC and nvcc may contract the model body differently; same steps, values within rounding
    a = solve(dg, "tsit5", p, [0, 1], 0.05, engine="lockstep", fp_mode="fast", func=func)
    b = solve(dg, "tsit5", p, [0, 1], 0.05, engine="lockstep", fp_mode="fast")
    assert np.array_equal(a["ts"], b["ts"]) and np.array_equal(a["naccept"], b["naccept"])
    assert (np.abs(a["us"] - b["us"]) / np.maximum(np.abs(b["us"]), 1.0)).max() < 5e-4
