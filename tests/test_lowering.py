"""RHS lowering front-end: host functions traced to CUDA C++ bodies.

CPU: the emitted bodies are compiled as plain C++ templates with g++ and compared with the host function
(values, Jacobian against central differences, time gradient), and compiled with NVRTC for every scalar type
the kernels instantiate (float / double, packed pairs in the fast build, forward-mode duals without a Jacobian).
GPU: traced models against the hand-written bodies / built-in structs and, through those, the oracle.
"""
import ctypes
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

from cases import U0_LORENZ, lorenz_sweep, rober_sweep  # noqa: E402

f32, f64 = np.float32, np.float64


def lorenz_py(u, p, t):
    return [p[0] * (u[1] - u[0]), u[0] * (p[1] - u[2]) - u[1], u[0] * u[1] - p[2] * u[2]]


def rober_py(u, p, t):
    return [-p[0] * u[0] + p[2] * u[1] * u[2], p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2], p[1] * u[1] ** 2]


def wild_py(u, p, t):
    # every construct the printer knows: literals, rationals, integer / half / real powers, the function set
    return [u[1] + np.sin(t) * np.exp(-u[0] ** 2),
            -u[0] + p[0] * np.cos(2.5 * t) - 0.1 * u[1] ** 3 + np.sqrt(1 + u[0] ** 2) / 3 + np.tanh(u[1])
            + (2 + u[0] ** 2) ** -2.5 + (u[0] + u[1]) ** 2 + np.log(3 + u[1] ** 2) + np.tan(0.1 * u[0])
            + np.sinh(0.2 * u[1]) + np.cosh(0.3 * u[0]) + (1 + u[1] ** 2) ** 1.5 + np.pi * p[1]]


HARNESS = r"""
#include <cmath>
using std::sin; using std::cos; using std::exp; using std::log; using std::sqrt;
template <class T, int N> struct M {
    static void rhs(T* du, const T* u, const T* p, T t) { %(rhs)s }
    static void jac(T (&J)[N][N], const T* u, const T* p, T t) { %(jac)s }
    static void tgrad(T* dT, const T* u, const T* p, T t) { %(tgrad)s }
};
template <class T> static void run(T* du, T* Jo, T* dT, const T* u, const T* p, T t) {
    constexpr int N = %(n)d;
    M<T, N>::rhs(du, u, p, t);
    T J[N][N];
    for (int i = 0; i < N; ++i) { dT[i] = 0; for (int j = 0; j < N; ++j) J[i][j] = 0; }
    M<T, N>::jac(J, u, p, t);
    M<T, N>::tgrad(dT, u, p, t);
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) Jo[i * N + j] = J[i][j];
}
extern "C" void run_f32(float* du, float* J, float* dT, const float* u, const float* p, float t) { run<float>(du, J, dT, u, p, t); }
extern "C" void run_f64(double* du, double* J, double* dT, const double* u, const double* p, double t) { run<double>(du, J, dT, u, p, t); }
"""


def build_host(tmp_path, func, n, name):
    src = tmp_path / f"{name}.cpp"
    src.write_text(HARNESS % dict(rhs=func.rhs, jac=func.jac or "", tgrad=func.tgrad or "", n=n))
    so = tmp_path / f"{name}.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-ffp-contract=off", str(src), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("name,py,n,npar", [("lorenz", lorenz_py, 3, 3), ("rober", rober_py, 3, 3), ("wild", wild_py, 2, 2)])
def test_lowered_bodies_evaluate_like_the_host_function(tmp_path, name, py, n, npar):
    import diffeqgpu_b200 as dg
    func = dg.ODEFunction.from_python(py, n, npar, jac=True)
    lib = build_host(tmp_path, func, n, name)
    rng = np.random.default_rng(3)
    for dtype, fn, tol in ((f64, lib.run_f64, 1e-13), (f32, lib.run_f32, 2e-5)):
        ct = ctypes.c_double if dtype == f64 else ctypes.c_float
        for _ in range(5):
            u = rng.uniform(0.2, 1.5, n).astype(dtype); p = rng.uniform(0.5, 2.0, npar).astype(dtype); t = dtype(rng.uniform(0, 3))
            du = np.zeros(n, dtype); J = np.zeros((n, n), dtype); dT = np.zeros(n, dtype)
            ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)      # noqa: E731
            fn(ptr(du), ptr(J), ptr(dT), ptr(u), ptr(p), ct(float(t)))
            ud, pd, td = u.astype(f64), p.astype(f64), float(t)
            want = np.asarray(py(ud, pd, td), f64)
            assert np.allclose(du, want, rtol=tol, atol=tol), (name, dtype)
            assert np.allclose(func.python(ud, pd, td), want)
            h = 1e-6
            Jfd = np.stack([(np.asarray(py(ud + h * e, pd, td)) - np.asarray(py(ud - h * e, pd, td))) / (2 * h) for e in np.eye(n)], 1)
            Tfd = (np.asarray(py(ud, pd, td + h)) - np.asarray(py(ud, pd, td - h))) / (2 * h)
            assert np.allclose(J, Jfd, rtol=max(tol, 1e-6), atol=max(tol, 2e-6) * np.abs(Jfd).max()), (name, dtype)
            assert np.allclose(dT, Tfd, rtol=max(tol, 1e-6), atol=max(tol, 2e-6) * max(1.0, np.abs(Tfd).max())), (name, dtype)


def test_lowering_keeps_two_term_expression_trees():
    """Lorenz has only two-operand sums and products, so the traced body rounds like the hand-written one"""
    import diffeqgpu_b200 as dg
    func = dg.ODEFunction.from_python(lorenz_py, 3, 3, jac=True)
    assert "du[0] = p[0]*(-u[0] + u[1]);" in func.rhs and "du[1] = u[0]*(p[1] - u[2]) - u[1];" in func.rhs
    assert func.jac.count("J[") == 8 and "J[0][2]" not in func.jac and func.tgrad is None          # autonomous: no time-gradient body at all (TGRAD_ZERO)
    plain = dg.ODEFunction.from_python(lorenz_py, 3, 3)
    assert plain.jac is None and plain.tgrad is None and plain.n_state == 3 and plain.n_param == 3
    # float-literal exponents that are integers or halves stay products / square roots (no exp-log detour, which
    # would fail for negative bases)
    pw = dg.ODEFunction.from_python(lambda u, p, t: [u[0] ** 2.0 + u[0] ** 0.5 + u[0] ** -1.0], 1, 0)
    assert "u[0] * u[0]" in pw.rhs and "sqrt(u[0])" in pw.rhs and "(T)1 / (u[0])" in pw.rhs and "log" not in pw.rhs
    # numpy arrays of traced values: matrix-vector products, broadcasting, ufuncs
    A = dg.ODEFunction.from_python(lambda u, p, t: np.array([[0.0, p[0]], [-1.0, 0.0]]) @ np.array(u) + np.sin(np.array(u)) * p[0], 2, 1)
    assert "du[0] = p[0]*u[1] + p[0]*sin(u[0]);" in A.rhs.replace("p[0]*sin(u[0]) + p[0]*u[1]", "p[0]*u[1] + p[0]*sin(u[0])")
    # shared subexpressions are hoisted once
    rob = dg.ODEFunction.from_python(rober_py, 3, 3)
    assert rob.rhs.count("const T x_") == 2 and rob.rhs.count("p[2]*u[1]*u[2]") == 1
    # constant mass matrix -> Mm body, zeros skipped
    dae = dg.ODEFunction.from_python(lambda u, p, t: [-p[0] * u[0], u[0] + u[1] - 1], 2, 1, mass_matrix=[[1, 0], [0, 0]])
    assert dae.mass_matrix == "    Mm[0][0] = (T)1.0;\n"
    with pytest.raises(ValueError):
        dg.ODEFunction.from_python(lorenz_py, 3, 3, mass_matrix=np.eye(2))


def test_odeproblem_takes_a_host_function():
    """`ODEProblem{false}(f, u0, tspan, p)` with a plain function, as in the reference's tests"""
    import diffeqgpu_b200 as dg
    prob = dg.ODEProblem(lorenz_py, np.array([1, 0, 0], f32), (0.0, 10.0), np.array([10, 28, 8 / 3], f32))
    assert isinstance(prob.f, dg.ODEFunction) and prob.f.n_state == 3 and prob.f.n_param == 3 and "du[2]" in prob.f.rhs
    assert prob.dtype == f32 and np.allclose(prob.f.python(prob.u0, prob.p, 0.0), [-10, 28, 0])
    again = dg.remake(prob, p=np.array([1, 2, 3], f32))
    assert again.f is prob.f                      # remake keeps the lowered function (one program for the ensemble)


def test_lowering_refuses_what_cannot_be_traced():
    import diffeqgpu_b200 as dg
    from diffeqgpu_b200.lowering import LoweringError
    with pytest.raises(LoweringError, match="control flow"):
        dg.ODEFunction.from_python(lambda u, p, t: [u[0] if u[0] > 0 else -u[0]], 1, 0)
    with pytest.raises(LoweringError, match="abs"):
        dg.ODEFunction.from_python(lambda u, p, t: [abs(u[0])], 1, 0)
    with pytest.raises(LoweringError, match="shape"):
        dg.ODEFunction.from_python(lambda u, p, t: [u[0], u[0]], 1, 0)
    with pytest.raises(LoweringError, match="comparison"):
        dg.DiscreteCallback.from_python(lambda u, t, p: u[0], lambda u, p, t: [u[0]], 1, 0)
    with pytest.raises(LoweringError, match="state values"):
        dg.DiscreteCallback.from_python(lambda u, t, p: u[0] > 1, lambda u, p, t: [u[0], u[0]], 1, 0)


def test_lowered_bodies_compile_with_nvrtc_for_every_scalar_type():
    import diffeqgpu_b200 as dg
    from diffeqgpu_b200 import _lib
    full = dg.ODEFunction.from_python(wild_py, 2, 2, jac=True)
    nojac = dg.ODEFunction.from_python(wild_py, 2, 2)
    for func, alg, dt, fp in ((full, 0, _lib.F64, _lib.FP_STRICT),     # GPUTsit5 Float64
                              (full, 5, _lib.F32, _lib.FP_FAST),       # GPURodas5P, packed pairs with analytic jac / tgrad
                              (nojac, 3, _lib.F32, _lib.FP_STRICT)):   # GPURosenbrock23, forward-mode duals through every function
        d = _lib.make_desc(rhs_src=func.rhs, jac_src=func.jac, tgrad_src=func.tgrad, n_state=2, n_param=2, dtype=dt, alg=alg, fp_mode=fp)
        st, nb, log = _lib.jit_compile_check(d)
        assert st == 0 and nb > 0, log
    # callbacks: discrete (comparison + new u / new p / terminate) and continuous (root function, None affect_neg)
    cbs = dg.CallbackSet(
        dg.DiscreteCallback.from_python(lambda u, t, p: (t == 2.4) | ((u[0] > 3) & ~(u[1] < 0)), lambda u, p, t: [u[0] + 10, u[1]], 2, 2),
        dg.DiscreteCallback.from_python(lambda u, t, p: t >= 4, lambda u, p, t: ([u[1], u[0]], [2 * p[0], p[1]]), 2, 2),
        dg.DiscreteCallback.from_python(lambda u, t, p: u[0] < -50, lambda u, p, t: "terminate", 2, 2),
        dg.ContinuousCallback.from_python(lambda u, t, p: u[0] - 0.5 * p[0], lambda u, p, t: [u[0], -0.9 * u[1]], 2, 2, affect_neg=None))
    d = _lib.make_desc(rhs_src=full.rhs, n_state=2, n_param=2, dtype=_lib.F32, alg=0, callbacks=cbs.key(), ccallbacks=cbs.ckey())
    st, nb, log = _lib.jit_compile_check(d)
    assert st == 0 and nb > 0, log
    # SDE: diagonal (GPUEM, GPUSIEA) and general noise (GPUEM)
    sd = dg.SDEFunction.from_python(wild_py, lambda u, p, t: [p[1] * u[0], 0.1 + np.sin(t)], 2, 2)
    sg = dg.SDEFunction.from_python(wild_py, lambda u, p, t: [[p[1] * u[0], 0, 0.5], [0, np.exp(-u[1] ** 2), 1]], 2, 2, noise="general", n_noise=3)
    for sf, alg, kind, m in ((sd, 7, _lib.NOISE_DIAGONAL, 2), (sg, 6, _lib.NOISE_GENERAL, 3)):
        d = _lib.make_desc(rhs_src=sf.f.rhs, noise_src=sf.g, n_state=2, n_param=2, n_noise=m, noise_kind=kind, dtype=_lib.F32, alg=alg)
        st, nb, log = _lib.jit_compile_check(d)
        assert st == 0 and nb > 0, log


def test_lowered_sde_bodies():
    import diffeqgpu_b200 as dg
    s = dg.SDEFunction.from_python(lambda u, p, t: [p[0] * u[i] for i in range(3)], lambda u, p, t: [p[1] * u[i] for i in range(3)], 3, 2)
    assert s.noise == "diagonal" and s.g.count("g[") == 3 and s.f.rhs.count("du[") == 3
    g = dg.SDEFunction.from_python(lambda u, p, t: [p[0] * u[0], p[0] * u[1]],
                                   lambda u, p, t: [[p[1] * u[0], 0, 0.5, 0], [0, p[1] * u[1], 0, 1]], 2, 2, noise="general", n_noise=4)
    assert g.n_noise == 4 and "G[0][2] = (T)0.5;" in g.g and "G[0][1]" not in g.g
    with pytest.raises(ValueError):
        dg.SDEFunction.from_python(lorenz_py, lorenz_py, 3, 3, noise="scalar")


# ---------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle
    return oracle


def _solve(dg, func, alg, u0, p, tspan, *, adaptive, fp_mode="strict", callback=None, **kw):
    import torch
    u0 = np.asarray(u0, f32); p = np.asarray(p, f32)
    prob = dg.ODEProblem(func, u0[0] if u0.ndim == 2 else u0, tuple(tspan), p[0] if p.ndim == 2 else p)
    n = max(u0.shape[0] if u0.ndim == 2 else 1, p.shape[0] if p.ndim == 2 else 1)
    probs = dg.ProblemBatch.from_arrays(prob, u0=u0 if u0.ndim == 2 else None, p=p if p.ndim == 2 else None, n_traj=n, device="cuda:0")
    kw = {k: (f32(v) if k in ("dt", "abstol", "reltol") else v) for k, v in kw.items()}
    solve = dg.vectorized_asolve if adaptive else dg.vectorized_solve
    ts, us, st = solve(probs, prob, alg, fp_mode=fp_mode, stats=True, callback=callback, **kw)
    torch.cuda.synchronize()
    return dict(ts=ts.cpu().numpy(), us=us.cpu().numpy(), naccept=st["naccept"].cpu().numpy(), retcode=st["retcode"].cpu().numpy())


@pytest.mark.gpu
def test_gpu_traced_lorenz_is_bit_identical_to_the_oracle(oracle):
    import diffeqgpu_b200 as dg
    func = dg.ODEFunction.from_python(lorenz_py, 3, 3, jac=True)
    p = lorenz_sweep(300, seed=31)
    sv = np.arange(0, 11, dtype=f32)
    kw = dict(dt=0.1, abstol=1e-6, reltol=1e-6, saveat=sv)
    g = _solve(dg, func, dg.GPUTsit5(), U0_LORENZ, p, [0, 10], adaptive=True, **kw)
    r = oracle.solve("lorenz", "tsit5", U0_LORENZ, p, [0, 10], adaptive=True, **kw)
    assert np.array_equal(g["us"], r["us"]) and np.array_equal(g["naccept"], r["naccept"])
    # the symbolic Jacobian has the hand-written one's entries: Rodas5P matches the oracle too
    g = _solve(dg, func, dg.GPURodas5P(), U0_LORENZ, p, [0, 2], adaptive=True, dt=0.01, abstol=1e-6, reltol=1e-6, saveat=sv[:3])
    r = oracle.solve("lorenz", "rodas5p", U0_LORENZ, p, [0, 2], adaptive=True, dt=0.01, abstol=1e-6, reltol=1e-6, saveat=sv[:3])
    assert np.array_equal(g["us"], r["us"]) and np.array_equal(g["naccept"], r["naccept"])
    # fast build (packed pairs)
    gf = _solve(dg, func, dg.GPUTsit5(), U0_LORENZ, p, [0, 1], adaptive=True, fp_mode="fast", dt=0.1, abstol=1e-6, reltol=1e-6, saveat=sv[:2])
    r1 = oracle.solve("lorenz", "tsit5", U0_LORENZ, p, [0, 1], adaptive=True, dt=0.1, abstol=1e-6, reltol=1e-6, saveat=sv[:2])
    assert np.allclose(gf["us"], r1["us"], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_gpu_traced_robertson_and_callbacks(oracle):
    import diffeqgpu_b200 as dg
    k = rober_sweep(128)
    sv = np.array([1.0, 10.0, 1e3, 1e5], f32)
    kw = dict(dt=1e-4, abstol=1e-8, reltol=1e-4, saveat=sv)
    r = oracle.solve("rober", "rodas5p", U0_LORENZ, k, [0, 1e5], adaptive=True, **kw)
    for jac in (True, False):                 # symbolic Jacobian / forward-mode duals of the traced body
        func = dg.ODEFunction.from_python(rober_py, 3, 3, jac=jac)
        g = _solve(dg, func, dg.GPURodas5P(), U0_LORENZ, k, [0, 1e5], adaptive=True, **kw)
        assert (g["retcode"] == 1).all() and np.allclose(g["us"], r["us"], rtol=2e-3, atol=1e-7)   # re-associated products: tolerance
        assert np.abs(g["us"].sum(axis=2) - 1).max() < 1e-4
    # the kick of gpu_ode_discrete_callbacks.jl:26-60, condition / affect! given as host functions
    decay = dg.ODEFunction.from_python(lambda u, p, t: [-p[0] * u[0]], 1, 1)
    cb = dg.DiscreteCallback.from_python(lambda u, t, p: t == 2.4, lambda u, p, t: [u[0] + 10], 1, 1)
    g = _solve(dg, decay, dg.GPUTsit5(), [10.0], [1.0], [0, 10], adaptive=False, dt=0.5, tstops=[2.4], callback=cb)
    ro = oracle.solve("decay", "tsit5", [10.0], [1.0], [0, 10], dt=0.5, length=g["us"].shape[1], tstops=[2.4],
                      callbacks=[(("t_eq", 0, 2.4), ("u_add", 0, 10.0))])
    assert np.array_equal(g["ts"], ro["ts"]) and np.array_equal(g["us"], ro["us"])
    # bouncing ball, ContinuousCallback from host functions
    ball = dg.ODEFunction.from_python(lambda u, p, t: [u[1], -p[0]], 2, 1)
    cc = dg.ContinuousCallback.from_python(lambda u, t, p: u[0], lambda u, p, t: [u[0], -u[1]], 2, 1)
    g = _solve(dg, ball, dg.GPUTsit5(), [45.0, 0.0], [10.0], [0, 10], adaptive=True, dt=0.1, abstol=1e-6, reltol=1e-6,
               saveat=np.array([0.0, 9.1], f32), callback=cc)
    ro = oracle.solve("ball", "tsit5", [45.0, 0.0], [10.0], [0, 10], dt=0.1, adaptive=True, abstol=1e-6, reltol=1e-6,
                      saveat=np.array([0.0, 9.1], f32),
                      continuous_callbacks=[dict(condition=("u_minus", 0, 0.0), affect=("u_scale", 1, -1.0))])
    assert np.array_equal(g["us"], ro["us"])


@pytest.mark.gpu
def test_gpu_traced_sde_matches_builtin_pathwise():
    import torch
    import diffeqgpu_b200 as dg
    traced = dg.SDEFunction.from_python(lambda u, p, t: [p[0] * u[i] for i in range(3)], lambda u, p, t: [p[1] * u[i] for i in range(3)], 3, 2)
    out = []
    for func in (traced, dg.models.gbm):
        prob = dg.SDEProblem(func, np.full(3, 0.1, f32), (0.0, 1.0), np.array([1.5, 0.2], f32), seed=1234)
        probs = dg.ProblemBatch.from_arrays(prob, n_traj=512, device="cuda:0", seed=1234)
        ts, us, st = dg.vectorized_solve(probs, prob, dg.GPUEM(), dt=f32(1 / 64), save_everystep=False, stats=True)
        torch.cuda.synchronize()
        out.append(us.cpu().numpy())
    assert np.array_equal(out[0], out[1])


# ---------------------------------------------------------------------------------------- fuzz (CPU)
def _random_tree(rng, depth):
    """a random expression over u[0..2], p[0..1], t as (source string in numpy terms)"""
    if depth == 0 or rng.random() < 0.2:
        k = rng.integers(0, 8)
        return [f"u[{k}]" for k in range(3)][k] if k < 3 else ["p[0]", "p[1]", "t", f"{rng.uniform(-2, 2):.6f}", str(int(rng.integers(1, 4)))][k - 3]
    op = rng.integers(0, 11)
    a = _random_tree(rng, depth - 1)
    if op <= 3:
        b = _random_tree(rng, depth - 1)
        return f"({a} {'+-*'[min(op, 2)]} {b})" if op < 3 else f"({a} / (1.5 + ({b}) ** 2))"
    if op == 4:
        return f"({a}) ** {int(rng.integers(2, 5))}"
    if op == 5:
        return f"np.sin({a})"
    if op == 6:
        return f"np.cos({a})"
    if op == 7:
        return f"np.exp(-({a}) ** 2)"
    if op == 8:
        return f"np.log(1.25 + ({a}) ** 2)"
    if op == 9:
        return f"np.sqrt(0.5 + ({a}) ** 2)"
    return f"np.tanh({a})"


def test_lowering_fuzz_against_the_host_functions(tmp_path):
    """25 random 3-state models: the lowered rhs / symbolic jac / tgrad, compiled as C++, against the host function"""
    import diffeqgpu_b200 as dg
    rng = np.random.default_rng(2024)
    n, npar, nfun = 3, 2, 25
    pys, funcs = [], []
    for _ in range(nfun):
        src = "lambda u, p, t: [" + ", ".join(_random_tree(rng, 4) for _ in range(n)) + "]"
        py = eval(src, {"np": np})
        pys.append(py)
        funcs.append(dg.ODEFunction.from_python(py, n, npar, jac=True))
    code = ["#include <cmath>", "using std::sin; using std::cos; using std::exp; using std::log; using std::sqrt;"]
    for k, f in enumerate(funcs):
        code.append(f"""extern "C" void f{k}(double* du, double* Jo, double* dT, const double* u, const double* p, double t) {{
    typedef double T; T J[{n}][{n}] = {{}}; for (int i = 0; i < {n}; ++i) dT[i] = 0;
    {{ {f.rhs} }} {{ {f.jac} }} {{ {f.tgrad or ""} }}
    for (int i = 0; i < {n}; ++i) for (int j = 0; j < {n}; ++j) Jo[i * {n} + j] = J[i][j]; }}""")
    src = tmp_path / "fuzz.cpp"
    src.write_text("\n".join(code))
    so = tmp_path / "fuzz.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-ffp-contract=off", str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)      # noqa: E731
    for k, py in enumerate(pys):
        for _ in range(3):
            u = rng.uniform(-1.5, 1.5, n); p = rng.uniform(0.5, 2.0, npar); t = float(rng.uniform(0, 2))
            du = np.zeros(n); J = np.zeros((n, n)); dT = np.zeros(n)
            getattr(lib, f"f{k}")(ptr(du), ptr(J), ptr(dT), ptr(u), ptr(p), ctypes.c_double(t))
            want = np.asarray(py(u, p, t), f64)
            assert np.allclose(du, want, rtol=1e-11, atol=1e-11), (k, funcs[k].rhs)
            h = 1e-6
            Jfd = np.stack([(np.asarray(py(u + h * e, p, t)) - np.asarray(py(u - h * e, p, t))) / (2 * h) for e in np.eye(n)], 1)
            Tfd = (np.asarray(py(u, p, t + h)) - np.asarray(py(u, p, t - h))) / (2 * h)
            scale = max(1.0, np.abs(Jfd).max())
            assert np.allclose(J, Jfd, rtol=1e-5, atol=1e-6 * scale), (k, funcs[k].jac)
            assert np.allclose(dT, Tfd, rtol=1e-5, atol=1e-6 * max(1.0, np.abs(Tfd).max())), (k, funcs[k].tgrad)
