"""CPU restatement (numpy, small cases) of the DAE initialisation the kernels run before the first step
(csrc/device/degk_dae_init.cuh): trust-region Newton on the algebraic rows of a mass-matrix DAE.

Reference: gpu_initialization_solve (src/ensemblegpukernel/nlsolve/initialization.jl:1-54, call sites kernels.jl:19-25,
93-99) hands ModelingToolkit's generated initialisation problem to SimpleNonlinearSolve's SimpleTrustRegion.  Neither is
vendored under /root/reference and the reference's only test of the path is `@test_broken`
(test/gpu_kernel_de/stiff_ode/gpu_ode_modelingtoolkit_dae.jl), so this is a restatement of the published algorithm:
PARITY UNPINNED.  Test infrastructure only (tests/test_mass_matrix.py)."""
import numpy as np


def dae_initialize(f, jac, mass, u0, p, t0, abstol=1e-6, reltol=1e-6, dtype=np.float32):
    """-> (u, ok).  f(u, p, t) -> du, jac(u, p, t) -> J (n x n), mass: constant n x n matrix."""
    T = np.dtype(dtype).type
    u = np.asarray(u0, dtype=dtype).copy()
    M = np.asarray(mass, dtype=dtype)
    n = u.size
    alg = np.array([not M[i, :].any() and not M[:, i].any() for i in range(n)])
    if not alg.any():
        return u, True
    F = np.asarray(f(u, p, t0), dtype=dtype)
    fn2 = T(np.sum(F[alg] * F[alg]))
    rmax = max(T(np.sqrt(fn2)), T(u.max() - u.min()))
    radius = T(rmax / T(11))
    shrinks = 0
    for _ in range(1000):
        finf = np.abs(F[alg]).max()
        if not finf > abstol:
            return u, bool(np.isfinite(finf))
        J = np.asarray(jac(u, p, t0), dtype=dtype)
        A = np.eye(n, dtype=dtype)
        for i in range(n):
            for j in range(n):
                if alg[i] and alg[j]:
                    A[i, j] = J[i, j]
        b = np.where(alg, -F, T(0)).astype(dtype)
        try:
            d = np.linalg.solve(A.astype(np.float64), b.astype(np.float64)).astype(dtype)
        except np.linalg.LinAlgError:
            return u, False
        dn = T(np.sqrt(np.sum(d * d)))
        if not np.isfinite(dn):
            return u, False
        clipped = dn > radius
        if clipped:
            d = (d * (radius / dn)).astype(dtype)
            dn = radius
        Fp = F.copy()
        for i in range(n):
            if alg[i]:
                Fp[i] = F[i] + sum(J[i, j] * d[j] for j in range(n) if alg[j])
        un = (u + d).astype(dtype)
        Fn = np.asarray(f(un, p, t0), dtype=dtype)
        fn2_new = T(np.sum(Fn[alg] * Fn[alg]))
        pred = T(fn2 - np.sum(Fp[alg] * Fp[alg]))
        with np.errstate(all="ignore"):
            rho = (fn2 - fn2_new) / pred
        if rho > 1e-4 and np.isfinite(fn2_new):
            u, F, fn2 = un, Fn, fn2_new
            if dn <= reltol * np.abs(u).max() + abstol and np.abs(F[alg]).max() <= abstol:
                return u, True
        if not rho >= 0.25:
            radius = T(radius * T(0.25))
            shrinks += 1
            if shrinks >= 32:
                return u, False
        else:
            shrinks = 0
            if rho > 0.75 and clipped:
                radius = min(T(2) * radius, rmax)
    return u, False
