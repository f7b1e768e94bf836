// degk_dae_init.cuh -- consistent initialisation of a mass-matrix DAE before the first step.
//
// Reference: kernels.jl:19-25, 93-99 call `gpu_initialization_solve(prob, SimpleTrustRegion(), abstol, reltol)`
// (nlsolve/initialization.jl:1-54) when the function carries initialization data: ModelingToolkit's generated
// nonlinear "initializeprob" is solved with SimpleNonlinearSolve's trust-region Newton and its solution mapped back
// to u0 (and p).  Neither the generated problem nor SimpleNonlinearSolve is vendored, so the algorithm is restated from
// its published description (parity unpinned; the reference's own test of this path is `@test_broken`):
//
//   unknowns   the algebraic states: i with row i and column i of the constant mass matrix all zero
//   residual   F_i = f_i(u, p, t0) on those rows, the differential states keep their given values
//   iteration  Newton step d = -J_aa \ F (analytic / dual / finite-difference Jacobian as the stepper uses),
//              clipped to the trust radius; rho = actual / predicted reduction of |F|^2; accepted when
//              rho > 1e-4; radius * 1/4 when rho < 1/4, * 2 (up to the maximum) when rho > 3/4 on a clipped step
//   start      max radius = max(|F0|, max(u0) - min(u0)), initial radius = max radius / 11
//   stop       |F|_inf <= abstol (success), or a step below reltol * |u|_inf + abstol (success), 1000 iterations,
//              32 consecutive shrinks or a singular J_aa (failure)
//
// On failure the trajectory is not integrated: row 1 (and row 2 for endpoints-only output) get prob.u0 / t0
// (kernels.jl:63-70, 143-150) and the return code is DEGK_RC_INIT_FAILURE.
#pragma once
#include "degk_rosenbrock.cuh"

namespace degk {

enum { RC_INIT_FAILURE = 7 };

template <class T, class Model>
DEGK_DEV bool dae_initialize(T (&u)[Model::N], const T* p, T t0, T abstol, T reltol) {
    constexpr int N = Model::N;
    if constexpr (!has_mass_of<Model>::value) {
        (void)u; (void)p; (void)t0; (void)abstol; (void)reltol;
        return true;
    } else {
        T Mm[N][N];
        DEGK_UNROLL for (int i = 0; i < N; ++i) DEGK_UNROLL for (int j = 0; j < N; ++j) Mm[i][j] = (T)0;
        Model::template mass<T>(Mm);
        bool alg[N];
        bool any = false;
        DEGK_UNROLL for (int i = 0; i < N; ++i) {
            bool z = true;
            DEGK_UNROLL for (int j = 0; j < N; ++j) z = z && Mm[i][j] == (T)0 && Mm[j][i] == (T)0;
            alg[i] = z; any = any || z;
        }
        if (!any) return true;
        T f[N];
        Model::template f<T>(f, u, p, t0);
        auto norm2 = [&](const T (&v)[N]) { T s = (T)0; DEGK_UNROLL for (int i = 0; i < N; ++i) if (alg[i]) s = s + v[i] * v[i]; return s; };
        T fn2 = norm2(f);
        T umax = u[0], umin = u[0];
        DEGK_UNROLL for (int i = 1; i < N; ++i) { umax = fmax_(umax, u[i]); umin = fmin_(umin, u[i]); }
        const T rmax = fmax_(sqrt_(fn2), umax - umin);
        T radius = rmax / (T)11;
        int shrinks = 0;
        for (int it = 0; it < 1000; ++it) {
            T finf = (T)0;
            DEGK_UNROLL for (int i = 0; i < N; ++i) if (alg[i]) finf = fmax_(finf, abs_(f[i]));
            if (!(finf > abstol)) return finf == finf;          // converged (a NaN residual fails)
            T J[N][N], A[N][N], b[N], d[N];
            DEGK_UNROLL for (int i = 0; i < N; ++i) DEGK_UNROLL for (int j = 0; j < N; ++j) J[i][j] = (T)0;
            eval_jac<T, Model>(J, u, p, t0);
            DEGK_UNROLL for (int i = 0; i < N; ++i) {
                DEGK_UNROLL for (int j = 0; j < N; ++j) A[i][j] = (alg[i] && alg[j]) ? J[i][j] : ((i == j) ? (T)1 : (T)0);
                b[i] = alg[i] ? -f[i] : (T)0;
            }
            LinSolve<T, N> ls;
            if (!ls.factor(A)) return false;
            ls.solve(b, d);
            T dn = (T)0;
            DEGK_UNROLL for (int i = 0; i < N; ++i) dn = dn + d[i] * d[i];
            dn = sqrt_(dn);
            if (!(dn == dn) || !finite_(dn)) return false;       // singular closed-form solve (n <= 3) shows up here
            const bool clipped = dn > radius;
            if (clipped) { const T s = radius / dn; DEGK_UNROLL for (int i = 0; i < N; ++i) d[i] = d[i] * s; dn = radius; }
            // predicted residual F + J_aa d
            T fp[N];
            DEGK_UNROLL for (int i = 0; i < N; ++i) {
                T s = f[i];
                DEGK_UNROLL for (int j = 0; j < N; ++j) if (alg[i] && alg[j]) s = s + J[i][j] * d[j];
                fp[i] = s;
            }
            T un[N], fnw[N];
            DEGK_UNROLL for (int i = 0; i < N; ++i) un[i] = u[i] + d[i];
            Model::template f<T>(fnw, un, p, t0);
            const T fn2_new = norm2(fnw), pred = fn2 - norm2(fp);
            const T rho = (fn2 - fn2_new) / pred;
            if (rho > (T)1e-4 && fn2_new == fn2_new) {
                DEGK_UNROLL for (int i = 0; i < N; ++i) { u[i] = un[i]; f[i] = fnw[i]; }
                fn2 = fn2_new;
                T uinf = (T)0;
                DEGK_UNROLL for (int i = 0; i < N; ++i) uinf = fmax_(uinf, abs_(u[i]));
                if (dn <= reltol * uinf + abstol) {
                    T fi = (T)0;
                    DEGK_UNROLL for (int i = 0; i < N; ++i) if (alg[i]) fi = fmax_(fi, abs_(f[i]));
                    if (fi <= abstol) return true;
                }
            }
            if (!(rho >= (T)0.25)) { radius = radius * (T)0.25; if (++shrinks >= 32) return false; }
            else { shrinks = 0; if (rho > (T)0.75 && clipped) radius = fmin_((T)2 * radius, rmax); }
        }
        return false;
    }
}

}  // namespace degk
