// degk_kvaerno.cuh -- ESDIRK steppers GPUKvaerno3 / GPUKvaerno5 with the Newton `nlsolve`.
//
//   Kvaerno3   reference perform_step/gpu_kvaerno3_perform_step.jl:1-89 (fixed), :91-232 (adaptive)
//   Kvaerno5   reference perform_step/gpu_kvaerno5_perform_step.jl:1-143, :145-300
//   nlsolve    reference nlsolve/utils.jl:1-26, build_nlsolver nlsolve/type.jl:157-185
//              (maxiters = 30, W(u,p,t) = -M + gamma*dt*J(u,p,t), M = I)
//   tableaus   reference tableaus/kvaerno_tableaus.jl:24-52, :137-183
//
// Every Newton iteration re-evaluates J at the current stage point and solves with a fresh W
// (the reference does not freeze the Jacobian): LinSolve::factor + solve per iteration, all in
// registers.  The Jacobian comes from eval_jac (analytic body, forward-mode duals or finite
// differences, degk_dual.cuh).
//
// Method interface as in degk_rosenbrock.cuh.  attempt<false> is the fixed-dt step (the nlsolver
// is built after integ.t advanced: its time base is t + dt, gpu_kvaerno3_perform_step.jl:15-41),
// attempt<true> the adaptive attempt (time base t, k1 = f(uprev) re-evaluated every attempt,
// :131-143, error estimate = W \ sum(btilde_i z_i), :178-186).
//
// The reference defines no `_ode_interpolant` for the Kvaerno integrators (saveat and the
// interpolate-back-to-tf branch would raise a MethodError there).  interp() here is the cubic
// Hermite interpolant through (uprev, k1) and (u, k2) -- a documented extension.
#pragma once
#include "degk_rosenbrock.cuh"

namespace degk {

template <class T, class Model, bool K5>
struct Kvaerno {
    static constexpr int N = Model::N;
    static constexpr int ORDER = K5 ? 5 : 3;
    static constexpr bool FSAL = false;
    static constexpr bool ALWAYS_SOLVED = false;
    static constexpr int NS = K5 ? 7 : 4;
    static constexpr int NF_ATTEMPT = NS;
    // dt_nl / tb_nl: step and time base the fixed-dt stepper hands to build_nlsolver when a tstop shortened the
    // step -- the nominal integ.dt and the tstop itself (gpu_kvaerno3_perform_step.jl:16-24, 36-41); 0 = unset
    struct Keep { T k1[N]; T k1next[N]; T k2[N]; T dt_nl; T tb_nl; };

    static DEGK_DEV T dtmin() { return (T)1.0e-14f; }     // convert(T, 1.0f-14)
    static DEGK_DEV T land()  { return (T)1.0e-14f; }

    static DEGK_DEV void init(Keep& K, const T (&u0)[N], const T* p, T t0) {
        Model::template f<T>(K.k1, u0, p, t0);             // u_modified = true at the first step
        DEGK_UNROLL for (int c = 0; c < N; ++c) { K.k1next[c] = K.k1[c]; K.k2[c] = (T)0; }
        K.dt_nl = (T)0; K.tb_nl = (T)0;
    }
    static DEGK_DEV void accepted(Keep& K) { DEGK_UNROLL for (int c = 0; c < N; ++c) K.k1[c] = K.k1next[c]; }
    static DEGK_DEV void on_accept(Keep&) {}
    static DEGK_DEV void init_sel(Keep& K, const T (&u0)[N], const T* p, T t0, unsigned m) { if (m & 1u) init(K, u0, p, t0); }
    static DEGK_DEV void accepted_sel(Keep& K, unsigned m) { if (m & 1u) accepted(K); }
    static DEGK_DEV void accepted_if(Keep& K, const bool* acc) { if (acc[0]) accepted(K); }

    // W = -mass_matrix + gamma dt J (nlsolve/type.jl:139); identity: the literal -1 on the diagonal
    template <int MN>
    static DEGK_DEV void set_W(T (&W)[N][N], const T (&J)[N][N], const T (&Mm)[MN][MN], T gdt) {
        if constexpr (MN == N && has_mass_of<Model>::value) {
            DEGK_UNROLL for (int i = 0; i < N; ++i)
                DEGK_UNROLL for (int j = 0; j < N; ++j) W[i][j] = -Mm[i][j] + gdt * J[i][j];
        } else {
            DEGK_UNROLL for (int i = 0; i < N; ++i)
                DEGK_UNROLL for (int j = 0; j < N; ++j) W[i][j] = (i == j) ? (T)-1 + gdt * J[i][j] : gdt * J[i][j];
        }
    }

    // nlsolve/utils.jl:1-26.  z: in = predictor, out = solution; tmp = explicit part of the stage
    static DEGK_DEV bool nlsolve(T (&z)[N], const T (&tmp)[N], T gam, T c, T dt, T tb, const T* p) {
        const T abstol = (T)100 * (sizeof(T) == 4 ? (T)1.1920928955078125e-7f : (T)2.220446049250313e-16);
        const T ts = tb + c * dt;
        const T gdt = gam * dt;
        // a constant mass matrix enters here only: W = -M + gamma dt J, residual dt f - M z (utils.jl:10-21, type.jl:139)
        constexpr bool MASS = has_mass_of<Model>::value;
        T Mm[MASS ? N : 1][MASS ? N : 1];
        if constexpr (MASS) Model::template mass<T>(Mm);
        for (int it = 0; it < 30; ++it) {
            T us[N], J[N][N], W[N][N], fe[N], rhs[N], dz[N], mz[N];
            DEGK_UNROLL for (int i = 0; i < N; ++i) us[i] = tmp[i] + gam * z[i];
            eval_jac<T, Model>(J, us, p, ts);
            set_W(W, J, Mm, gdt);
            Model::template f<T>(fe, us, p, ts);
            if constexpr (MASS) mass_mul<T, N>(Mm, z, mz); else { DEGK_UNROLL for (int i = 0; i < N; ++i) mz[i] = z[i]; }
            DEGK_UNROLL for (int i = 0; i < N; ++i) rhs[i] = dt * fe[i] - mz[i];
            LinSolve<T, N> F;
            if (!F.factor(W)) return false;
            F.solve(rhs, dz);
            DEGK_UNROLL for (int i = 0; i < N; ++i) z[i] = z[i] - dz[i];
            DEGK_UNROLL for (int i = 0; i < N; ++i) us[i] = tmp[i] + gam * z[i];
            Model::template f<T>(fe, us, p, ts);
            if constexpr (MASS) mass_mul<T, N>(Mm, z, mz); else { DEGK_UNROLL for (int i = 0; i < N; ++i) mz[i] = z[i]; }
            T acc = (T)0;
            DEGK_UNROLL for (int i = 0; i < N; ++i) { const T r = dt * fe[i] - mz[i]; acc = (i == 0) ? r * r : acc + r * r; }
            if (sqrt_(acc / (T)N) < abstol) break;           // diffeqgpunorm, src/utils.jl:1
        }
        return true;
    }

    template <bool WANT_ERR>
    static DEGK_DEV bool attempt(Keep& K, const T (&uprev)[N], const T* p, T t, T h,
                                 T (&unew)[N], T (&err)[N]) {
        const bool nominal = !WANT_ERR && K.dt_nl > (T)0;
        const T tb = WANT_ERR ? t : (nominal ? K.tb_nl : t + h);     // nlsolver.t (see the header comment)
        const T hnl = nominal ? K.dt_nl : h;                         // nlsolver.dt
        if (!WANT_ERR) K.dt_nl = (T)0;
        T k1[N];
        if (WANT_ERR) Model::template f<T>(k1, uprev, p, t);  // `k1 = f(uprev, p, t)` inside the retry loop
        else { DEGK_UNROLL for (int c = 0; c < N; ++c) k1[c] = K.k1[c]; }
        T z[NS][N], tmp[N];
        DEGK_UNROLL for (int c = 0; c < N; ++c) z[0][c] = h * k1[c];
        T gam;
        if (!K5) {
            gam = (T)0.4358665215;
            const T a31 = (T)0.490563388419108, a32 = (T)0.073570090080892;
            const T a41 = (T)0.308809969973036, a42 = (T)1.490563388254106, a43 = -(T)1.235239879727145;
            const T c3 = (T)1;
            // alpha31/32 are evaluated in the working precision (kvaerno_tableaus.jl:41-46)
            const T c2 = (T)2 * gam, th = c3 / c2;
            const T th2 = th * th;
            const T w = ((T)6 * th) * ((T)1 - th) / c2;
            const T al31 = ((T)1 + ((T)-4 * th + (T)3 * th2)) + w * gam;
            const T al32 = ((T)-2 * th + (T)3 * th2) + w * gam;
            DEGK_UNROLL for (int c = 0; c < N; ++c) { z[1][c] = z[0][c]; tmp[c] = uprev[c] + gam * z[0][c]; }
            if (!nlsolve(z[1], tmp, gam, gam, hnl, tb, p)) return false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                z[2][c] = al31 * z[0][c] + al32 * z[1][c];
                tmp[c] = (uprev[c] + a31 * z[0][c]) + a32 * z[1][c];
            }
            if (!nlsolve(z[2], tmp, gam, c3, hnl, tb, p)) return false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                z[3][c] = (a31 * z[0][c] + a32 * z[1][c]) + gam * z[2][c];      // yhat as prediction
                tmp[c] = ((uprev[c] + a41 * z[0][c]) + a42 * z[1][c]) + a43 * z[2][c];
            }
            if (!nlsolve(z[3], tmp, gam, (T)1, hnl, tb, p)) return false;
        } else {
            gam = (T)0.26;
            const T a31 = (T)0.13, a32 = (T)0.84033320996790809;
            const T a41 = (T)0.22371961478320505, a42 = (T)0.47675532319799699, a43 = -(T)0.06470895363112615;
            const T a51 = (T)0.16648564323248321, a52 = (T)0.1045001884159172, a53 = (T)0.03631482272098715, a54 = -(T)0.13090704451073998;
            const T a61 = (T)0.13855640231268224, a63 = -(T)0.04245337201752043, a64 = (T)0.02446657898003141, a65 = (T)0.61943039072480676;
            const T a71 = (T)0.13659751177640291, a73 = -(T)0.05496908796538376, a74 = -(T)0.04118626728321046, a75 = (T)0.62993304899016403, a76 = (T)0.06962479448202728;
            const T al31 = (T)-1.366025403784441, al32 = (T)2.3660254037844357;
            const T al41 = (T)-0.19650552613122207, al42 = (T)0.8113579546496623, al43 = (T)0.38514757148155954;
            const T al51 = (T)0.10375304369958693, al52 = (T)0.937994698066431, al53 = (T)-0.04174774176601781;
            const T al61 = (T)-0.17281112873898072, al62 = (T)0.6235784481025847, al63 = (T)0.5492326806363959;
            const T c3 = (T)1.230333209967908, c4 = (T)0.895765984350076, c5 = (T)0.436393609858648, c6 = (T)1;
            DEGK_UNROLL for (int c = 0; c < N; ++c) { z[1][c] = z[0][c]; tmp[c] = uprev[c] + gam * z[0][c]; }
            if (!nlsolve(z[1], tmp, gam, gam, hnl, tb, p)) return false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                z[2][c] = al31 * z[0][c] + al32 * z[1][c];
                tmp[c] = (uprev[c] + a31 * z[0][c]) + a32 * z[1][c];
            }
            if (!nlsolve(z[2], tmp, gam, c3, hnl, tb, p)) return false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                z[3][c] = (al41 * z[0][c] + al42 * z[1][c]) + al43 * z[2][c];
                tmp[c] = ((uprev[c] + a41 * z[0][c]) + a42 * z[1][c]) + a43 * z[2][c];
            }
            if (!nlsolve(z[3], tmp, gam, c4, hnl, tb, p)) return false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                z[4][c] = (al51 * z[0][c] + al52 * z[1][c]) + al53 * z[2][c];
                tmp[c] = (((uprev[c] + a51 * z[0][c]) + a52 * z[1][c]) + a53 * z[2][c]) + a54 * z[3][c];
            }
            if (!nlsolve(z[4], tmp, gam, c5, hnl, tb, p)) return false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                z[5][c] = (al61 * z[0][c] + al62 * z[1][c]) + al63 * z[2][c];
                tmp[c] = (((uprev[c] + a61 * z[0][c]) + a63 * z[2][c]) + a64 * z[3][c]) + a65 * z[4][c];
            }
            if (!nlsolve(z[5], tmp, gam, c6, hnl, tb, p)) return false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                z[6][c] = (((a61 * z[0][c] + a63 * z[2][c]) + a64 * z[3][c]) + a65 * z[4][c]) + gam * z[5][c];
                tmp[c] = ((((uprev[c] + a71 * z[0][c]) + a73 * z[2][c]) + a74 * z[3][c]) + a75 * z[4][c]) + a76 * z[5][c];
            }
            if (!nlsolve(z[6], tmp, gam, (T)1, hnl, tb, p)) return false;
        }
        DEGK_UNROLL for (int c = 0; c < N; ++c) {
            unew[c] = tmp[c] + gam * z[NS - 1][c];
            K.k2[c] = z[NS - 1][c] / h;
        }
        if (WANT_ERR) {
            // W_eval = W(tmp + gamma z_s, p, t + c dt), c = 1;  err = W \ sum(btilde_i z_i)
            T b[N];
            if (!K5) {
                const T bt1 = (T)0.181753418446072, bt2 = (T)-1.416993298173214, bt3 = (T)1.671106401227145, bt4 = -gam;
                DEGK_UNROLL for (int c = 0; c < N; ++c) b[c] = ((bt1 * z[0][c] + bt2 * z[1][c]) + bt3 * z[2][c]) + bt4 * z[3][c];
            } else {
                const T bt1 = (T)0.00195889053627933, bt3 = (T)0.01251571594786333, bt4 = (T)0.06565284626324187;
                const T bt5 = -(T)0.01050265826535727, bt6 = (T)0.19037520551797272, bt7 = -gam;
                DEGK_UNROLL for (int c = 0; c < N; ++c)
                    b[c] = ((((bt1 * z[0][c] + bt3 * z[2][c]) + bt4 * z[3][c]) + bt5 * z[4][c]) + bt6 * z[5][c]) + bt7 * z[6][c];
            }
            T J[N][N], W[N][N];
            const T gdt = gam * h;
            eval_jac<T, Model>(J, unew, p, tb + (T)1 * h);
            constexpr bool MASS = has_mass_of<Model>::value;
            T Mm[MASS ? N : 1][MASS ? N : 1];
            if constexpr (MASS) Model::template mass<T>(Mm);
            set_W(W, J, Mm, gdt);
            LinSolve<T, N> F;
            if (!F.factor(W)) return false;
            F.solve(b, err);
            // integ.k1 = k1 on accept: the slope the interpolant sees (k1next) is the one at uprev
            DEGK_UNROLL for (int c = 0; c < N; ++c) { K.k1[c] = k1[c]; K.k1next[c] = k1[c]; }
        } else {
            // integ.k1 = f(integ.u, p, t) with t the time captured before the step (:78-81)
            Model::template f<T>(K.k1next, unew, p, t);
        }
        return true;
    }

    // The integrators without an `_ode_interpolant` method of their own fall to the default Hermite one
    // (nonstiff/interpolants.jl:1-21, `@muladd`), evaluated with whatever integ.k1 / integ.k2 hold when
    // savevalues! runs: adaptive k1 = f(uprev, p, t), fixed dt k1 = f(u_new, p, t_old) (the FSAL value stored
    // at the end of the step, gpu_kvaerno3_perform_step.jl:78-81) -- K.k1next in both cases -- and k2 = z_s / dt.
    //   (1-Θ) y0 + Θ y1 + Θ (Θ-1) ((1-2Θ)(y1-y0) + (Θ-1) dt k1 + Θ dt k2)
    // MuladdMacro keeps the first product of a sum and folds the later ones in as muladd(prod(1..n-1), last, acc).
    static DEGK_DEV void interp(const Keep& K, T theta, T h, const T (&uprev)[N],
                                const T (&unew)[N], const T* p, T tprev, T (&out)[N]) {
        (void)p; (void)tprev;
        const T th1 = (T)1 - theta, thm = theta - (T)1;
        const T w0 = fma_((T)-2, theta, (T)1);
        const T w1 = thm * h, w2 = theta * h, w3 = theta * thm;
        DEGK_UNROLL for (int c = 0; c < N; ++c) {
            T inner = w0 * (unew[c] - uprev[c]);
            inner = fma_(w1, K.k1next[c], inner);
            inner = fma_(w2, K.k2[c], inner);
            out[c] = fma_(w3, inner, fma_(theta, unew[c], th1 * uprev[c]));
        }
    }
    // the deferred-save replay (degk_ode_saves.cuh) must rebuild the *adaptive* step's k1 / k2
    static constexpr bool REPLAY_ADAPTIVE = true;
};

template <class T, class M> using Kvaerno3M = Kvaerno<T, M, false>;
template <class T, class M> using Kvaerno5M = Kvaerno<T, M, true>;

}  // namespace degk
