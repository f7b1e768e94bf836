// degk_ode_events.cuh -- ensemble kernels with tstops and discrete callbacks.
//
// The event-capable variants of the two ODE drivers, one thread per trajectory:
//   ode_solve_events_body    fixed dt   kernels.jl:1-72   + tstops   gpu_tsit5_perform_step.jl:18-27
//   ode_asolve_events_body   adaptive   kernels.jl:74-152 + tstops   gpu_tsit5_perform_step.jl:155-169
// and, after every step, handle_callbacks! -> apply_discrete_callback! (integrator_utils.jl:69-150,
// 271-330): each GPUDiscreteCallback whose condition holds first saves (savevalues!), then marks
// u_modified (the FSAL stage is recomputed at the next step) and runs its affect!;
// terminate!(integrator) ends the trajectory with ReturnCode.Terminated (:52-66).
//
// These kernels mirror the reference's control flow statement by statement (nested retry loop,
// static thread-to-trajectory mapping) instead of the persistent packed kernel of the headline
// path: callbacks are arbitrary user code, lowered from the Julia closures to CUDA-C bodies and
// spliced in by NVRTC (degk_jit.cpp), so the programs that use them are JIT-built and carry this
// pair of kernels in place of the generation-2/3 adaptive kernel.
//
// Continuous callbacks (GPUContinuousCallback, callbacks.jl:38-124) run first, as in handle_callbacks!:
// find_callback_time (integrator_utils.jl:383-442: sign change of the condition between tprev --
// nudged by repeat_nudge*dt off a root found in the previous step -- and t, dense output in between),
// the hand-written ITP root finder gpu_find_root (:326-381), the earliest event over the set
// (DiffEqBase.find_first_continuous_callback -- not vendored; restated from the published package),
// then apply_callback! (:232-269): move (t, u) back to the event by interpolation, save, mark
// u_modified and run affect! / affect_neg! by the sign before the event.  Two in-tree quirks are
// kept: the affects of continuous_callbacks[1] run whichever callback fired (:296-301), and
// last_event_error is never updated (stays 0).
//
// CB is a struct with
//   static constexpr int NCB, NCC;
//   template <class T> static bool condition(int c, const T (&u)[N], const T* p, T t);
//   template <class T> static void affect(int c, T (&u)[N], T* p, T t, bool& terminate_);
//   template <class T> static T    ccondition(int c, const T (&u)[N], const T* p, T t);
//   template <class T> static void caffect(int c, bool neg, T (&u)[N], T* p, T t, bool& terminate_);
//   static bool has_caffect(int c, bool neg);  static int rootfind(int c);     // 0 Left, 1 Right, 2 None
//   static double cc_abstol(int c), cc_repeat_nudge(int c), cc_dtrelax(int c);
#pragma once
#include "degk_common.cuh"

namespace degk {

enum { RC_TERMINATED = 6 };

struct NoCallbacks {
    static constexpr int NCB = 0, NCC = 0;
    template <class T, int N> static DEGK_DEV bool condition(int, const T (&)[N], const T*, T) { return false; }
    template <class T, int N> static DEGK_DEV void affect(int, T (&)[N], T*, T, bool&) {}
    template <class T, int N> static DEGK_DEV T ccondition(int, const T (&)[N], const T*, T) { return (T)1; }
    template <class T, int N> static DEGK_DEV void caffect(int, bool, T (&)[N], T*, T, bool&) {}
    static DEGK_DEV bool has_caffect(int, bool) { return false; }
    static DEGK_DEV int rootfind(int) { return 0; }
    static DEGK_DEV double cc_abstol(int) { return 0.0; }
    static DEGK_DEV double cc_repeat_nudge(int) { return 0.0; }
    static DEGK_DEV double cc_dtrelax(int) { return 1.0; }
};

// Base.sign, Base.eps(x), nextfloat for the ITP root finder
template <class T> DEGK_DEV T sign_(T x) { return x > (T)0 ? (T)1 : (x < (T)0 ? (T)-1 : x); }
DEGK_DEV float eps_of(float x) {
    if (!finite_(x)) return x - x;
    const float ax = fabsf(x);
    return ax >= 1.17549435e-38f ? ldexpf(1.1920928955078125e-7f, ilogbf(ax)) : 1.401298464e-45f;
}
DEGK_DEV double eps_of(double x) {
    if (!finite_(x)) return x - x;
    const double ax = fabs(x);
    return ax >= 2.2250738585072014e-308 ? ldexp(2.220446049250313e-16, ilogb(ax)) : 4.9406564584124654e-324;
}
DEGK_DEV float next_up(float x) { return nextafterf(x, __int_as_float(0x7f800000)); }
DEGK_DEV double next_up(double x) { return nextafter(x, __longlong_as_double(0x7ff0000000000000LL)); }
DEGK_DEV float copysign_(float a, float b) { return copysignf(a, b); }
DEGK_DEV double copysign_(double a, double b) { return copysign(a, b); }

// gpu_find_root (integrator_utils.jl:326-381): ITP with scaled_k1 = 0.2, k2 = 2, n0 = 10
template <class T, class F>
DEGK_DEV T itp_root(F&& fz, T left, T right, int rootfind) {
    T fl = fz(left), fr = fz(right);
    const T span0 = right - left;
    const T k1 = (T)0.2 / span0;
    T eps_s = span0 * (T)512;
    for (int it = 0; it < 100; ++it) {
        const T span = right - left;
        const T mid = (left + right) / (T)2;
        const T r = eps_s - span / (T)2;
        const T x_f = left + span * fl / (fl - fr);
        const T delta = jl_max(k1 * span * span, eps_of(x_f));
        const T diff = mid - x_f;
        const T xt = (delta <= abs_(diff)) ? x_f + copysign_(delta, diff) : mid;
        const T xp = (abs_(xt - mid) <= r) ? xt : mid - copysign_(r, diff);
        const T yp = fz(xp);
        const T yps = yp * sign_(fr);
        if (yps > (T)0) { right = xp; fr = yp; }
        else if (yps < (T)0) { left = xp; fl = yp; }
        else { left = xp; right = xp; break; }
        eps_s = eps_s / (T)2;
        if (next_up(left) >= right) break;
    }
    return rootfind == 0 ? left : right;
}

// handle_callbacks!, continuous part.  hdt = integ.dt (the step the dense output belongs to); dtnew is
// the adaptive integrator's next step (nullptr for fixed dt).  Returns saved_in_cb.
// fixed-dt steppers that build their nonlinear solver from integ.dt / integ.t (not the tstop-shortened local dt)
// keep both in their Keep; everyone else ignores the call
template <class K, class T>
DEGK_DEV auto set_nominal_step(K& k, T dt, T tnew, int) -> decltype((void)(k.dt_nl = dt)) { k.dt_nl = dt; k.tb_nl = tnew; }
template <class K, class T>
DEGK_DEV void set_nominal_step(K&, T, T, long) {}

template <class T, class Model, class Method, class CB, class SaveF>
DEGK_DEV bool handle_continuous(const typename Method::Keep& K, T (&u)[Model::N], const T (&uprev)[Model::N], T* p,
                                T& t, T tprev, T hdt, T tf, T* dtnew, i64& step_idx, int& event_last_time,
                                bool& u_modified, bool& terminated, SaveF&& savevalues) {
    constexpr int N = Model::N;
    const T last_event_error = (T)0;
    auto get_condition = [&](int c, T abst) -> T {       // :460-479
        if (abst == t) return CB::ccondition(c, u, p, abst);
        if (abst == tprev) return CB::ccondition(c, uprev, p, abst);
        T v[N];
        Method::interp(K, (abst - tprev) / hdt, hdt, uprev, u, p, tprev, v);
        return CB::ccondition(c, v, p, abst);
    };
    bool occurred = false;
    T tmin = t, up = (T)0;
    int idx = 0;
    DEGK_UNROLL for (int c = 0; c < CB::NCC; ++c) {
        T bottom_t = tprev;
        T bottom_condition = CB::ccondition(c, uprev, p, tprev);
        if (event_last_time == c + 1 && abs_(bottom_condition - last_event_error) <= (T)CB::cc_abstol(c)) {
            bottom_t = tprev + hdt * (T)CB::cc_repeat_nudge(c);
            bottom_condition = get_condition(c, bottom_t);
        }
        const T bottom_sign = sign_(bottom_condition);
        const T top_t = t;
        const T top_sign = sign_(get_condition(c, top_t));
        const bool ev = ((bottom_sign < (T)0 && CB::has_caffect(c, false)) || (bottom_sign > (T)0 && CB::has_caffect(c, true))) &&
                        bottom_sign * top_sign <= (T)0;
        if (ev) {
            T cbt;
            if (CB::rootfind(c) == 2 || top_sign == (T)0) cbt = top_t;
            else cbt = itp_root<T>([&](T x) { return get_condition(c, x); }, bottom_t, top_t, CB::rootfind(c));
            if (cbt < tmin || !occurred) { tmin = cbt; up = bottom_sign; occurred = true; idx = c + 1; }
        }
    }
    if (!occurred) { event_last_time = 0; return false; }
    event_last_time = idx;
    if (tmin != t) {                                     // change_t_via_interpolation!, :186-206
        T v[N];
        Method::interp(K, (tmin - tprev) / hdt, hdt, uprev, u, p, tprev, v);
        step_idx -= (i64)rint((double)((t - tmin) / hdt));
        DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = v[c];
        t = tmin;
    }
    if (dtnew != nullptr && *dtnew < (T)1.0e-12) {        // :240-249
        const T remaining = abs_(tf - t);
        *dtnew = jl_min((T)CB::cc_dtrelax(0) * hdt, remaining);
    }
    savevalues();
    u_modified = true;
    if (up < (T)0) { if (!CB::has_caffect(0, false)) u_modified = false; else CB::caffect(0, false, u, p, t, terminated); }
    else if (up > (T)0) { if (!CB::has_caffect(0, true)) u_modified = false; else CB::caffect(0, true, u, p, t, terminated); }
    return true;
}

// eps(T) of `T(100) * eps(T)` in the tstops test
template <class T> DEGK_DEV T eps_();
template <> DEGK_DEV float eps_<float>() { return 1.1920928955078125e-7f; }
template <> DEGK_DEV double eps_<double>() { return 2.220446049250313e-16; }

// `tstops[idx] - integ.t - integ.dt - T(100) * eps(T) < T(0)` (integ.t is still the old time)
template <class T>
DEGK_DEV bool tstop_hit(const KArgs& a, int idx, T told, T h) {
    if (idx >= a.n_tstops) return false;
    const T ts = ((const T*)a.tstops)[idx];
    return (ts - told - h - (T)100 * eps_<T>()) < (T)0;
}

// =====================================================================================
// fixed time step
// =====================================================================================
template <class T, class Model, class Method, class CB>
DEGK_DEV void ode_solve_events_body(const KArgs& a) {
    constexpr int N = Model::N;
    const i64 traj = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 nsteps = 0, nfail = 0;
    if (traj < a.n_traj) {
        T u[N], uprev[N], unew[N], err[N];
        T p[Model::NP > 0 ? Model::NP : 1];
        T t0, tf;
        load_problem<T, Model>(a, traj, u, p, t0, tf);
        const T dt = (T)a.dt;                // integ.dt: the nominal step, also when a tstop shortens one (Q4)
        const bool has_saveat = a.saveat != nullptr;
        const T* saveat = (const T*)a.saveat + (has_saveat ? traj * a.saveat_stride : 0);      // this trajectory's grid
        typename Method::Keep K;
        int cur = 0;                 // 1-based index of the next saveat entry
        i64 step_idx = 1;            // 0-based row of the next every-step save
        i64 ts_written = 0;
        if (has_saveat) {            // kernels.jl:34-47
            cur = 1;
            if (t0 == saveat[0]) { cur = 2; store_u<T, N>(a, traj, 0, u); store_t<T>(a, traj, 0, t0); }
        } else {
            store_t<T>(a, traj, 0, t0);
            store_u<T, N>(a, traj, 0, u);
            ts_written = 1;
        }
        Method::init(K, u, p, t0);
        T t = t0, tprev = t0;
        int rc = RC_SUCCESS;
        int tstops_idx = 0;
        bool first = true, u_modified = false, terminated = false;
        int event_last_time = 0;
        i64 iters = 0;
        // savevalues! (integrator_utils.jl:13-50)
        auto savevalues = [&]() {
            if (!has_saveat) {
                if (a.save_everystep) {
                    store_u<T, N>(a, traj, step_idx, u);
                    store_t<T>(a, traj, step_idx, t);
                    ++step_idx;
                    ts_written = step_idx < a.n_rows ? step_idx : a.n_rows;
                }
            } else {
                while (cur <= a.n_saveat && saveat[cur - 1] <= t) {
                    const T savet = saveat[cur - 1];
                    const T theta = (savet - tprev) / dt;
                    T v[N];
                    Method::interp(K, theta, dt, uprev, u, p, tprev, v);
                    store_u<T, N>(a, traj, cur - 1, v);
                    store_t<T>(a, traj, cur - 1, savet);
                    ++cur;
                }
            }
        };
        while (t < tf && !terminated) {
            if (!first) {                        // FSAL shift deferred so the last step's stages
                if (u_modified) Method::init(K, u, p, t);   // survive for the final interpolation;
                else Method::accepted(K);                   // after an affect! k1 = f(u, p, t) again
            }
            first = false; u_modified = false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) uprev[c] = u[c];
            tprev = t;
            T h = dt;
            if (tstop_hit<T>(a, tstops_idx, t, dt)) {
                t = ((const T*)a.tstops)[tstops_idx];
                h = t - tprev;
                ++tstops_idx;
            } else {
                t = t + dt;                      // integ.t += dt precedes the stages
            }
            set_nominal_step(K, dt, t, 0);       // steppers whose nonlinear solver keeps integ.dt / integ.t (Kvaerno)
            if (!Method::template attempt<false>(K, uprev, p, tprev, h, unew, err)) {
                rc = RC_SINGULAR; ++nfail; break;
            }
            Method::on_accept(K);
            DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = unew[c];
            ++nsteps;
            bool saved_in_cb = false;
            if (CB::NCC > 0)
                saved_in_cb = handle_continuous<T, Model, Method, CB>(K, u, uprev, p, t, tprev, dt, tf, (T*)nullptr, step_idx,
                                                                      event_last_time, u_modified, terminated, savevalues);
            if (CB::NCB > 0) saved_in_cb = false;    // handle_callbacks! returns the discrete pass's flag (:318-326)
            DEGK_UNROLL for (int c = 0; c < CB::NCB; ++c) {
                if (CB::condition(c, u, p, t)) {
                    savevalues();
                    saved_in_cb = true;
                    u_modified = true;
                    CB::affect(c, u, p, t, terminated);
                }
            }
            if (!saved_in_cb) savevalues();
            if (++iters >= a.max_iters) { rc = RC_MAXITERS; ++nfail; break; }
        }
        if (rc == RC_SUCCESS) {
            if (t > tf && !has_saveat) {         // kernels.jl:53-57
                const T theta = (tf - tprev) / dt;
                T v[N];
                Method::interp(K, theta, dt, uprev, u, p, tprev, v);
                store_u<T, N>(a, traj, a.n_rows - 1, v);
                store_t<T>(a, traj, a.n_rows - 1, tf);
            }
            if (!has_saveat && !a.save_everystep) {   // kernels.jl:59-62
                store_u<T, N>(a, traj, 1, u);
                store_t<T>(a, traj, 1, t);
                ts_written = 2;
            }
            bool fin = true;
            DEGK_UNROLL for (int c = 0; c < N; ++c) fin = fin && finite_(u[c]);
            if (terminated) rc = RC_TERMINATED;
            else if (!fin) { rc = RC_UNSTABLE; ++nfail; }
        }
        // rows this trajectory did not reach keep t0; the overshoot interpolation above may have
        // written the last row, which must survive
        {
            const i64 first_unwritten = has_saveat ? (i64)(cur - 1) : ts_written;
            const bool last_written = rc != RC_SINGULAR && rc != RC_MAXITERS && t > tf && !has_saveat;
            if (a.ts != nullptr)
                for (i64 k = first_unwritten; k < a.n_rows - (last_written ? 1 : 0); ++k) store_t<T>(a, traj, k, t0);
        }
        if (a.retcode) a.retcode[traj] = rc;
        if (a.naccept) a.naccept[traj] = (int)nsteps;
        if (a.nreject) a.nreject[traj] = 0;
    }
    add_totals<T>(a, nsteps, 0u, nfail);
}

// =====================================================================================
// adaptive time step (nested retry loop like gpu_tsit5_perform_step.jl:101)
// =====================================================================================
template <class T, class Model, class Method, class CB>
DEGK_DEV void ode_asolve_events_body(const KArgs& a) {
    constexpr int N = Model::N;
    typedef Ctl<T, Method::ORDER> C;
    const T abstol = (T)a.abstol, reltol = (T)a.reltol;
    const i64 traj = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 nacc = 0, nrej = 0, nfail = 0;
    if (traj < a.n_traj) {
        T u[N], uprev[N], unew[N], err[N];
        T p[Model::NP > 0 ? Model::NP : 1];
        T t0, tf;
        load_problem<T, Model>(a, traj, u, p, t0, tf);
        const bool has_saveat = a.saveat != nullptr;
        const T* saveat = (const T*)a.saveat + (has_saveat ? traj * a.saveat_stride : 0);      // this trajectory's grid
        typename Method::Keep K;
        int cur = 0;
        if (has_saveat) {            // kernels.jl:116-126
            cur = 1;
            if (t0 == saveat[0]) { cur = 2; store_u<T, N>(a, traj, 0, u); store_t<T>(a, traj, 0, t0); }
        } else {
            store_t<T>(a, traj, 0, t0);
            store_u<T, N>(a, traj, 0, u);
        }
        Method::init(K, u, p, t0);
        T t = t0, tprev = t0, h = (T)a.dt, step_dt = (T)a.dt, qold = C::qoldinit();
        int rc = RC_DEFAULT;
        int tstops_idx = 0;
        bool first = true, u_modified = false, terminated = false;
        int event_last_time = 0;
        i64 dummy_step_idx = 1;      // adaptive integrators do not save every step (Q3)
        i64 iters = 0;
        auto savevalues = [&]() {    // adaptive integrators are built with save_everystep = false (Q3)
            if (has_saveat) {
                while (cur <= a.n_saveat && saveat[cur - 1] <= t) {
                    const T savet = saveat[cur - 1];
                    const T theta = (savet - tprev) / step_dt;
                    T v[N];
                    Method::interp(K, theta, step_dt, uprev, u, p, tprev, v);
                    store_u<T, N>(a, traj, cur - 1, v);
                    store_t<T>(a, traj, cur - 1, savet);
                    ++cur;
                }
            }
        };
        while (t < tf && !terminated && rc == RC_DEFAULT) {
            if (!first) {
                if (u_modified) Method::init(K, u, p, t);
                else Method::accepted(K);
            }
            first = false; u_modified = false;
            DEGK_UNROLL for (int c = 0; c < N; ++c) uprev[c] = u[c];
            const T tcur = t;
            for (;;) {                                       // `while EEst > 1`
                if (h < Method::dtmin()) { rc = RC_DT_LESS_THAN_MIN; break; }   // `dt < dtmin && error(...)`
                if (!Method::template attempt<true>(K, uprev, p, tcur, h, unew, err)) { rc = RC_SINGULAR; break; }
                // tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol); ODE_DEFAULT_NORM
                T acc = (T)0;
                DEGK_UNROLL for (int c = 0; c < N; ++c) {
                    const T sc = abstol + jl_max(abs_(uprev[c]), abs_(unew[c])) * reltol;
                    const T v = ctl_div(err[c], sc);
                    const T sq = v * v;
                    acc = (c == 0) ? sq : acc + sq;
                }
                const T EEst = sqrt_(mean_<T, N>(acc));
                T q, q11 = (T)0;
                if (EEst == (T)0) {
                    q = (T)1 / C::qmax();
                } else {
                    q11 = pow_(EEst, C::beta1());
                    q = ctl_div(q11, pow_(qold, C::beta2()));
                }
                if (EEst > (T)1) {                           // reject
                    h = ctl_div(h, jl_min((T)1 / C::qmin(), ctl_div(q11, C::gamma())));
                    ++nrej;
                    continue;
                }
                q = jl_max((T)1 / C::qmax(), jl_min((T)1 / C::qmin(), ctl_div(q, C::gamma())));
                qold = jl_max(EEst, C::qoldinit());
                T dtnew = ctl_div(h, q);
                dtnew = jl_min(abs_(dtnew), abs_(tf - tcur - h));
                step_dt = h;                                 // integ.dt
                tprev = tcur;
                Method::on_accept(K);
                DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = unew[c];
                if ((tf - tcur - h) < Method::land()) {
                    t = tf;
                } else if (tstop_hit<T>(a, tstops_idx, tcur, h)) {
                    // integ.t = tstop; integ.u = integ(integ.t): dense output of the step just taken
                    t = ((const T*)a.tstops)[tstops_idx];
                    T v[N];
                    Method::interp(K, (t - tprev) / step_dt, step_dt, uprev, unew, p, tprev, v);
                    DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = v[c];
                    ++tstops_idx;
                } else {
                    t = tcur + h;
                    if (t == tcur && (tf - tcur - h) <= h) t = tf;   // see DESIGN.md, deviations (sub-ulp remaining span)
                }
                h = dtnew;
                ++nacc;
                break;
            }
            if (rc != RC_DEFAULT) break;
            bool saved_in_cb = false;
            if (CB::NCC > 0)
                saved_in_cb = handle_continuous<T, Model, Method, CB>(K, u, uprev, p, t, tprev, step_dt, tf, &h, dummy_step_idx,
                                                                      event_last_time, u_modified, terminated, savevalues);
            if (CB::NCB > 0) saved_in_cb = false;
            DEGK_UNROLL for (int c = 0; c < CB::NCB; ++c) {
                if (CB::condition(c, u, p, t)) {
                    savevalues();
                    saved_in_cb = true;
                    u_modified = true;
                    CB::affect(c, u, p, t, terminated);
                }
            }
            if (!saved_in_cb) savevalues();
            if (++iters >= a.max_iters) { rc = RC_MAXITERS; break; }
        }
        if (rc == RC_DEFAULT) {
            if (t > tf && !has_saveat) {                     // kernels.jl:133-137
                const T theta = (tf - tprev) / step_dt;
                T v[N];
                Method::interp(K, theta, step_dt, uprev, u, p, tprev, v);
                store_u<T, N>(a, traj, a.n_rows - 1, v);
                store_t<T>(a, traj, a.n_rows - 1, tf);
            }
            if (!has_saveat && !a.save_everystep) {          // kernels.jl:139-142
                store_u<T, N>(a, traj, 1, u);
                store_t<T>(a, traj, 1, t);
            }
            bool fin = true;
            DEGK_UNROLL for (int c = 0; c < N; ++c) fin = fin && finite_(u[c]);
            rc = terminated ? (int)RC_TERMINATED : (fin ? (int)RC_SUCCESS : (int)RC_UNSTABLE);
        }
        if (rc != RC_SUCCESS && rc != RC_TERMINATED) ++nfail;
        if (a.ts != nullptr) {
            // unwritten rows keep t0 (lowerlevel_solve.jl:318 fill!)
            i64 first_unwritten;
            if (has_saveat) first_unwritten = cur - 1;
            else first_unwritten = (rc == RC_SUCCESS || rc == RC_TERMINATED) && !a.save_everystep ? 2 : 1;
            const bool last_written = (rc == RC_SUCCESS || rc == RC_TERMINATED) && t > tf && !has_saveat;
            for (i64 k = first_unwritten; k < a.n_rows - (last_written ? 1 : 0); ++k) store_t<T>(a, traj, k, t0);
        }
        if (a.retcode) a.retcode[traj] = rc;
        if (a.naccept) a.naccept[traj] = (int)nacc;
        if (a.nreject) a.nreject[traj] = (int)nrej;
    }
    add_totals<T>(a, nacc, nrej, nfail);
}

}  // namespace degk
