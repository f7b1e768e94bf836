// degk_ode_kernels.cuh -- the two ODE ensemble kernels (one thread = one trajectory).
//
//   ode_solve_kernel   fixed dt      replaces reference kernels.jl:1-72  (ode_solve_kernel)
//   ode_asolve_kernel  adaptive dt   replaces reference kernels.jl:74-152 (ode_asolve_kernel)
//
// Both are generic in <T, Model, Method>; Method is one of the generated explicit RK structs
// (gen_erk_*.cuh) or a Rosenbrock struct (degk_rosenbrock.cuh) exposing
//   Keep, init(), attempt<WANT_ERR>(), accepted(), interp(), dtmin(), land(), ORDER.
//
// B200 design notes (vs the reference's KernelAbstractions kernel):
//  * The reference nests two data-dependent loops (steps, and retries inside step!,
//    gpu_tsit5_perform_step.jl:101).  Here the adaptive kernel is ONE flat loop whose body is a
//    single attempt; accept/reject bookkeeping is predicated.  A warp therefore never waits on
//    a lane that is retrying.
//  * Adaptive step counts differ per trajectory.  With SCHED_QUEUE the kernel is persistent:
//    a lane that finishes its trajectory claims the next unclaimed index from a global
//    counter (one warp-aggregated atomicAdd per refill), so lanes stay busy until the queue
//    drains instead of idling until the slowest lane of their warp finishes.
//  * State, parameters and all stage vectors live in registers; tableau coefficients are
//    immediates; saved states go straight to HBM.  In the REF layout each trajectory owns a
//    contiguous (len*n*sizeof T)-byte strip that is filled over the trajectory's lifetime, so the
//    sectors are merged in the 126 MB L2 before they are written back.
#pragma once
#include "degk_common.cuh"
#include "degk_dae_init.cuh"

namespace degk {

// =====================================================================================
// fixed time step
// =====================================================================================
// Every-step saves in the reference layout are 12-byte stores at a stride of len*n*sizeof(T)
// bytes across the lanes of a warp: 32 sectors per store instruction.  With a.stage_rows = R > 0
// (host: REF layout, save_everystep, no saveat) each lane buffers R rows of (u, t) in shared
// memory and the warp then writes every trajectory's R*n contiguous values with consecutive
// lanes: the same number of store instructions, one eighth of the sectors.
template <class T, class Model, class Method>
DEGK_DEV void ode_solve_body(const KArgs& a, unsigned char* smem_raw) {
    constexpr int N = Model::N;
    const i64 traj = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = traj < a.n_traj;
    const u32 lane = lane_id();
    u32 nsteps = 0, nfail = 0;
    T u[N], uprev[N], unew[N], err[N];
    T p[Model::NP > 0 ? Model::NP : 1];
    T t0 = (T)0, tf = (T)0;
    const T dt = (T)a.dt;
    const bool has_saveat = a.saveat != nullptr;
    const T* saveat = (const T*)a.saveat + (has_saveat && valid ? traj * a.saveat_stride : 0);   // this trajectory's grid
    typename Method::Keep K;
    int cur = 0;                 // 1-based index of the next saveat entry
    i64 step_idx = 1;            // 0-based row of the next every-step save
    i64 ts_written = 0;          // rows [0, ts_written) of ts hold real values
    T t = (T)0, tprev = (T)0;
    int rc = RC_SUCCESS;
    bool first = true;
    i64 iters = 0;
    // staging buffers of this warp: [32 lanes][N*R + 1] state words, then [32 lanes][R + 1] times
    const int R = a.stage_rows;
    T* wu = nullptr; T* wt = nullptr;
    int nbuf = 0; i64 k0 = 0;
    if (R > 0) {
        const size_t per_warp = (size_t)32 * ((size_t)N * R + 1 + R + 1);
        wu = (T*)smem_raw + (size_t)(threadIdx.x >> 5) * per_warp;
        wt = wu + (size_t)32 * ((size_t)N * R + 1);
    }
    auto flush = [&]() {         // warp-cooperative: every lane of the warp calls it
        __syncwarp();
        // lock-step warp (the usual case: same tspan and dt for 32 consecutive trajectories): one flattened copy of
        // the [32 trajectories][n rows][N] block -- lane-consecutive words, one index update per 32 words
        const int n0 = __shfl_sync(0xffffffffu, nbuf, 0);
        const i64 kk0 = __shfl_sync(0xffffffffu, k0, 0), tr0 = __shfl_sync(0xffffffffu, traj, 0);
        if (__all_sync(0xffffffffu, nbuf == n0 && k0 == kk0 && traj == tr0 + (i64)lane) && n0 > 0 && kk0 + n0 <= a.n_rows) {
            const int L = n0 * N, S = N * R + 1;
            const size_t rs = (size_t)a.n_rows * N;
            T* out = (T*)a.us + ((size_t)tr0 * a.n_rows + (size_t)kk0) * N;
            int j = 0, w = (int)lane;
            while (w >= L) { w -= L; ++j; }
            for (int it = 0; it < L; ++it) {
                out[(size_t)j * rs + w] = wu[j * S + w];
                w += 32;
                while (w >= L) { w -= L; ++j; }
            }
            if (a.ts != nullptr) {
                T* tout = (T*)a.ts + (size_t)tr0 * a.n_rows + (size_t)kk0;
                j = 0; w = (int)lane;
                while (w >= n0) { w -= n0; ++j; }
                for (int it = 0; it < n0; ++it) {
                    tout[(size_t)j * a.n_rows + w] = wt[j * (R + 1) + w];
                    w += 32;
                    while (w >= n0) { w -= n0; ++j; }
                }
            }
        } else {
            for (int j = 0; j < 32; ++j) {
                const int n = __shfl_sync(0xffffffffu, nbuf, j);
                if (n == 0) continue;
                const i64 tr = __shfl_sync(0xffffffffu, traj, j);
                const i64 kk = __shfl_sync(0xffffffffu, k0, j);
                i64 nn = a.n_rows - kk;          // rows past len are dropped (the reference would write out of bounds)
                nn = nn < 0 ? 0 : (nn < n ? nn : n);
                const T* ub = wu + (size_t)j * ((size_t)N * R + 1);
                T* ud = (T*)a.us + (tr * a.n_rows + kk) * N;
                for (int w = (int)lane; w < (int)nn * N; w += 32) ud[w] = ub[w];
                if (a.ts != nullptr && (i64)lane < nn) ((T*)a.ts)[tr * a.n_rows + kk + lane] = wt[(size_t)j * (R + 1) + lane];
            }
        }
        nbuf = 0;
        __syncwarp();
    };
    bool init_failed = false;
    if (valid) {
        load_problem<T, Model>(a, traj, u, p, t0, tf);
        // DAE initialisation (kernels.jl:19-25: SimpleTrustRegion, tolerances 1e-6); row 1 of the every-step path
        // is prob.u0, the saveat path stores the initialised value (kernels.jl:40 vs :44)
        T u_given[N];
        DEGK_UNROLL for (int c = 0; c < N; ++c) u_given[c] = u[c];
        if (a.reserved & 2) init_failed = !dae_initialize<T, Model>(u, p, t0, (T)1.0e-6, (T)1.0e-6);
        if (init_failed) {                       // kernels.jl:63-70: store the initial values and bail out
            store_u<T, N>(a, traj, 0, u_given); store_t<T>(a, traj, 0, t0);
            ts_written = 1; cur = 2;
            if (!has_saveat && !a.save_everystep) { store_u<T, N>(a, traj, 1, u_given); store_t<T>(a, traj, 1, t0); ts_written = 2; }
            rc = RC_INIT_FAILURE; ++nfail;
        } else if (has_saveat) {                 // kernels.jl:34-47
            cur = 1;
            if (t0 == saveat[0]) { cur = 2; store_u<T, N>(a, traj, 0, u); store_t<T>(a, traj, 0, t0); }
        } else {
            store_t<T>(a, traj, 0, t0);
            store_u<T, N>(a, traj, 0, u_given);
            ts_written = 1;
        }
        Method::init(K, u, p, t0);
        t = t0; tprev = t0;
    }
    bool active = valid && !init_failed && (t < tf);
    while (__any_sync(0xffffffffu, active)) {
        if (active) {
            if (!first) Method::accepted(K);     // FSAL shift deferred so the last step's
            first = false;                       // stages survive for the final interpolation
            DEGK_UNROLL for (int c = 0; c < N; ++c) uprev[c] = u[c];
            tprev = t;
            t = t + dt;                          // integ.t += dt precedes the stages
            if (!Method::template attempt<false>(K, uprev, p, tprev, dt, unew, err)) {
                rc = RC_SINGULAR; ++nfail; active = false;
            } else {
                Method::on_accept(K);
                DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = unew[c];
                ++nsteps;
                if (!has_saveat) {
                    if (a.save_everystep) {          // integrator_utils.jl:28-33
                        if (R > 0) {
                            if (nbuf == 0) k0 = step_idx;
                            T* ub = wu + (size_t)lane * ((size_t)N * R + 1) + (size_t)nbuf * N;
                            DEGK_UNROLL for (int c = 0; c < N; ++c) ub[c] = u[c];
                            wt[(size_t)lane * (R + 1) + nbuf] = t;
                            ++nbuf;
                        } else {
                            store_u<T, N>(a, traj, step_idx, u);
                            store_t<T>(a, traj, step_idx, t);
                        }
                        ++step_idx;
                        ts_written = step_idx;
                    }
                } else {                             // integrator_utils.jl:34-47
                    while (cur <= a.n_saveat && saveat[cur - 1] <= t) {
                        const T savet = saveat[cur - 1];
                        const T theta = (savet - tprev) / dt;
                        T v[N];
                        Method::interp(K, theta, dt, uprev, u, p, tprev, v);
                        store_u<T, N>(a, traj, cur - 1, v);
                        store_t<T>(a, traj, cur - 1, savet);
                        ts_written = cur;
                        ++cur;
                    }
                }
                if (++iters >= a.max_iters) { rc = RC_MAXITERS; ++nfail; active = false; }
                else active = t < tf;
            }
        }
        if (R > 0 && __any_sync(0xffffffffu, nbuf >= R)) flush();
    }
    if (R > 0) flush();
    if (valid) {
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
            if (!fin) { rc = RC_UNSTABLE; ++nfail; }
        }
        fill_unwritten_ts<T>(a, traj, has_saveat ? (i64)(cur - 1) : ts_written, t0);
        if (a.retcode) a.retcode[traj] = rc;
        if (a.naccept) a.naccept[traj] = (int)nsteps;
        if (a.nreject) a.nreject[traj] = 0;
    }
    add_totals<T>(a, nsteps, 0u, nfail);
}

// =====================================================================================
// adaptive time step: flat attempt loop + optional persistent lane refill
// =====================================================================================
template <class T, class Model, class Method>
DEGK_DEV void ode_asolve_body(const KArgs& a) {
    constexpr int N = Model::N;
    typedef Ctl<T, Method::ORDER> C;
    const T abstol = (T)a.abstol, reltol = (T)a.reltol;
    const T* saveat = (const T*)a.saveat;              // re-pointed at the trajectory's own grid when it is claimed
    const bool has_saveat = saveat != nullptr;
    const u32 lane = lane_id();
    const u32 lt_mask = (1u << lane) - 1u;

    // per-lane trajectory state
    T u[N], unew[N], err[N];
    T p[Model::NP > 0 ? Model::NP : 1];
    typename Method::Keep K;
    T t = (T)0, t0 = (T)0, tf = (T)0, h = (T)0, qold = (T)0, next_save = (T)0;
    int cur = 0;
    u32 nacc = 0, nrej = 0;              // current trajectory
    u32 tot_acc = 0, tot_rej = 0, tot_fail = 0;
    i64 iters = 0;
    i64 traj = -1;
    bool have = false;
    // SCHED_STATIC: every thread integrates exactly trajectory <global thread id>.
    // SCHED_QUEUE : persistent kernel, trajectories are claimed from a.work_counter (starts at 0).
    const bool queue = (a.schedule == SCHED_QUEUE);
    bool exhausted = false;                            // warp-uniform
    i64 claim = a.n_traj;

    for (;;) {
        // ---------------- (re)fill idle lanes ----------------
        const u32 need = __ballot_sync(0xffffffffu, !have);
        if (need) {
            if (exhausted) {
                claim = a.n_traj;    // nothing left
            } else if (queue) {
                const int cnt = __popc(need);
                const int leader = __ffs(need) - 1;
                i64 base = 0;
                if ((int)lane == leader) base = (i64)atomicAdd(a.work_counter, (u64)cnt);
                base = __shfl_sync(0xffffffffu, base, leader);
                claim = base + __popc(need & lt_mask);
                if (base + cnt >= a.n_traj) exhausted = true;
            } else {
                claim = (i64)blockIdx.x * blockDim.x + threadIdx.x;
                exhausted = true;
            }
            if (!have && claim < a.n_traj) {
                traj = claim;
                claim = a.n_traj;
                have = true;
                load_problem<T, Model>(a, traj, u, p, t0, tf);
                bool init_ok = true;
                if (a.reserved & 2) {                    // kernels.jl:93-99 (tolerances of the solve)
                    T u_given[N];
                    DEGK_UNROLL for (int c = 0; c < N; ++c) u_given[c] = u[c];
                    init_ok = dae_initialize<T, Model>(u, p, t0, abstol, reltol);
                    if (!init_ok) {                      // kernels.jl:143-150
                        store_u<T, N>(a, traj, 0, u_given); store_t<T>(a, traj, 0, t0);
                        if (!has_saveat && !a.save_everystep) { store_u<T, N>(a, traj, 1, u_given); store_t<T>(a, traj, 1, t0); }
                        fill_unwritten_ts<T>(a, traj, (!has_saveat && !a.save_everystep) ? 2 : 1, t0);
                        if (a.retcode) a.retcode[traj] = RC_INIT_FAILURE;
                        if (a.naccept) a.naccept[traj] = 0;
                        if (a.nreject) a.nreject[traj] = 0;
                        ++tot_fail;
                        have = false;
                    }
                }
                if (init_ok) {
                t = t0;
                h = (T)a.dt;
                qold = C::qoldinit();
                nacc = 0; nrej = 0; iters = 0;
                // kernels.jl:116-126
                cur = 0;
                if (has_saveat) {
                    saveat = (const T*)a.saveat + traj * a.saveat_stride;
                    cur = 1;
                    if (t0 == saveat[0]) {
                        cur = 2;
                        store_u<T, N>(a, traj, 0, u);
                        store_t<T>(a, traj, 0, t0);
                    }
                    next_save = (cur <= a.n_saveat) ? saveat[cur - 1] : (T)0;
                } else {
                    store_t<T>(a, traj, 0, t0);
                    store_u<T, N>(a, traj, 0, u);
                    fill_unwritten_ts<T>(a, traj, 1, t0);   // rows 1.. start as t0 (lowerlevel_solve.jl:318)
                }
                Method::init(K, u, p, t0);
                if (!(t < tf)) {     // empty time span: nothing to integrate
                    if (!has_saveat && !a.save_everystep) { store_u<T, N>(a, traj, 1, u); store_t<T>(a, traj, 1, t); }
                    if (a.retcode) a.retcode[traj] = RC_SUCCESS;
                    if (a.naccept) a.naccept[traj] = 0;
                    if (a.nreject) a.nreject[traj] = 0;
                    have = false;
                }
                }   // init_ok
            }
            if (__all_sync(0xffffffffu, !have)) {
                if (exhausted) break;
                continue;
            }
        }

        // ---------------- one attempt ----------------
        if (have) {
            int rc = RC_DEFAULT;
            if (h < Method::dtmin()) {                       // `dt < dtmin && error(...)`
                rc = RC_DT_LESS_THAN_MIN;
            } else if (!Method::template attempt<true>(K, u, p, t, h, unew, err)) {
                rc = RC_SINGULAR;
            } else {
                // tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol); ODE_DEFAULT_NORM
                T acc = (T)0;
                DEGK_UNROLL for (int c = 0; c < N; ++c) {
                    const T sc = abstol + jl_max(abs_(u[c]), abs_(unew[c])) * reltol;
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
                } else {                                     // accept
                    q = jl_max((T)1 / C::qmax(), jl_min((T)1 / C::qmin(), ctl_div(q, C::gamma())));
                    qold = jl_max(EEst, C::qoldinit());
                    T dtnew = ctl_div(h, q);
                    dtnew = jl_min(abs_(dtnew), abs_(tf - t - h));
                    const T tprev = t;
                    // a step that cannot advance t (remaining span below ulp(t)) lands on tf: the
                    // reference would loop forever here (deviations)
                    T tnew = ((tf - t - h) < Method::land()) ? tf : t + h;
                    if (tnew == t && (tf - t - h) <= h) tnew = tf;
                    ++nacc;
                    Method::on_accept(K);
                    if (has_saveat) {                        // integrator_utils.jl:34-47
                        while (cur <= a.n_saveat && next_save <= tnew) {
                            const T theta = (next_save - tprev) / h;
                            T v[N];
                            Method::interp(K, theta, h, u, unew, p, tprev, v);
                            store_u<T, N>(a, traj, cur - 1, v);
                            store_t<T>(a, traj, cur - 1, next_save);
                            ++cur;
                            next_save = (cur <= a.n_saveat) ? saveat[cur - 1] : (T)0;
                        }
                    }
                    if (!(tnew < tf)) {                      // trajectory finished
                        if (tnew > tf && !has_saveat) {      // kernels.jl:133-137 (first step overshoot)
                            const T theta = (tf - tprev) / h;
                            T v[N];
                            Method::interp(K, theta, h, u, unew, p, tprev, v);
                            store_u<T, N>(a, traj, a.n_rows - 1, v);
                            store_t<T>(a, traj, a.n_rows - 1, tf);
                        }
                        if (!has_saveat && !a.save_everystep) {   // kernels.jl:139-142
                            store_u<T, N>(a, traj, 1, unew);
                            store_t<T>(a, traj, 1, tnew);
                        }
                        bool fin = true;
                        DEGK_UNROLL for (int c = 0; c < N; ++c) fin = fin && finite_(unew[c]);
                        rc = fin ? RC_SUCCESS : RC_UNSTABLE;
                    } else {
                        Method::accepted(K);
                        DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = unew[c];
                        t = tnew;
                        h = dtnew;
                        // (a NaN dtnew is attempted like any other step, as in the reference: that attempt is accepted
                        //  -- NaN > 1 is false --, tnew becomes NaN and the `!(tnew < tf)` branch above ends the
                        //  trajectory with the NaN end point and Unstable)
                    }
                }
                if (rc == RC_DEFAULT && ++iters >= a.max_iters) rc = RC_MAXITERS;
            }
            if (rc != RC_DEFAULT) {                          // retire this trajectory
                if (has_saveat) fill_unwritten_ts<T>(a, traj, cur - 1, t0);
                if (a.retcode) a.retcode[traj] = rc;
                if (a.naccept) a.naccept[traj] = (int)nacc;
                if (a.nreject) a.nreject[traj] = (int)nrej;
                tot_acc += nacc; tot_rej += nrej;
                if (rc != RC_SUCCESS) ++tot_fail;
                have = false;
            }
        }
    }
    add_totals<T>(a, tot_acc, tot_rej, tot_fail);
}

}  // namespace degk
