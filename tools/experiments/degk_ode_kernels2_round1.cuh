// Round-1 second-generation adaptive kernel body (strict build of round 1), kept for A/B runs with tools/c2_probe.cu.
// Not part of libdegk: the library runs degk_ode_kernels4.cuh for both fp modes.
#pragma once
#include "device/degk_ode_saves.cuh"

namespace degk {

// ------------------------------------------------------------------------------------------
template <class T, class Model, template <class, class> class MethodT, int W>
DEGK_DEV void ode_asolve2_body(const KArgs& a, unsigned char* smem_raw) {
    typedef PackOf<T, W> PO;
    typedef typename PO::type V;
    typedef MethodT<V, Model> MethodV;       // stepping (packed when W == 2)
    typedef MethodT<T, Model> MethodS;       // scalar: deferred saves, final interpolation
    typedef Ctl<T, MethodS::ORDER> C;
    constexpr int N = Model::N;
    constexpr int NPA = Model::NP > 0 ? Model::NP : 1;
    constexpr int QCAP = asolve2_qcap<T, N, W>();
    typedef SaveRec<T, N> Rec;

    const T abstol = (T)a.abstol, reltol = (T)a.reltol;
    const bool has_saveat = a.saveat != nullptr;
    const int nsv = has_saveat ? a.n_saveat : 0;
    const u32 lane = lane_id();
    const u32 lt_mask = (1u << lane) - 1u;
    const int warp_in_block = (int)(threadIdx.x >> 5);
    const int nwarps = (int)(blockDim.x >> 5);
    const u32 max_it = a.max_iters > 0x7fffffffLL ? 0x7fffffffu : (u32)a.max_iters;
    const T kInf = (T)__longlong_as_double(0x7ff0000000000000LL);   // +inf: "no further save point"

    // shared memory: [per-warp save queues][per-warp problem pools][saveat copy]
    Rec* queue = (Rec*)smem_raw + (size_t)warp_in_block * QCAP;
    T* pool = (T*)(smem_raw + (size_t)nwarps * QCAP * sizeof(Rec)) + (size_t)warp_in_block * 32 * (N + Model::NP + 2);
    T* sv_s = (T*)(smem_raw + (size_t)nwarps * QCAP * sizeof(Rec)) + (size_t)nwarps * 32 * (N + Model::NP + 2);
    const T* sv = (const T*)a.saveat;
    if (has_saveat && a.n_saveat <= 1024) {
        for (int i = (int)threadIdx.x; i < a.n_saveat; i += (int)blockDim.x) sv_s[i] = sv[i];
        __syncthreads();
        sv = sv_s;
    }
    int qcount = 0;                          // warp-uniform

    // per-thread state: W trajectories ("slots"); flags are bit masks over the slots
    V u[N], unew[N], err[N], p[NPA];
    typename MethodV::Keep K;
    T t[W], h[W], tf[W], next_save[W], next_save2[W], lq[W];   // lq: qold (strict) or log2(qold) (fast)
    int cur[W];
    i64 traj[W];
    u32 nacc[W], nrej[W];
    u32 havem = 0;                           // slots integrating a trajectory
    u32 donem = 0, failm = 0, singm = 0;     // finished and waiting for the batched retire / failed / singular W
    DEGK_UNROLL for (int s = 0; s < W; ++s) {
        traj[s] = -1; cur[s] = 0; nacc[s] = 0; nrej[s] = 0;
        t[s] = (T)0; h[s] = (T)1; tf[s] = (T)0; next_save[s] = kInf; next_save2[s] = kInf; lq[s] = (T)0;
    }
    DEGK_UNROLL for (int c = 0; c < N; ++c) { u[c] = V((T)0); unew[c] = V((T)0); err[c] = V((T)0); }
    DEGK_UNROLL for (int c = 0; c < NPA; ++c) p[c] = V((T)0);
    DEGK_UNROLL for (int j = 0; j < (int)(sizeof(K) / sizeof(V)); ++j) ((V*)&K)[j] = V((T)0);
    u32 tot_acc = 0, tot_rej = 0, tot_fail = 0;

    const bool queue_sched = (a.schedule == SCHED_QUEUE);
    bool exhausted = false;                  // warp-uniform
    bool static_done = false;
    const i64 warp_global = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr u32 ALLM = (1u << W) - 1u;
    const int RETIRE_BATCH = a.retire_batch > 0 ? a.retire_batch : (DEGK_RETIRE_BATCH * W) / 2;

    // Claimed-but-not-started trajectories are staged in a per-warp pool in shared memory: one
    // atomicAdd and one round of coalesced global loads per 32 trajectories, instead of an
    // atomic + dependent loads (~1000 cycles of exposed latency) every time a lane retires.
    constexpr int PW = N + Model::NP + 2;    // words per pooled problem: u0, p, t0, tf
    i64 pool_base = 0;                       // warp-uniform
    int pool_n = 0, pool_pos = 0;

    // start trajectory `claim` (problem data in us_/ps_/t0_/tf_) in slot s of this lane
    auto start_slot = [&](int s, i64 claim, const T (&us_)[N], const T (&ps_)[NPA], T t0_, T tf_, u32& freshm) {
        traj[s] = claim;
        DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = PO::set(u[c], s, us_[c]);
        DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) p[c] = PO::set(p[c], s, ps_[c]);
        t[s] = t0_; tf[s] = tf_;
        h[s] = (T)a.dt;
#if DEGK_STRICT
        lq[s] = C::qoldinit();
#else
        lq[s] = (T)-13.287712379549449;          // log2(qoldinit = 1e-4)
#endif
        nacc[s] = 0; nrej[s] = 0;
        cur[s] = 0;                               // kernels.jl:116-126
        if (has_saveat) {
            cur[s] = 1;
            if (t0_ == sv[0]) {
                cur[s] = 2;
                store_u<T, N>(a, claim, 0, us_);
                store_t<T>(a, claim, 0, t0_);
            }
            next_save[s] = (cur[s] <= nsv) ? sv[cur[s] - 1] : kInf;
            next_save2[s] = (cur[s] + 1 <= nsv) ? sv[cur[s]] : kInf;
        } else {
            next_save[s] = kInf;
            next_save2[s] = kInf;
            store_t<T>(a, claim, 0, t0_);
            store_u<T, N>(a, claim, 0, us_);
            // without saveat only row 1 (endpoints) or the last row (first-step overshoot) can be
            // written later: pre-fill the rest of ts with t0 now (lowerlevel_solve.jl:318 fill!)
            fill_unwritten_ts<T>(a, claim, 1, t0_);
        }
        if (t0_ < tf_) {
            havem |= (1u << s);
            freshm |= (1u << s);
        } else {                                  // empty time span
            if (!has_saveat && !a.save_everystep) { store_u<T, N>(a, claim, 1, us_); store_t<T>(a, claim, 1, t0_); }
            if (a.retcode) a.retcode[claim] = RC_SUCCESS;
            if (a.naccept) a.naccept[claim] = 0;
            if (a.nreject) a.nreject[claim] = 0;
            if (has_saveat) {
                fill_unwritten_ts<T>(a, claim, cur[s] - 1, t0_);
                if (a.nsaved) a.nsaved[claim] = cur[s] - 1;
            }
        }
    };

    for (;;) {
        // ---------------- retire finished / failed trajectories, in batches ----------------
        // Retiring (and then refilling) runs with one or two active lanes, so it is done only when
        // RETIRE_BATCH slots of the warp are waiting (or nothing is left to integrate): the
        // ~300 divergent instructions are paid once per batch instead of once per trajectory.
        {
            int ndone = 0;
            DEGK_UNROLL for (int s = 0; s < W; ++s) ndone += __popc(__ballot_sync(0xffffffffu, (donem >> s) & 1u));
            if (ndone >= RETIRE_BATCH || (ndone > 0 && __all_sync(0xffffffffu, havem == 0))) {
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    if ((donem >> s) & 1u) {
                        int rc = RC_SUCCESS;
                        T uf[N];
                        DEGK_UNROLL for (int c = 0; c < N; ++c) uf[c] = PO::get(u[c], s);
                        if ((failm >> s) & 1u) {
                            if ((singm >> s) & 1u) rc = RC_SINGULAR;
                            else if (h[s] < MethodS::dtmin()) rc = RC_DT_LESS_THAN_MIN;
                            else if (nacc[s] + nrej[s] + 1u >= max_it) rc = RC_MAXITERS;
                            else rc = RC_UNSTABLE;
                        } else {
                            if (!has_saveat && !a.save_everystep) {  // kernels.jl:139-142
                                store_u<T, N>(a, traj[s], 1, uf);
                                store_t<T>(a, traj[s], 1, t[s]);
                            }
                            bool fin = true;
                            DEGK_UNROLL for (int c = 0; c < N; ++c) fin = fin && finite_(uf[c]);
                            if (!fin) rc = RC_UNSTABLE;
                        }
                        if (has_saveat) {
                            fill_unwritten_ts<T>(a, traj[s], cur[s] - 1, ((const T*)a.tspan)[traj[s] * a.tspan_stride]);
                            if (a.nsaved) a.nsaved[traj[s]] = cur[s] - 1;
                        }
                        if (a.retcode) a.retcode[traj[s]] = rc;
                        if (a.naccept) a.naccept[traj[s]] = (int)nacc[s];
                        if (a.nreject) a.nreject[traj[s]] = (int)nrej[s];
                        tot_acc += nacc[s]; tot_rej += nrej[s];
                        if (rc != RC_SUCCESS) ++tot_fail;
                    }
                }
                donem = 0; failm = 0; singm = 0;
            }
        }

        // ---------------- (re)fill free slots ----------------
        if (__any_sync(0xffffffffu, (havem | donem) != ALLM)) {
            u32 freshm = 0;
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                const bool mine = !((havem | donem) & (1u << s));
                const u32 need = __ballot_sync(0xffffffffu, mine);
                if (need == 0) continue;
                if (queue_sched) {
                    const int cnt = __popc(need);
                    const int rank = __popc(need & lt_mask);
                    int served = 0;
                    while (served < cnt) {
                        if (pool_pos == pool_n) {                 // pool empty: claim the next 32
                            if (exhausted) break;
                            i64 base = 0;
                            if (lane == 0) base = (i64)atomicAdd(a.work_counter, (u64)32);
                            base = __shfl_sync(0xffffffffu, base, 0);
                            i64 left = a.n_traj - base;
                            int n = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
                            if (base + 32 >= a.n_traj) exhausted = true;
                            if ((int)lane < n) {
                                T us_[N], ps_[NPA], t0_, tf_;
                                load_problem<T, Model>(a, base + lane, us_, ps_, t0_, tf_);
                                T* e = pool + lane * PW;
                                DEGK_UNROLL for (int c = 0; c < N; ++c) e[c] = us_[c];
                                DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) e[N + c] = ps_[c];
                                e[N + Model::NP] = t0_; e[N + Model::NP + 1] = tf_;
                            }
                            __syncwarp();
                            pool_base = base; pool_n = n; pool_pos = 0;
                            if (n == 0) break;
                        }
                        const int avail = pool_n - pool_pos;
                        const int take = avail < cnt - served ? avail : cnt - served;
                        if (mine && rank >= served && rank < served + take) {
                            const int ei = pool_pos + rank - served;
                            const T* e = pool + ei * PW;
                            T us_[N], ps_[NPA];
                            DEGK_UNROLL for (int c = 0; c < N; ++c) us_[c] = e[c];
                            DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) ps_[c] = e[N + c];
                            start_slot(s, pool_base + ei, us_, ps_, e[N + Model::NP], e[N + Model::NP + 1], freshm);
                        }
                        pool_pos += take;
                        served += take;
                    }
                    __syncwarp();
                } else if (!static_done) {
                    const i64 claim = (warp_global * W + s) * 32 + lane;
                    if (mine && claim < a.n_traj) {
                        T us_[N], ps_[NPA], t0_, tf_;
                        load_problem<T, Model>(a, claim, us_, ps_, t0_, tf_);
                        start_slot(s, claim, us_, ps_, t0_, tf_, freshm);
                    }
                }
            }
            if (!queue_sched) { static_done = true; exhausted = true; }
            if (__all_sync(0xffffffffu, havem == 0)) {
                if (__any_sync(0xffffffffu, donem != 0)) continue;          // retire them first
                if (exhausted && pool_pos == pool_n) break;
                continue;
            }
            if (__any_sync(0xffffffffu, freshm != 0)) {
                T tt0[W];
                DEGK_UNROLL for (int s = 0; s < W; ++s) tt0[s] = t[s];
                MethodV::init_sel(K, u, p, PO::make(tt0), freshm);
            }
        }

        // ---------------- one attempt for every slot ----------------
        T tt[W], hh[W];
        DEGK_UNROLL for (int s = 0; s < W; ++s) { tt[s] = t[s]; hh[s] = h[s]; }
        const bool solved = MethodV::template attempt<true>(K, u, p, PO::make(tt), PO::make(hh), unew, err);

        // ---------------- step-size control, per slot ----------------
        u32 accm = 0, pushm = 0;                         // accepted / save crossing
        T tnew_[W];
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            // tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol); ODE_DEFAULT_NORM
            T accn = (T)0;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                const T uo = PO::get(u[c], s), un = PO::get(unew[c], s), e = PO::get(err[c], s);
#if DEGK_STRICT
                const T sc = abstol + jl_max(abs_(uo), abs_(un)) * reltol;
                const T v = e / sc;
#else
                const T sc = fma_(fmax_(abs_(uo), abs_(un)), reltol, abstol);
                const T v = e * rcp_(sc);
#endif
                const T sq = v * v;
                accn = (c == 0) ? sq : accn + sq;
            }
            const T rem = tf[s] - t[s] - h[s];                  // tf - t - dt
            // a step that cannot advance t (remaining span below ulp(t)) lands on tf: the reference
            // would loop forever here (deviations)
            const T tsum = t[s] + h[s];
            const T tn = ((rem < MethodS::land()) | ((tsum == t[s]) & (rem <= h[s]))) ? tf[s] : tsum;
            T h_next, lq_next;
            bool reject;
#if DEGK_STRICT
            const T EEst = sqrt_(mean_<T, N>(accn));
            T q, q11 = (T)0;
            if (EEst == (T)0) {
                q = (T)1 / C::qmax();
            } else {
                q11 = pow_(EEst, C::beta1());
                q = q11 / pow_(lq[s], C::beta2());
            }
            reject = EEst > (T)1;
            if (reject) {
                h_next = h[s] / jl_min((T)1 / C::qmin(), q11 / C::gamma());
                lq_next = lq[s];
            } else {
                q = jl_max((T)1 / C::qmax(), jl_min((T)1 / C::qmin(), q / C::gamma()));
                lq_next = jl_max(EEst, C::qoldinit());
                h_next = jl_min(abs_(h[s] / q), abs_(rem));
            }
#else
            const T m = mean_<T, N>(accn);                      // EEst^2
            const T lE = (T)0.5 * log2_(m);                     // log2(EEst); -inf when EEst == 0
            reject = m > (T)1;
            // accept: q/gamma = 2^(b1*lE - b2*lq - log2(gamma)) clamped to [1/qmax, 1/qmin], dtnew = dt/q
            // reject: dt / min(1/qmin, q11/gamma) = dt * max(qmin, 2^(log2(gamma) - b1*lE))
            // -> one EX2 on the selected exponent
            T e2 = fma_(C::beta1(), lE, fma_(-C::beta2(), lq[s], (T)0.15200309344504995));
            e2 = fmax_((T)-3.321928094887362, fmin_((T)2.321928094887362, e2));
            const T ex = reject ? fma_(-C::beta1(), lE, (T)-0.15200309344504995) : -e2;
            const T fac = exp2_(ex);
            const T h_acc = fmin_(abs_(h[s] * fac), abs_(rem));
            const T h_rej = h[s] * fmax_(C::qmin(), fac);
            h_next = reject ? h_rej : h_acc;
            lq_next = reject ? lq[s] : fmax_(lE, (T)-13.287712379549449);
#endif
            // flags as 0/1 integers combined with bitwise ops: no short-circuit branches
            const u32 live = (havem >> s) & 1u;
            const u32 too_small = h[s] < MethodS::dtmin();      // `dt < dtmin && error(...)`
            const u32 ok = live & (u32)solved & (too_small ^ 1u);
            const u32 accept = ok & ((u32)reject ^ 1u);
            const u32 finished = accept & (u32)!(tn < tf[s]);
            // non-finite time or step size: the reference would spin on NaNs; report Unstable
            const u32 bad_num = accept & (finished ^ 1u) & (u32)!(abs_(tn + h_next) < (T)(sizeof(T) == 4 ? 3.0e38 : 1.0e300));
            const u32 too_many = ok & (u32)(nacc[s] + nrej[s] + 1u >= max_it);
            const u32 fail = live & (((u32)solved ^ 1u) | too_small | bad_num | too_many);
            tnew_[s] = tn;
            nacc[s] += accept;
            nrej[s] += ok & (u32)reject;
            pushm |= (accept & (u32)(next_save[s] <= tn)) << s;   // next_save = +inf past the last point
            accm |= accept << s;
            const u32 stop = finished | fail;                     // leaves the integration loop
            havem &= ~(stop << s);
            donem |= stop << s;
            failm |= fail << s;
            singm |= (live & ((u32)solved ^ 1u)) << s;
            h[s] = ok ? h_next : h[s];
            lq[s] = ok ? lq_next : lq[s];
            t[s] = accept ? tn : t[s];
        }

        // ---------------- queue the deferred saves ----------------
        if (__any_sync(0xffffffffu, pushm != 0)) {
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                const bool ps = (pushm >> s) & 1u;
                const u32 pm = __ballot_sync(0xffffffffu, ps);
                if (ps) {
                    Rec r;
                    r.traj = (int)traj[s];
                    r.cur = cur[s];
                    r.tprev = tt[s];
                    r.h = hh[s];
                    r.tnew = tnew_[s];
                    DEGK_UNROLL for (int c = 0; c < N; ++c) r.u[c] = PO::get(u[c], s);
                    rec_copy(queue + qcount + __popc(pm & lt_mask), &r);
                    // advance to the next save point; its time was prefetched into next_save2 at
                    // the previous crossing, so the shared-memory load issued here is not waited on
                    ++cur[s];
                    next_save[s] = next_save2[s];
                    while (cur[s] <= nsv && next_save[s] <= tnew_[s]) {   // several save points in one step
                        ++cur[s];
                        next_save[s] = (cur[s] <= nsv) ? sv[cur[s] - 1] : kInf;
                    }
                    next_save2[s] = (cur[s] + 1 <= nsv) ? sv[cur[s]] : kInf;
                }
                qcount += __popc(pm);
            }
            __syncwarp();
            while (qcount >= 32) {
                process_saves<T, Model, MethodS>(a, queue, qcount - 32, 32, sv);
                qcount -= 32;
                __syncwarp();
            }
        }

        // (kernels.jl:133-137 interpolates back when integ.t > tf.  In the adaptive path that cannot
        //  happen: a step with t + dt >= tf always has (tf - t - dt) < 1e-14 and lands on tf.)

        // ---------------- commit accepted steps ----------------
        DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = blendm(accm, unew[c], u[c]);
        MethodV::accepted_sel(K, accm);
    }
    // flush the remaining deferred saves
    __syncwarp();
    while (qcount > 0) {
        const int n = qcount < 32 ? qcount : 32;
        process_saves<T, Model, MethodS>(a, queue, qcount - n, n, sv);
        qcount -= n;
        __syncwarp();
    }
    add_totals<T>(a, tot_acc, tot_rej, tot_fail);
}

}  // namespace degk
