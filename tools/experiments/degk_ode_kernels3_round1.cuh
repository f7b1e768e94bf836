// Round-1 third-generation adaptive kernel body (fast build of round 1: 125.8 G steps/s on C2), kept for A/B runs with
// tools/c2_probe.cu.  Not part of libdegk: the library runs degk_ode_kernels4.cuh for both fp modes.
#pragma once
#include "device/degk_ode_saves.cuh"

namespace degk {

// ------------------------------------------------------------------------------------------
template <class T, class Model, template <class, class> class MethodT, int W>
DEGK_DEV void ode_asolve3_body(const KArgs& a, unsigned char* smem_raw) {
    typedef PackOf<T, W> PO;
    typedef typename PO::type V;
    typedef MethodT<V, Model> MethodV;       // stepping (packed when W == 2)
    typedef MethodT<T, Model> MethodS;       // scalar: deferred saves
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
    const T dtmin = MethodS::dtmin();
    const T kDead = (T)-1;                   // h of a slot that is not integrating

    // controller constants in the L = log2(N * EEst^2) representation
    const double lgN = log2((double)N), lgGamma = log2(9.0 / 10.0);
    const T b1h = (T)(0.5 * 7.0 / (10.0 * MethodS::ORDER));
    const T b2h = (T)(0.5 * 2.0 / (5.0 * MethodS::ORDER));
    const T k0 = (T)((0.5 * 7.0 / (10.0 * MethodS::ORDER) - 0.5 * 2.0 / (5.0 * MethodS::ORDER)) * lgN + lgGamma);
    const T lqZero = (T)lgN;                                  // qold = 1 (reject branch ignores qold)
    const T lqInit = (T)(2.0 * log2(1.0e-4) + lgN);           // qoldinit = 1e-4
    const T exLo = (T)-2.321928094887362, exHi = (T)3.321928094887362;   // fac in [qmin, qmax] = [1/5, 10]

    // shared memory: [per-warp save queues][per-warp problem pools][saveat copy + 2 x inf]
    Rec* queue = (Rec*)smem_raw + (size_t)warp_in_block * QCAP;
    const u32 queue_saddr = opaque((u32)__cvta_generic_to_shared(queue));
    T* pool = (T*)(smem_raw + (size_t)nwarps * QCAP * sizeof(Rec)) + (size_t)warp_in_block * 32 * (N + Model::NP + 2);
    T* sv_s = (T*)(smem_raw + (size_t)nwarps * QCAP * sizeof(Rec)) + (size_t)nwarps * 32 * (N + Model::NP + 2);
    const T* sv = (const T*)a.saveat;
    const bool sv_staged = nsv <= 1024;      // must match the host's shared-memory sizing
    if (sv_staged) {
        for (int i = (int)threadIdx.x; i < nsv; i += (int)blockDim.x) sv_s[i] = sv[i];
        if (threadIdx.x < 2) sv_s[nsv + (int)threadIdx.x] = kInf;
        __syncthreads();
        sv = sv_s;
    }
    const u32 sv_saddr = opaque((u32)__cvta_generic_to_shared(sv_s) - (u32)sizeof(T));   // 1-based
    // time of the 1-based save index c; +inf past the end
    auto save_time = [&](int c) -> T {
        if (sv_staged) return lds_(sv_saddr + (u32)c * (u32)sizeof(T), (T)0);
        return c <= nsv ? sv[c - 1] : kInf;
    };
    int qcount = 0;                          // warp-uniform

    // per-thread state: W trajectories ("slots")
    V u[N], unew[N], err[N], p[NPA];
    typename MethodV::Keep K;
    T t[W], h[W], tf[W], next_save[W], next_save2[W], lq[W];   // lq: log2(N * qold^2)
    int cur[W], traj[W];
    u32 natt[W], nacc[W];
    u32 singm = 0;
    DEGK_UNROLL for (int s = 0; s < W; ++s) {
        traj[s] = -1; cur[s] = 1; natt[s] = 0; nacc[s] = 0;
        t[s] = (T)0; h[s] = kDead; tf[s] = (T)0; next_save[s] = kInf; next_save2[s] = kInf; lq[s] = lqInit;
    }
    DEGK_UNROLL for (int c = 0; c < N; ++c) { u[c] = V((T)0); unew[c] = V((T)0); err[c] = V((T)0); }
    DEGK_UNROLL for (int c = 0; c < NPA; ++c) p[c] = V((T)0);
    DEGK_UNROLL for (int j = 0; j < (int)(sizeof(K) / sizeof(V)); ++j) ((V*)&K)[j] = V((T)0);
    u32 tot_acc = 0, tot_rej = 0, tot_fail = 0;

    const bool queue_sched = (a.schedule == SCHED_QUEUE);
    bool exhausted = false;                  // warp-uniform
    bool static_done = false;
    const i64 warp_global = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int RETIRE_BATCH = a.retire_batch > 0 ? a.retire_batch : (DEGK_RETIRE_BATCH * W) / 2;

    constexpr int PW = N + Model::NP + 2;    // words per pooled problem: u0, p, t0, tf
    int pool_base = 0;                       // warp-uniform
    int pool_n = 0, pool_pos = 0;

    // start trajectory `claim` (problem data in us_/ps_/t0_/tf_) in slot s of this lane
    auto start_slot = [&](int s, int claim, const T (&us_)[N], const T (&ps_)[NPA], T t0_, T tf_, u32& freshm) {
        DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = PO::set(u[c], s, us_[c]);
        DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) p[c] = PO::set(p[c], s, ps_[c]);
        t[s] = t0_; tf[s] = tf_;
        lq[s] = lqInit;
        natt[s] = 0; nacc[s] = 0;
        cur[s] = 1;                               // kernels.jl:116-126
        if (has_saveat) {
            if (t0_ == save_time(1)) {
                cur[s] = 2;
                store_u<T, N>(a, claim, 0, us_);
                store_t<T>(a, claim, 0, t0_);
            }
        } else {
            store_t<T>(a, claim, 0, t0_);
            store_u<T, N>(a, claim, 0, us_);
            // without saveat only row 1 (endpoints) can be written later: pre-fill the rest of ts
            // with t0 now (lowerlevel_solve.jl:318 fill!)
            fill_unwritten_ts<T>(a, claim, 1, t0_);
        }
        next_save[s] = save_time(cur[s]);
        next_save2[s] = save_time(cur[s] + 1);
        if (t0_ < tf_) {
            traj[s] = claim;
            const T h0 = (T)a.dt;
            // dt0 < dtmin errors at the first attempt; non-finite time data cannot be integrated:
            // both park the slot (dead), the retire path derives the return code
            const bool valid = finite_(t0_) & finite_(tf_) & finite_(h0);
            h[s] = valid ? fmax_(h0, (T)0) : kDead;      // dt0 <= 0 fails like dt0 < dtmin
            if (h[s] >= dtmin) freshm |= (1u << s);
        } else {                                  // empty time span: nothing to integrate
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

    bool service = true;                     // warp-uniform: a slot stopped (or start of the kernel)
    u32 iter = 0;                            // warp-uniform
    bool all_done = false;
    bool started = false;                    // warp-uniform: the first service pass (initial fill) ran
    for (;;) {
        if (service) {
            u32 freshm = 0;
            for (;;) {
                // slot states: integrating (h >= dtmin) / stopped, waiting to retire / free
                u32 havem = 0, donem = 0;
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    const bool hv = h[s] >= dtmin;
                    havem |= (u32)hv << s;
                    donem |= (u32)(!hv & (traj[s] >= 0)) << s;
                }
                // several save points inside one accepted step: the queued record covers all of
                // them (process_saves loops), skip the cursor past them
                // (after the push next_save is the old next_save2 and t the end of the step, so
                //  `next_save <= t` after at least one accepted step identifies exactly those slots;
                //  before the first accepted step a save point may legitimately lie before t0 -- it
                //  is extrapolated from the first step like integrator_utils.jl:34-47 does)
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    if (next_save[s] <= t[s] && nacc[s] != 0u) {
                        while (cur[s] <= nsv && save_time(cur[s]) <= t[s]) ++cur[s];
                        next_save[s] = save_time(cur[s]);
                        next_save2[s] = save_time(cur[s] + 1);
                    }
                }
                // ---------------- retire stopped trajectories, in batches ----------------
                int ndone = 0;
                DEGK_UNROLL for (int s = 0; s < W; ++s) ndone += __popc(__ballot_sync(0xffffffffu, (donem >> s) & 1u));
                const bool none_live = __all_sync(0xffffffffu, havem == 0);
                // most service entries only find fewer stopped slots than a batch: nothing to do
                // (free slots exist only once the work queue is exhausted -- otherwise the pass
                //  that retired them refilled them -- so there is nothing to refill either)
                if (started && ndone < RETIRE_BATCH && !none_live) break;
                started = true;
                if (ndone >= RETIRE_BATCH || (ndone > 0 && none_live)) {
                    DEGK_UNROLL for (int s = 0; s < W; ++s) {
                        if ((donem >> s) & 1u) {
                            int rc = RC_SUCCESS;
                            T uf[N];
                            DEGK_UNROLL for (int c = 0; c < N; ++c) uf[c] = PO::get(u[c], s);
                            if (t[s] >= tf[s]) {
                                if (!has_saveat && !a.save_everystep) {  // kernels.jl:139-142
                                    store_u<T, N>(a, traj[s], 1, uf);
                                    store_t<T>(a, traj[s], 1, t[s]);
                                }
                                bool fin = true;
                                DEGK_UNROLL for (int c = 0; c < N; ++c) fin = fin && finite_(uf[c]);
                                if (!fin) rc = RC_UNSTABLE;
                            } else if ((singm >> s) & 1u) rc = RC_SINGULAR;
                            else if (natt[s] >= max_it) rc = RC_MAXITERS;
                            else if (h[s] >= (T)0) rc = RC_DT_LESS_THAN_MIN;
                            else rc = RC_UNSTABLE;
                            if (has_saveat) {
                                fill_unwritten_ts<T>(a, traj[s], cur[s] - 1, ((const T*)a.tspan)[(i64)traj[s] * a.tspan_stride]);
                                if (a.nsaved) a.nsaved[traj[s]] = cur[s] - 1;
                            }
                            if (a.retcode) a.retcode[traj[s]] = rc;
                            if (a.naccept) a.naccept[traj[s]] = (int)nacc[s];
                            if (a.nreject) a.nreject[traj[s]] = (int)(natt[s] - nacc[s]);
                            tot_acc += nacc[s]; tot_rej += natt[s] - nacc[s];
                            if (rc != RC_SUCCESS) ++tot_fail;
                            traj[s] = -1;
                            h[s] = kDead;
                            singm &= ~(1u << s);
                        }
                    }
                    donem = 0;
                }
                // ---------------- (re)fill free slots ----------------
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    const bool mine = !(((havem | donem) >> s) & 1u);
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
                                pool_base = (int)base; pool_n = n; pool_pos = 0;
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
                            start_slot(s, (int)claim, us_, ps_, t0_, tf_, freshm);
                        }
                    }
                }
                if (!queue_sched) { static_done = true; exhausted = true; }
                // anything integrating now?
                bool mine_live = false, mine_wait = false;
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    const bool hv = h[s] >= dtmin;
                    mine_live |= hv;
                    mine_wait |= !hv & (traj[s] >= 0);
                }
                if (__any_sync(0xffffffffu, mine_live)) break;
                if (__any_sync(0xffffffffu, mine_wait)) continue;              // retire them first
                if (exhausted && pool_pos == pool_n) { all_done = true; break; }
            }
            if (all_done) break;
            if (__any_sync(0xffffffffu, freshm != 0)) {
                MethodV::init_sel(K, u, p, PO::make(t), freshm);
            }
            service = false;
        }

        // ---------------- one attempt for every slot ----------------
        const V tv = PO::make(t), hv = PO::make(h), tfv = PO::make(tf);
        const bool solved = MethodV::template attempt<true>(K, u, p, tv, hv, unew, err);

        // ---------------- error norm and step-size factor, packed ----------------
        // tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol), sum of squares (ODE_DEFAULT_NORM)
        V accn;
        DEGK_UNROLL for (int c = 0; c < N; ++c) {
            const V sc = fma_(vmaxabs(u[c], unew[c]), V(reltol), V(abstol));
            const V v = err[c] * vrcp(sc);
            accn = (c == 0) ? v * v : fma_(v, v, accn);
        }
        const V L = vlog2(accn);                                 // log2(N * EEst^2)
        bool rej[W];
        T lqe[W];
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            rej[s] = PO::get(accn, s) > (T)N;                    // EEst > 1
            lqe[s] = rej[s] ? lqZero : lq[s];
        }
        const V ex = vclamp(fma_(V(-b1h), L, fma_(V(b2h), PO::make(lqe), V(k0))), exLo, exHi);
        const V hf = hv * vexp2(ex);                             // dt * fac
        const V rem = (tfv - tv) - hv;                           // tf - t - dt
        const V tsum = tv + hv;

        // ---------------- per-slot flags ----------------
        bool push[W], acc_[W], ok_[W], stop_[W];
        T tnew_[W], hnext_[W];
        bool any_evt = false;
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            const T rem_s = PO::get(rem, s), tsum_s = PO::get(tsum, s), hf_s = PO::get(hf, s);
            // land on tf (gpu_tsit5_perform_step.jl:155-156); a step that cannot advance t
            // (remaining span below ulp(t)) lands too -- the reference would loop forever
            const bool land = (rem_s < MethodS::land()) | ((tsum_s == t[s]) & (rem_s <= h[s]));
            const T tn = land ? tf[s] : tsum_s;
            hnext_[s] = rej[s] ? hf_s : fmin_(abs_(hf_s), abs_(rem_s));
            const bool live = h[s] >= dtmin;                     // dead slots carry h < dtmin
            const bool ok = live & solved;                       // W factorised
            const bool accept = ok & !rej[s];
            inc_if(ok, natt[s]);
            inc_if(accept, nacc[s]);
            const bool fin = accept & !(tn < tf[s]);
            const bool many = natt[s] >= max_it;
            // (a step size below dtmin needs no test here: the slot is simply not live any more
            //  in the next iteration -- `dt < dtmin && error(...)` -- and retires as DtLessThanMin)
            const bool stop = ok & (fin | many);
            const bool mult = accept & (next_save2[s] <= tn);
            push[s] = accept & (next_save[s] <= tn);
            acc_[s] = accept; ok_[s] = ok; stop_[s] = stop;
            tnew_[s] = tn;
            any_evt |= stop | mult;
            if (!MethodS::ALWAYS_SOLVED) {
                const bool sing = live & !solved;
                singm |= (u32)sing << s;
                any_evt |= sing;
                stop_[s] |= sing; ok_[s] |= sing;                // park the slot (h = -1)
            }
        }

        // ---------------- queue the deferred saves (branch-free) ----------------
        {
            int pos = qcount;
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                const u32 pm = __ballot_sync(0xffffffffu, push[s]);
                Rec r;
                r.traj = traj[s];
                r.cur = cur[s];
                r.tprev = t[s];
                r.h = h[s];
                r.tnew = tnew_[s];
                DEGK_UNROLL for (int c = 0; c < N; ++c) r.u[c] = PO::get(u[c], s);
                rec_store_if(push[s], queue_saddr + (u32)(pos + __popc(pm & lt_mask)) * (u32)sizeof(Rec), r);
                pos += __popc(pm);
                inc_if(push[s], cur[s]);
                next_save[s] = push[s] ? next_save2[s] : next_save[s];
                next_save2[s] = save_time(cur[s] + 1);
            }
            qcount = pos;
        }

        // ---------------- state update (selects only) ----------------
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            const T hn = stop_[s] ? kDead : hnext_[s];
            h[s] = ok_[s] ? hn : h[s];
            lq[s] = acc_[s] ? fmax_(PO::get(L, s), lqInit) : lq[s];
            t[s] = acc_[s] ? tnew_[s] : t[s];
        }

        // (a loop, not an `if`: up to 32 * W records can arrive in one iteration, and the queue
        //  holds 32 + 32 * W -- it must be drained below 32 before the next push)
        while (qcount >= 32) {
            __syncwarp();
            process_saves<T, Model, MethodS>(a, queue, qcount - 32, 32, sv);
            qcount -= 32;
            __syncwarp();
        }

        // ---------------- commit accepted steps ----------------
        DEGK_UNROLL for (int c = 0; c < N; ++c) assign_if(acc_, u[c], unew[c]);
        MethodV::accepted_if(K, acc_);

        service = __any_sync(0xffffffffu, any_evt) | ((++iter & 255u) == 0u);
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
