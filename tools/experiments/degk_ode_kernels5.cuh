// degk_ode_kernels5.cuh -- the adaptive ensemble kernel, two phase-shifted streams per thread.
//
// Same job, same building blocks and same arithmetic policies as degk_ode_kernels4.cuh (reference
// kernels.jl:74-152 + the adaptive step! of each solver; persistent warps, problem pool, deferred batched saves,
// strict / fast StepMath).  What changes is the shape of the loop, and it comes from the ncu profile of the
// fourth generation (profiles/r2_c2_v4_fast.txt):  an SM sub-partition of sm_100 has two half-rate pipes that
// matter here -- the FMA pipe (packed FFMA2/FMUL2/FADD2: 2 cycles each) and the ALU pipe (compares, selects,
// min/max, integer and predicate work: 2 cycles each) -- and one issue slot per cycle.  One attempt of the
// one-stream kernel is ~155 FMA-pipe instructions (the stages) followed by ~140 ALU-pipe instructions (error norm
// tail, step-size control, flags, save queue, commits).  Each pipe alone would be busy ~60 % of the time, but the
// two phases of a warp cannot overlap (the control needs the stages, the next stages need the control), and with
// four resident warps per sub-partition the phases of different warps overlap only by chance: measured, the
// kernel needs 588 cycles per pass where either pipe needs ~350.
//
// Here every thread carries TWO independent streams (each W trajectories: a packed pair in the fast build), half a
// pass apart: while stream A runs its stages, stream B runs its control, then they swap.  The two halves are
// independent straight-line code in one basic block, and the generated steppers call a hook after every stage
// (gen_erk_*.cuh::attempt<..., Hook>) through which the other stream's control is spliced in piece by piece, so
// the instruction stream of every warp alternates between the two pipes by construction.
//
// The service path (retire / refill), the save queue and the replay are the fourth generation's, applied to the
// stream that has just finished its control (its state is committed; the other stream has an attempt in flight
// and is left alone).  The strict build instantiates the same body with one scalar trajectory per stream and is
// compared bit for bit with the CPU oracle, like generation four.
#pragma once
#include "device/degk_ode_kernels4.cuh"

#ifndef DEGK5_CHUNKS
#define DEGK5_CHUNKS 6      // pieces the control of one stream is cut into (spliced between the other stream's stages)
#endif

namespace degk {

template <class M, class = void> struct has_hooks_of { static constexpr bool value = false; };
template <class M> struct has_hooks_of<M, typename replay_void_<decltype(M::NHOOK)>::type> { static constexpr bool value = true; };

template <class T, class Model, template <class, class> class MethodT, int W>
struct Asolve5 {
    static constexpr bool FAST = !DEGK_STRICT;
    static constexpr bool PACKED = slots4_packed<T, W>();
    typedef Slots4<T, W, PACKED> PO;
    typedef typename PO::type V;
    typedef MethodT<V, Model> MethodV;       // stepping (packed when PACKED)
    typedef MethodT<T, Model> MethodS;       // scalar: deferred saves, constants
    typedef StepMath<T, MethodS::ORDER, Model::N, false> SM;
    static constexpr int N = Model::N;
    static constexpr int NPA = Model::NP > 0 ? Model::NP : 1;
    static constexpr int QCAP = asolve4_qcap<T, N, W>();
    static constexpr int QBATCH = 32;
    static constexpr bool HK = use_hk<MethodV>();
    static constexpr int PW = asolve4_pool_words<T, N, Model::NP>();
    typedef SaveRec<T, N> Rec;

    // ---- one stream: W trajectories ("slots") of this thread ----
    struct Stream {
        V u[N], unew[N], err[N], p[NPA];
        typename MethodV::Keep K;
        T t[W], h[W], tf[W], next_save[W], next_save2[W], lq[W];
        int cur[W], traj[W];
        u32 natt[W], nacc[W];
        u32 singm;
        bool solved;                         // outcome of the attempt in flight (linear solve of the stiff steppers)
        bool idle;                           // warp-uniform: nothing left to integrate in this stream
        bool svc;                            // warp-uniform: look at this stream's stopped slots after its next control
    };
    // ---- what the control of one pass hands from piece to piece ----
    struct Ctl {
        V accn, L, ex, hf, rem, tsum;        // packed quantities of the fast build
        bool rej[W], push[W], acc_[W], stop_[W];
        T hf_[W], lqa_[W], rem_[W], tsum_[W], tnew_[W], hnext_[W];
        bool any_evt;
    };
    // ---- per-warp context and launch constants ----
    struct Ctx {
        T abstol, reltol, dtmin, kInf, kDead, lqInit;
        T b1h, b2h, k0, lqZero, exLo, exHi;
        bool has_saveat, queue_sched, exhausted, static_done;
        int nsv, RETIRE_BATCH, qcount, pool_base, pool_n, pool_pos;
        u32 lane, lt_mask, max_it, queue_saddr, sv_saddr, iter;
        i64 warp_global;
        Rec* queue; T* pool; T* sv_s;
        u32 tot_acc, tot_rej, tot_fail;
    };

    static DEGK_DEV T save_time(const Ctx& c, int k) { return lds_(c.sv_saddr + (u32)k * (u32)sizeof(T), (T)0); }   // k <= nsv + 2

    // Load up to 32 problems [base, base + n) into the pool, one per lane, and do everything that does not depend
    // on the integration with the full warp (see degk_ode_kernels4.cuh::load_pool).
    static DEGK_DEV void load_pool(const KArgs& a, Ctx& c, i64 base, int n) {
        if ((int)c.lane < n) {
            const i64 claim = base + c.lane;
            T us_[N], ps_[NPA], t0_, tf_;
            load_problem<T, Model>(a, claim, us_, ps_, t0_, tf_);
            int c1 = 1;
            if (c.has_saveat) {
                if (t0_ == save_time(c, 1)) { c1 = 2; store_u<T, N>(a, claim, 0, us_); store_t<T>(a, claim, 0, t0_); }
                if (a.ts != nullptr) for (i64 k = c1 - 1; k < a.n_rows; ++k) store_t<T>(a, claim, k, t0_);
            } else {
                store_u<T, N>(a, claim, 0, us_);
                if (a.ts != nullptr) for (i64 k = 0; k < a.n_rows; ++k) store_t<T>(a, claim, k, t0_);
            }
            if (!(t0_ < tf_)) {                          // empty time span: nothing to integrate
                if (!c.has_saveat && !a.save_everystep) { store_u<T, N>(a, claim, 1, us_); store_t<T>(a, claim, 1, t0_); }
                if (a.retcode) a.retcode[claim] = RC_SUCCESS;
                if (a.naccept) a.naccept[claim] = 0;
                if (a.nreject) a.nreject[claim] = 0;
                if (c.has_saveat && a.nsaved) a.nsaved[claim] = c1 - 1;
                c1 = 0;
            }
            T* e = c.pool + c.lane * PW;
            DEGK_UNROLL for (int k = 0; k < N; ++k) e[k] = us_[k];
            DEGK_UNROLL for (int k = 0; k < Model::NP; ++k) e[N + k] = ps_[k];
            e[N + Model::NP] = t0_; e[N + Model::NP + 1] = tf_;
            ((int*)e)[(N + Model::NP + 2) * (int)(sizeof(T) / sizeof(int))] = c1;
        }
        __syncwarp();
    }

    // start the pooled trajectory `ei` in slot s of stream S of this lane
    static DEGK_DEV void start_slot(const KArgs& a, Ctx& c, Stream& S, int s, int ei, u32& freshm) {
        const T* e = c.pool + ei * PW;
        const int c1 = ((const int*)e)[(N + Model::NP + 2) * (int)(sizeof(T) / sizeof(int))];
        if (c1 == 0) return;                              // finished at load time; the slot stays free
        DEGK_UNROLL for (int k = 0; k < N; ++k) S.u[k] = PO::set(S.u[k], s, e[k]);
        DEGK_UNROLL for (int k = 0; k < Model::NP; ++k) S.p[k] = PO::set(S.p[k], s, e[N + k]);
        const T t0_ = e[N + Model::NP], tf_ = e[N + Model::NP + 1];
        S.t[s] = t0_; S.tf[s] = tf_;
        S.lq[s] = c.lqInit;
        S.natt[s] = 0; S.nacc[s] = 0;
        S.cur[s] = c1;
        S.next_save[s] = save_time(c, c1);
        S.next_save2[s] = save_time(c, c1 + 1);
        S.traj[s] = c.pool_base + ei;
        const T h0 = (T)a.dt;
        // dt0 < dtmin errors at the first attempt; non-finite time data cannot be integrated:
        // both park the slot (dead), the retire path derives the return code
        const bool valid = finite_(t0_) & finite_(tf_) & finite_(h0);
        S.h[s] = valid ? fmax_(h0, (T)0) : c.kDead;       // dt0 <= 0 fails like dt0 < dtmin
        if (S.h[s] >= c.dtmin) freshm |= (1u << s);
    }

    // ---- service path of one stream: cursor fix-ups, batched retire, refill (state of S is committed) ----
    static DEGK_DEV void service(const KArgs& a, Ctx& c, Stream& S, bool& started) {
        u32 freshm = 0;
        for (;;) {
            // slot states: integrating (h >= dtmin) / stopped, waiting to retire / free
            u32 havem = 0, donem = 0;
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                const bool hv = S.h[s] >= c.dtmin;
                havem |= (u32)hv << s;
                donem |= (u32)(!hv & (S.traj[s] >= 0)) << s;
            }
            // several save points inside one accepted step: the queued record covers all of them (the replay
            // loops), skip the cursor past them (see degk_ode_kernels4.cuh for why `next_save <= t` identifies them)
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                if (S.next_save[s] <= S.t[s] && S.nacc[s] != 0u) {
                    while (S.cur[s] <= c.nsv && save_time(c, S.cur[s]) <= S.t[s]) ++S.cur[s];
                    S.next_save[s] = save_time(c, S.cur[s]);
                    S.next_save2[s] = save_time(c, S.cur[s] + 1);
                }
            }
            // ---------------- retire stopped trajectories, in batches ----------------
            int ndone = 0;
            DEGK_UNROLL for (int s = 0; s < W; ++s) ndone += __popc(__ballot_sync(0xffffffffu, (donem >> s) & 1u));
            const bool none_live = __all_sync(0xffffffffu, havem == 0);
            // most entries only find fewer stopped slots than a batch: nothing to do (free slots exist only once the
            // work queue is exhausted -- otherwise the pass that retired them refilled them)
            if (started && ndone < c.RETIRE_BATCH && !none_live) break;
            started = true;
            if (ndone >= c.RETIRE_BATCH || (ndone > 0 && none_live)) {
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    if ((donem >> s) & 1u) {
                        int rc = RC_SUCCESS;
                        const u32 natt_ = S.natt[s];
                        if (S.t[s] >= S.tf[s]) {
                            T uf[N];
                            DEGK_UNROLL for (int k = 0; k < N; ++k) uf[k] = PO::get(S.u[k], s);
                            if (!c.has_saveat && !a.save_everystep) {  // kernels.jl:139-142
                                store_u<T, N>(a, S.traj[s], 1, uf);
                                store_t<T>(a, S.traj[s], 1, S.t[s]);
                            }
                            bool fin = true;
                            DEGK_UNROLL for (int k = 0; k < N; ++k) fin = fin && finite_(uf[k]);
                            if (!fin) rc = RC_UNSTABLE;
                        } else if ((S.singm >> s) & 1u) rc = RC_SINGULAR;
                        else if (natt_ >= c.max_it) rc = RC_MAXITERS;
                        else if (S.h[s] >= (T)0) rc = RC_DT_LESS_THAN_MIN;
                        else rc = RC_UNSTABLE;
                        if (c.has_saveat && a.nsaved) a.nsaved[S.traj[s]] = S.cur[s] - 1;
                        if (a.retcode) a.retcode[S.traj[s]] = rc;
                        if (a.naccept) a.naccept[S.traj[s]] = (int)S.nacc[s];
                        if (a.nreject) a.nreject[S.traj[s]] = (int)(natt_ - S.nacc[s]);
                        c.tot_acc += S.nacc[s]; c.tot_rej += natt_ - S.nacc[s];
                        if (rc != RC_SUCCESS) ++c.tot_fail;
                        S.traj[s] = -1;
                        S.h[s] = c.kDead;
                        S.singm &= ~(1u << s);
                    }
                }
                donem = 0;
            }
            // ---------------- (re)fill free slots ----------------
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                const bool mine = !(((havem | donem) >> s) & 1u);
                const u32 need = __ballot_sync(0xffffffffu, mine);
                if (need == 0) continue;
                if (c.queue_sched) {
                    const int cnt = __popc(need);
                    const int rank = __popc(need & c.lt_mask);
                    int served = 0;
                    while (served < cnt) {
                        if (c.pool_pos == c.pool_n) {                 // pool empty: claim the next 32
                            if (c.exhausted) break;
                            i64 base = 0;
                            if (c.lane == 0) base = (i64)atomicAdd(a.work_counter, (u64)32);
                            base = __shfl_sync(0xffffffffu, base, 0);
                            i64 left = a.n_traj - base;
                            int n = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
                            if (base + 32 >= a.n_traj) c.exhausted = true;
                            load_pool(a, c, base, n);
                            c.pool_base = (int)base; c.pool_n = n; c.pool_pos = 0;
                            if (n == 0) break;
                        }
                        const int avail = c.pool_n - c.pool_pos;
                        const int take = avail < cnt - served ? avail : cnt - served;
                        if (mine && rank >= served && rank < served + take) start_slot(a, c, S, s, c.pool_pos + rank - served, freshm);
                        c.pool_pos += take;
                        served += take;
                    }
                    __syncwarp();
                }
            }
            // anything integrating now?
            bool mine_live = false, mine_wait = false;
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                const bool hv = S.h[s] >= c.dtmin;
                mine_live |= hv;
                mine_wait |= !hv & (S.traj[s] >= 0);
            }
            if (__any_sync(0xffffffffu, mine_live)) break;
            if (__any_sync(0xffffffffu, mine_wait)) continue;              // retire them first
            if (c.exhausted && c.pool_pos == c.pool_n) { S.idle = true; break; }
        }
        if (__any_sync(0xffffffffu, freshm != 0)) MethodV::init_sel(S.K, S.u, S.p, PO::make(S.t), freshm);
    }

    // static schedule: stream q, slot s of warp w owns trajectories ((w * 2 + q) * W + s) * 32 + lane
    static DEGK_DEV void fill_static(const KArgs& a, Ctx& c, Stream& S, int q) {
        u32 freshm = 0;
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            const i64 base = ((c.warp_global * 2 + q) * W + s) * 32;
            const i64 left = a.n_traj - base;
            const int n = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
            load_pool(a, c, base, n);
            c.pool_base = (int)base;
            if ((int)c.lane < n) start_slot(a, c, S, s, (int)c.lane, freshm);
            __syncwarp();
        }
        c.pool_n = c.pool_pos = 0;
        if (__any_sync(0xffffffffu, freshm != 0)) MethodV::init_sel(S.K, S.u, S.p, PO::make(S.t), freshm);
    }

    // ---- the control of one pass, in DEGK5_CHUNKS pieces ----
    template <int PIECE>
    static DEGK_DEV void control(Ctx& c, Stream& S, Ctl& x) {
        if constexpr (PIECE == 0) {
            // tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol), sum of squares (ODE_DEFAULT_NORM)
            if constexpr (FAST) {
                DEGK_UNROLL for (int k = 0; k < N; ++k) {
                    const V sc = fma_(vmaxabs(S.u[k], S.unew[k]), V(c.reltol), V(c.abstol));
                    const V v = S.err[k] * vrcp(sc);
                    x.accn = (k == 0) ? v * v : fma_(v, v, x.accn);
                }
                x.L = vlog2(x.accn);                                     // log2(N * EEst^2)
            } else {
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    T accn = (T)0;
                    DEGK_UNROLL for (int k = 0; k < N; ++k) {
                        const T sq = SM::scaled_sq(PO::get(S.u[k], s), PO::get(S.unew[k], s), PO::get(S.err[k], s), c.abstol, c.reltol);
                        accn = (k == 0) ? sq : accn + sq;
                    }
                    x.hf_[s] = accn;                                     // handed to piece 1
                }
            }
        } else if constexpr (PIECE == 1) {
            if constexpr (FAST) {
                T lqe[W];
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    x.rej[s] = PO::get(x.accn, s) > (T)N;                // EEst > 1
                    lqe[s] = x.rej[s] ? c.lqZero : S.lq[s];
                }
                x.ex = vclamp(fma_(V(-c.b1h), x.L, fma_(V(c.b2h), PO::make(lqe), V(c.k0))), c.exLo, c.exHi);
            } else {
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    const T accn = x.hf_[s];
                    SM::control(accn, S.lq[s], S.h[s], x.rej[s], x.hf_[s], x.lqa_[s]);
                }
            }
        } else if constexpr (PIECE == 2) {
            if constexpr (FAST) {
                const V hv = PO::make(S.h), tv = PO::make(S.t), tfv = PO::make(S.tf);
                x.hf = hv * vexp2(x.ex);                                 // dt * fac
                x.rem = (tfv - tv) - hv;                                 // tf - t - dt
                x.tsum = tv + hv;
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    x.hf_[s] = PO::get(x.hf, s); x.rem_[s] = PO::get(x.rem, s); x.tsum_[s] = PO::get(x.tsum, s);
                    x.lqa_[s] = fmax_(PO::get(x.L, s), c.lqInit);
                }
            } else {
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    x.rem_[s] = S.tf[s] - S.t[s] - S.h[s];
                    x.tsum_[s] = S.t[s] + S.h[s];
                }
            }
        } else if constexpr (PIECE == 3) {
            x.any_evt = false;
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                // land on tf (gpu_tsit5_perform_step.jl:155-156); a step that cannot advance t
                // (remaining span below ulp(t)) lands too -- the reference would loop forever
                const bool land = (x.rem_[s] < MethodS::land()) | ((x.tsum_[s] == S.t[s]) & (x.rem_[s] <= S.h[s]));
                const T tn = land ? S.tf[s] : x.tsum_[s];
                if constexpr (FAST) x.hnext_[s] = x.rej[s] ? x.hf_[s] : fmin_(abs_(x.hf_[s]), abs_(x.rem_[s]));
                else x.hnext_[s] = x.rej[s] ? x.hf_[s] : SM::next_h_accept(x.hf_[s], x.rem_[s]);
                const bool live = S.h[s] >= c.dtmin;                 // dead slots carry h < dtmin
                const bool ok = live & S.solved;                     // W factorised
                const bool accept = ok & !x.rej[s];
                inc_if(ok, S.natt[s]);
                inc_if(accept, S.nacc[s]);
                const bool fin = accept & !(tn < S.tf[s]);
                const bool many = S.natt[s] >= c.max_it;
                const bool mult = accept & (S.next_save2[s] <= tn);
                x.push[s] = accept & (S.next_save[s] <= tn);
                x.acc_[s] = accept; x.stop_[s] = ok & (fin | many);
                x.tnew_[s] = tn;
                // a stopped slot idles until the service path next looks (every DEGK4_SERVICE_PERIOD passes); only a
                // step across several save points needs it at once (the cursor has to be moved on)
                x.any_evt |= mult;
                bool okh = ok;
                if (!MethodS::ALWAYS_SOLVED) {
                    const bool sing = live & !S.solved;
                    S.singm |= (u32)sing << s;
                    x.stop_[s] |= sing; okh |= sing;                 // park the slot (h = -1)
                }
                const T hn = x.stop_[s] ? c.kDead : x.hnext_[s];
                x.hnext_[s] = okh ? hn : S.h[s];
            }
        } else if constexpr (PIECE == 4) {
            // queue the deferred saves (branch-free)
            int pos = c.qcount;
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                const u32 pm = __ballot_sync(0xffffffffu, x.push[s]);
                Rec r;
                r.traj = S.traj[s];
                r.cur = S.cur[s];
                r.tprev = S.t[s];
                r.h = S.h[s];
                r.tnew = x.tnew_[s];
                DEGK_UNROLL for (int k = 0; k < N; ++k) r.u[k] = PO::get(S.u[k], s);
                rec_store_if(x.push[s], c.queue_saddr + (u32)(pos + __popc(pm & c.lt_mask)) * (u32)sizeof(Rec), r);
                pos += __popc(pm);
                inc_if(x.push[s], S.cur[s]);
                S.next_save[s] = x.push[s] ? S.next_save2[s] : S.next_save[s];
                S.next_save2[s] = save_time(c, S.cur[s] + 1);
            }
            c.qcount = pos;
        } else if constexpr (PIECE == 5) {
            // state update and commit of the accepted steps
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                S.h[s] = x.hnext_[s];
                S.lq[s] = x.acc_[s] ? x.lqa_[s] : S.lq[s];
                S.t[s] = x.acc_[s] ? x.tnew_[s] : S.t[s];
            }
            DEGK_UNROLL for (int k = 0; k < N; ++k) assign_if(x.acc_, S.u[k], S.unew[k]);
            MethodV::accepted_if(S.K, x.acc_);
        }
    }

    // the other stream's control, spliced between the stages: piece q runs at hook J when J == q * NH / DEGK5_CHUNKS
    struct Splice {
        Ctx& c; Stream& S; Ctl& x;
        DEGK_DEV Splice(Ctx& c_, Stream& S_, Ctl& x_) : c(c_), S(S_), x(x_) {}
        template <int J, int NH> DEGK_DEV void at() {
            piece_at<J, NH, 0>();
        }
        template <int J, int NH, int Q> DEGK_DEV void piece_at() {
            if constexpr (Q < DEGK5_CHUNKS) {
                constexpr int NHE = NH < DEGK5_CHUNKS ? DEGK5_CHUNKS : NH;     // fewer hooks than pieces: several pieces per hook
                if constexpr ((Q * NHE) / DEGK5_CHUNKS == J * (NHE / NH)) control<Q>(c, S, x);
                piece_at<J, NH, Q + 1>();
            }
        }
        DEGK_DEV void all() {
            control<0>(c, S, x); control<1>(c, S, x); control<2>(c, S, x); control<3>(c, S, x); control<4>(c, S, x); control<5>(c, S, x);
        }
    };

    // stages of stream A with the control of stream B spliced in
    static DEGK_DEV void half_pass(Ctx& c, Stream& A, Stream& B, Ctl& xb) {
        const V tv = PO::make(A.t), hv = PO::make(A.h);
        Splice sp(c, B, xb);
        if constexpr (has_hooks_of<MethodV>::value) {
            if constexpr (HK) A.solved = MethodV::template attempt_hk<true>(A.K, A.u, A.p, tv, hv, A.unew, A.err, sp);
            else A.solved = MethodV::template attempt<true>(A.K, A.u, A.p, tv, hv, A.unew, A.err, sp);
        } else {
            sp.all();
            A.solved = MethodV::template attempt<true>(A.K, A.u, A.p, tv, hv, A.unew, A.err);
        }
    }

    static DEGK_DEV void init_stream(const Ctx& c, Stream& S) {
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            S.traj[s] = -1; S.cur[s] = 1; S.natt[s] = 0; S.nacc[s] = 0;
            S.t[s] = (T)0; S.h[s] = c.kDead; S.tf[s] = (T)0; S.next_save[s] = c.kInf; S.next_save2[s] = c.kInf; S.lq[s] = c.lqInit;
        }
        DEGK_UNROLL for (int k = 0; k < N; ++k) { S.u[k] = V((T)0); S.unew[k] = V((T)0); S.err[k] = V((T)0); }
        DEGK_UNROLL for (int k = 0; k < NPA; ++k) S.p[k] = V((T)0);
        DEGK_UNROLL for (int j = 0; j < (int)(sizeof(S.K) / sizeof(V)); ++j) ((V*)&S.K)[j] = V((T)0);
        S.singm = 0; S.solved = true; S.idle = false; S.svc = true;
    }

    static DEGK_DEV void drain(const KArgs& a, Ctx& c, int batch_min) {
        while (c.qcount >= batch_min && c.qcount > 0) {
            const int n = c.qcount < QBATCH ? c.qcount : QBATCH;
            __syncwarp();
            process_saves<T, Model, MethodS>(a, c.queue, c.qcount - n, n, c.sv_s);
            c.qcount -= n;
            __syncwarp();
        }
    }

    static DEGK_DEV void run(const KArgs& a, unsigned char* smem_raw) {
        Ctx c;
        c.abstol = (T)a.abstol; c.reltol = (T)a.reltol;
        c.has_saveat = a.saveat != nullptr;
        c.nsv = (int)opaque((u32)(c.has_saveat ? a.n_saveat : 0));
        c.lane = lane_id();
        c.lt_mask = (1u << c.lane) - 1u;
        const int warp_in_block = (int)(threadIdx.x >> 5);
        const int nwarps = (int)(blockDim.x >> 5);
        c.max_it = a.max_iters > 0x3fffffffLL ? 0x3fffffffu : (u32)a.max_iters;
        c.kInf = (T)__longlong_as_double(0x7ff0000000000000LL);   // +inf: "no further save point"
        c.dtmin = MethodS::dtmin();
        c.kDead = (T)-1;                     // h of a slot that is not integrating
        // fast controller constants in the L = log2(N * EEst^2) representation (degk_ode_kernels4.cuh, StepMath)
        const double lgN = log2((double)N), lgGamma = log2(9.0 / 10.0);
        c.b1h = (T)(0.5 * 7.0 / (10.0 * MethodS::ORDER));
        c.b2h = (T)(0.5 * 2.0 / (5.0 * MethodS::ORDER));
        c.k0 = (T)((0.5 * 7.0 / (10.0 * MethodS::ORDER) - 0.5 * 2.0 / (5.0 * MethodS::ORDER)) * lgN + lgGamma);
        c.lqZero = (T)lgN;                                      // qold = 1 (the reject branch ignores qold)
        c.exLo = (T)-2.321928094887362; c.exHi = (T)3.321928094887362;   // fac in [qmin, qmax] = [1/5, 10]
        c.lqInit = FAST ? (T)(2.0 * log2(1.0e-4) + lgN) : SM::lq_init();

        // shared memory: [per-warp save queues][per-warp problem pools][saveat copy + 2 x inf]
        c.queue = (Rec*)smem_raw + (size_t)warp_in_block * QCAP;
        c.queue_saddr = opaque((u32)__cvta_generic_to_shared(c.queue));
        c.pool = (T*)(smem_raw + (size_t)nwarps * QCAP * sizeof(Rec)) + (size_t)warp_in_block * 32 * PW;
        c.sv_s = (T*)(smem_raw + (size_t)nwarps * QCAP * sizeof(Rec)) + (size_t)nwarps * 32 * PW;
        for (int i = (int)threadIdx.x; i < c.nsv; i += (int)blockDim.x) c.sv_s[i] = ((const T*)a.saveat)[i];
        if (threadIdx.x < 2) c.sv_s[c.nsv + (int)threadIdx.x] = c.kInf;
        __syncthreads();
        c.sv_saddr = opaque((u32)__cvta_generic_to_shared(c.sv_s) - (u32)sizeof(T));   // 1-based
        c.qcount = 0;
        c.queue_sched = (a.schedule == SCHED_QUEUE);
        c.exhausted = !c.queue_sched; c.static_done = false;
        c.warp_global = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        c.RETIRE_BATCH = a.retire_batch > 0 ? a.retire_batch : (DEGK_RETIRE_BATCH * W) / 2;
        c.pool_base = 0; c.pool_n = 0; c.pool_pos = 0;
        c.tot_acc = c.tot_rej = c.tot_fail = 0;
        c.iter = 0;

        Stream A, B;
        Ctl xa, xb;
        init_stream(c, A); init_stream(c, B);
        bool startedA = false, startedB = false;
        if (c.queue_sched) { service(a, c, A, startedA); service(a, c, B, startedB); }
        else {
            fill_static(a, c, A, 0); fill_static(a, c, B, 1);
            // (a stream without work marks itself idle in its first service call)
            service(a, c, A, startedA); service(a, c, B, startedB);
        }
        // prologue: A's first stages (B's "control" of nothing: its slots carry no attempt yet -- run B's stages too, so
        // that the loop can start with [stages(A) done, control(A) || stages(B)] in its steady shape)
        {
            const V tv = PO::make(A.t), hv = PO::make(A.h);
            if constexpr (HK) A.solved = MethodV::template attempt_hk<true>(A.K, A.u, A.p, tv, hv, A.unew, A.err);
            else A.solved = MethodV::template attempt<true>(A.K, A.u, A.p, tv, hv, A.unew, A.err);
        }
        for (;;) {
            // ---- stages(B) with control(A) spliced in; then A is committed: drain, service A ----
            half_pass(c, B, A, xa);
            A.svc = __any_sync(0xffffffffu, xa.any_evt) | ((c.iter & (u32)(DEGK4_SERVICE_PERIOD - 1)) == 0u);
            drain(a, c, QBATCH);
            if (A.svc && !A.idle) service(a, c, A, startedA);
            // ---- stages(A) with control(B) spliced in; then B is committed: drain, service B ----
            half_pass(c, A, B, xb);
            B.svc = __any_sync(0xffffffffu, xb.any_evt) | ((c.iter & (u32)(DEGK4_SERVICE_PERIOD - 1)) == (u32)(DEGK4_SERVICE_PERIOD / 2));
            drain(a, c, QBATCH);
            if (B.svc && !B.idle) service(a, c, B, startedB);
            ++c.iter;
            if (A.idle && B.idle) break;
        }
        // flush the remaining deferred saves
        __syncwarp();
        drain(a, c, 1);
        add_totals<T>(a, c.tot_acc, c.tot_rej, c.tot_fail);
    }
};

template <class T, class Model, template <class, class> class MethodT, int W>
DEGK_DEV void ode_asolve5_body(const KArgs& a, unsigned char* smem_raw) {
    Asolve5<T, Model, MethodT, W>::run(a, smem_raw);
}

}  // namespace degk
