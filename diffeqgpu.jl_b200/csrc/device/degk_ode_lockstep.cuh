// degk_ode_lockstep.cuh -- fixed-dt ensemble kernel for the lock-step case (replaces reference kernels.jl:1-72 for it).
//
// The general fixed-dt kernel (degk_ode_kernels.cuh::ode_solve_body) serves every option of the reference's
// ode_solve_kernel with one thread per trajectory; ncu showed ~370 warp-instructions per step on BASELINE config 1
// of which ~120 are the stepper (profiles/r2_c1_fixed_fast_before_flat_flush.txt).  This kernel is the same
// algorithm specialised for the case that configuration is: every trajectory of the launch shares (t0, tf, dt)
// -- so all lanes take the same number of steps and the time loop is warp-uniform -- with every-step saves, no
// saveat grid, no events, an explicit RK stepper:
//   * the loop carries no per-lane activity flags, cursors or option tests;
//   * the fast Float32 build advances TWO trajectories per thread on packed register pairs (FFMA2) with the
//     h-scaled stage sums of the generated steppers, the strict build one trajectory with the reference's un-fused
//     arithmetic (bit-identical to ode_solve_body, which the oracle pins);
//   * the loop is unrolled twice so that u / unew swap roles instead of being copied;
//   * in the reference layout the rows are staged in shared memory, R steps at a time, as one strip per trajectory
//     and written out by the whole warp as one flattened copy (consecutive lanes = consecutive words); the save
//     times are the same for every trajectory, so they are staged once per warp, not once per lane; the flush
//     points of the warps are staggered so that stores and arithmetic overlap across the GPU.
// Semantics kept from the reference: row 0 is prob.u0, integ.t += dt precedes the stages, the loop runs while
// t < tf, a last step that overshoots tf is followed by the interpolated value at tf in the last row
// (kernels.jl:53-57), rows past `len` are dropped, unwritten ts rows keep t0.
#pragma once
#include "degk_ode_kernels4.cuh"

namespace degk {

// steppers that can run here say so (generated explicit RK: attempt() cannot fail, no on_accept work)
template <class M, class = void> struct lockstep_ok_of { static constexpr bool value = false; };
template <class M> struct lockstep_ok_of<M, typename replay_void_<decltype(M::LOCKSTEP_OK)>::type> { static constexpr bool value = M::LOCKSTEP_OK; };

template <class T> __host__ __device__ constexpr int lockstep_minblocks() { return sizeof(T) == 4 ? 4 : 1; }

// shared memory per block: per warp 32 W strips of (N R) | 1 words (odd: lanes hit distinct banks), then R save times
__host__ __device__ constexpr size_t lockstep_smem_bytes(int nwarps, int n, int w, int rows, size_t es) {
    return (size_t)nwarps * ((size_t)32 * w * (size_t)((n * rows) | 1) + (size_t)rows) * es;
}

// STAGED: rows go through shared memory (reference layout); otherwise straight to `us` / `ts`.  Two instantiations of
// one body, chosen once per launch, so that neither carries the other's pointers through the time loop.
template <bool STAGED, class T, class Model, template <class, class> class MethodT, int W>
DEGK_DEV void ode_solve_lockstep_run(const KArgs& a, unsigned char* smem_raw) {
    constexpr bool PACKED = slots4_packed<T, W>();
    typedef Slots4<T, W, PACKED> PO;
    typedef typename PO::type V;
    typedef MethodT<V, Model> MethodV;
    constexpr int N = Model::N;
    constexpr int NPA = Model::NP > 0 ? Model::NP : 1;
    constexpr bool HK = use_hk<MethodV>();
    static_assert(lockstep_ok_of<MethodV>::value, "stepper cannot run in the lock-step kernel");

    const u32 lane = lane_id();
    const i64 warp_first = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 * W);   // first trajectory of this warp
    if (warp_first >= a.n_traj) return;
    const i64 left = a.n_traj - warp_first;
    const int nstrips = left >= 32 * W ? 32 * W : (int)left;      // trajectories of this warp (warp-uniform)
    i64 traj[W];
    bool valid[W];
    DEGK_UNROLL for (int s = 0; s < W; ++s) { traj[s] = warp_first + lane + 32 * s; valid[s] = (int)lane + 32 * s < nstrips; }

    // problem data: slots without a trajectory shadow the warp's first one and never store
    V ua[N], ub[N], err[N], p[NPA];
    T t0 = (T)0, tf = (T)0;
    {
        T us_[W][N], ps_[W][NPA];
        DEGK_UNROLL for (int s = 0; s < W; ++s) load_problem<T, Model>(a, valid[s] ? traj[s] : warp_first, us_[s], ps_[s], t0, tf);
        DEGK_UNROLL for (int c = 0; c < N; ++c) { T x[W]; DEGK_UNROLL for (int s = 0; s < W; ++s) x[s] = us_[s][c]; ua[c] = PO::make(x); ub[c] = ua[c]; }
        DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) { T x[W]; DEGK_UNROLL for (int s = 0; s < W; ++s) x[s] = ps_[s][c]; p[c] = PO::make(x); }
        DEGK_UNROLL for (int s = 0; s < W; ++s)
            if (valid[s]) { store_u<T, N>(a, traj[s], 0, us_[s]); store_t<T>(a, traj[s], 0, t0); }     // row 0 = prob.u0
    }
    const T dt = (T)a.dt;
    const V dtv = V(dt);

    // staging buffers of this warp
    const int R = a.stage_rows;
    const int S = (N * R) | 1;                                    // words per strip (odd: conflict-free column access)
    T* wu = (T*)smem_raw + (size_t)(threadIdx.x >> 5) * ((size_t)32 * W * S + R);
    T* wt = wu + (size_t)32 * W * S;
    int nbuf = 0;                                                 // rows staged (warp-uniform)
    i64 k0 = 1;                                                   // row of the first staged step
    // Every warp of the launch runs the same loop in the same rhythm, so with a common flush period the whole GPU
    // would alternate between compute-only and store-only phases (measured on C1 at 10^6: the store bursts added
    // ~0.5 ms to 0.6 ms of compute).  The first flush of warp w comes after 1 + (w mod R) rows: at any time 1/R of
    // the warps are storing while the others compute.
    int room = STAGED ? 1 + (int)((warp_first / (32 * W)) % R) : 0;   // rows until the next flush
    T* srow[W];                                                   // where this lane stages the next row of its trajectories
    DEGK_UNROLL for (int s = 0; s < W; ++s) srow[s] = wu + (size_t)(lane + 32 * s) * S;
    // unstaged saves (trajectory-major layout): per-slot output pointers that advance by one row per step
    const bool soa = a.out_layout != LAYOUT_REF;
    const i64 ustep = soa ? (i64)N * a.n_traj : (i64)N;           // words between two rows of one trajectory
    const i64 ucomp = soa ? a.n_traj : 1;                         // words between two components of one row
    const i64 tstep = soa ? a.n_traj : 1;
    T* pu[W];
    T* pt[W];
    DEGK_UNROLL for (int s = 0; s < W; ++s) {
        pu[s] = (T*)a.us + (soa ? traj[s] : traj[s] * a.n_rows * N) + ustep;                 // row 1
        pt[s] = a.ts ? (T*)a.ts + (soa ? traj[s] : traj[s] * a.n_rows) + tstep : nullptr;
    }

    // Copy the staged [nstrips][n0 rows][N] block to the trajectories' strips of `us`: 32 consecutive words of one
    // strip per store instruction (column passes over the strip, the warp walks down the strips), so the only
    // address arithmetic in the loop is two constant increments.  The save times are the same for every strip; with
    // at most 16 rows staged two strips share one store instruction.
    auto flush = [&]() {
        __syncwarp();
        i64 n0 = a.n_rows - k0;                                   // rows past `len` are dropped
        n0 = n0 < 0 ? 0 : (n0 < nbuf ? n0 : nbuf);
        if (n0 > 0) {
            const int L = (int)n0 * N;
            const i64 rs = (i64)a.n_rows * N;
            T* const obase = (T*)a.us + ((i64)warp_first * a.n_rows + k0) * N;
            for (int w = (int)lane; w < L; w += 32) {
                T* o = obase + w;
                const T* sm = wu + w;
                int j = 0;
                // eight strips per batch: all loads first, then all stores (written out so that the eight values are
                // live at once -- left to itself the compiler recycles four registers and serialises load -> store)
                for (; j + 8 <= nstrips; j += 8) {
                    T v[8];
                    DEGK_UNROLL for (int q = 0; q < 8; ++q) v[q] = sm[q * S];
                    DEGK_UNROLL for (int q = 0; q < 8; ++q) { *o = v[q]; o += rs; }
                    sm += 8 * S;
                }
                for (; j < nstrips; ++j) { *o = *sm; o += rs; sm += S; }
            }
            if (a.ts != nullptr) {
                const int nn = (int)n0;
                T* const tbase = (T*)a.ts + (i64)warp_first * a.n_rows + k0;
                if (nn <= 16) {
                    const int half = (int)(lane >> 4), w = (int)(lane & 15u);
                    if (w < nn) {
                        const T v = wt[w];
                        T* o = tbase + (i64)half * a.n_rows + w;
                        const i64 two_rows = 2 * a.n_rows;
                        _Pragma("unroll 4")
                        for (int j = half; j < nstrips; j += 2) { *o = v; o += two_rows; }
                    }
                } else {
                    for (int w = (int)lane; w < nn; w += 32) {
                        const T v = wt[w];
                        T* o = tbase + w;
                        for (int j = 0; j < nstrips; ++j) { *o = v; o += a.n_rows; }
                    }
                }
            }
        }
        k0 += nbuf;
        nbuf = 0;
        DEGK_UNROLL for (int s = 0; s < W; ++s) srow[s] = wu + (size_t)(lane + 32 * s) * S;
        __syncwarp();
    };

    typename MethodV::Keep K;
    MethodV::init(K, ua, p, V(t0));
    T t = t0, tprev = t0;
    u32 nsteps = 0;
    i64 step_idx = 1;                                             // row of the next every-step save
    int rc = RC_SUCCESS;
    bool first = true;
    const i64 max_it = a.max_iters;

    // one step from `uin` into `uout`; returns false when the loop ends
    auto step = [&](V (&uin)[N], V (&uout)[N]) -> bool {
        if (!first) MethodV::accepted(K);                        // FSAL shift deferred so the last step's stages
        first = false;                                            // survive for the final interpolation
        tprev = t;
        t = t + dt;                                               // integ.t += dt precedes the stages
        if constexpr (HK) MethodV::template attempt_hk<false>(K, uin, p, V(tprev), dtv, uout, err);
        else MethodV::template attempt<false>(K, uin, p, V(tprev), dtv, uout, err);
        ++nsteps;
        if constexpr (STAGED) {                                   // integrator_utils.jl:28-33, staged
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                DEGK_UNROLL for (int c = 0; c < N; ++c) srow[s][c] = PO::get(uout[c], s);
                srow[s] += N;
            }
            if (lane == 0) wt[nbuf] = t;
            ++nbuf;
            if (--room == 0) { flush(); room = R; }
        } else {
            const bool in_range = step_idx < a.n_rows;            // rows past `len` are dropped
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                if (valid[s] & in_range) {
                    T* q = pu[s];
                    DEGK_UNROLL for (int c = 0; c < N; ++c) { *q = PO::get(uout[c], s); q += ucomp; }
                    if (pt[s]) *pt[s] = t;
                }
                pu[s] += ustep;
                if (pt[s]) pt[s] += tstep;
            }
        }
        ++step_idx;
        if ((i64)nsteps >= max_it) { rc = RC_MAXITERS; return false; }
        return t < tf;
    };

    bool last_in_a = true;                                        // the last step read ua and wrote ub
    if (t < tf) {
        for (;;) {
            last_in_a = true;
            if (!step(ua, ub)) break;
            last_in_a = false;
            if (!step(ub, ua)) break;
        }
    } else {
        DEGK_UNROLL for (int c = 0; c < N; ++c) ub[c] = ua[c];    // no step: u = u0
    }
    if (STAGED && nbuf > 0) flush();

    // final state and the step it came from
    V u[N], uprev[N];
    DEGK_UNROLL for (int c = 0; c < N; ++c) {
        u[c] = last_in_a ? ub[c] : ua[c];
        uprev[c] = last_in_a ? ua[c] : ub[c];
    }
    u32 nfail = 0, nmine = 0;
    if (rc == RC_SUCCESS && t > tf) {                             // kernels.jl:53-57: the value at tf goes into the last row
        const V theta = V((tf - tprev) / dt);
        V v[N];
        if constexpr (HK) MethodV::interp_hk(K, theta, dtv, uprev, u, p, V(tprev), v);
        else MethodV::interp(K, theta, dtv, uprev, u, p, V(tprev), v);
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            if (valid[s]) {
                T o[N];
                DEGK_UNROLL for (int c = 0; c < N; ++c) o[c] = PO::get(v[c], s);
                store_u<T, N>(a, traj[s], a.n_rows - 1, o);
                store_t<T>(a, traj[s], a.n_rows - 1, tf);
            }
        }
    }
    DEGK_UNROLL for (int s = 0; s < W; ++s) {
        if (!valid[s]) continue;
        int rcs = rc;
        if (rc == RC_SUCCESS) {
            bool fin = true;
            DEGK_UNROLL for (int c = 0; c < N; ++c) fin = fin && finite_(PO::get(u[c], s));
            if (!fin) rcs = RC_UNSTABLE;
        }
        if (rcs != RC_SUCCESS) ++nfail;
        fill_unwritten_ts<T>(a, traj[s], step_idx, t0);
        if (a.retcode) a.retcode[traj[s]] = rcs;
        if (a.naccept) a.naccept[traj[s]] = (int)nsteps;
        if (a.nreject) a.nreject[traj[s]] = 0;
        ++nmine;
    }
    add_totals<T>(a, nsteps * nmine, 0u, nfail);
}

template <class T, class Model, template <class, class> class MethodT, int W>
DEGK_DEV void ode_solve_lockstep_body(const KArgs& a, unsigned char* smem_raw) {
    if (a.stage_rows > 0) ode_solve_lockstep_run<true, T, Model, MethodT, W>(a, smem_raw);
    else ode_solve_lockstep_run<false, T, Model, MethodT, W>(a, smem_raw);
}

}  // namespace degk
