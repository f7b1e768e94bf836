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
//   * in the reference layout every trajectory owns a contiguous strip of `us` (len * n values) at an arbitrary
//     4-byte alignment, and what bounds the kernel is how the memory system digests the stores (measured: with the
//     stores removed the rest runs 2.2x faster; pieces that end inside a 32-byte sector make the L2 fetch the sector
//     from DRAM).  So the rows go through a 16- or 32-row buffer per trajectory in shared memory, and every 8 / 24
//     steps the warp writes, for each trajectory, the words up to THAT trajectory's last sector boundary -- whole
//     sectors, 16-byte loads and stores, four trajectories per instruction -- and the owning lane moves the few
//     words behind the boundary to the front of the buffer.  Only the first and the last piece of a trajectory are
//     ragged.  The save times are the same for every trajectory: one ring per warp.
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

// Rows a trajectory can buffer: 32 with one trajectory per thread, 16 with two (the same ~53 KB per block for a
// three-component Float32 state).  A flush comes every (rows - 8) steps -- at most 7 left-over rows of the save times
// plus (rows - 8 + 1) new ones fit -- so the one-trajectory builds, which are issue-bound, amortise the per-trajectory
// part of a flush over three times as many rows.
__host__ __device__ constexpr int lockstep_ring_rows(int w) { return 32 / w; }
// words between the buffers of two trajectories: the rows, rounded up to a multiple of four words plus four (16-byte
// aligned for the vector loads of the flush, and not a multiple of 32 banks)
__host__ __device__ constexpr int lockstep_buf_words(int n, int w) { return ((n * lockstep_ring_rows(w) + 3) & ~3) + 4; }
// shared memory per block: per warp 32 W buffers, then the ring of save times (one entry per buffered row)
__host__ __device__ constexpr size_t lockstep_smem_bytes(int nwarps, int n, int w, size_t es) {
    return (size_t)nwarps * ((size_t)32 * w * (size_t)lockstep_buf_words(n, w) + (size_t)lockstep_ring_rows(w)) * es;
}

// Sector-aligned window of one trajectory at a flush, in words relative to the trajectory's first word (NW words per
// row): `phi` = (absolute word index of the trajectory's first word) mod (words per sector), so relative word r sits on a
// sector boundary when (r + phi) is a multiple of SW.  F = what earlier flushes reached (the last boundary not beyond
// row kpr, or 0), E = where this one ends (the last boundary not beyond row khi; row khi itself when `final`),
// [Fa, Ea) = the whole sectors in between, [F, Fa) the ragged first words of a trajectory (F = 0 off a boundary),
// [Ea, E) the ragged last ones (`final`).  F of a flush equals E of the one before by construction.
template <int SW>
DEGK_DEV void lockstep_window(int phi, int wpr, int whi, bool final, int& F, int& Fa, int& Ea, int& E) {
    F = ((wpr + phi) & ~(SW - 1)) - phi;
    F = F > 0 ? F : 0;
    const int Ef = ((whi + phi) & ~(SW - 1)) - phi;
    E = final ? whi : (Ef > F ? Ef : F);
    Fa = ((F + phi + SW - 1) & ~(SW - 1)) - phi;                  // first boundary at or after F (F itself unless F = 0 is off one)
    Fa = Fa < E ? Fa : E;
    Ea = ((E + phi) & ~(SW - 1)) - phi;                           // last boundary at or before E
    Ea = Ea > Fa ? Ea : Fa;
}

// Flush one output array for every trajectory of the warp (warp-collective; LPS lanes serve one trajectory): the whole
// sectors of the window go out as 16-byte stores, the ragged first words of a trajectory (first flush) and last words
// (`final`) one by one.  LINEAR: trajectory j has its own buffer at buf0 + j * stride whose word 0 is relative word F
// (the owning lane compacts what is left after every flush), so in steady state the reads are 16-byte loads too;
// otherwise buf0 is one ring of RING rows shared by all trajectories (the save times: position = row mod RING).
// Not inlined: one copy per array kind keeps the time loop inside the instruction cache.
template <class T, int NW, int LPS, bool LINEAR, int RING>
__device__ __noinline__ void lockstep_flush_array(T* base, const T* buf0, int stride, i64 warp_first, int nstrips,
                                                  i64 n_rows, i64 k_prev, i64 k_hi, bool final) {
    constexpr int SW = 32 / (int)sizeof(T);                       // words per 32-byte sector
    constexpr int QW = 16 / (int)sizeof(T);                       // words per 16-byte store
    constexpr int PER_IT = 32 / LPS;
    static_assert(LINEAR || NW == 1, "the shared ring holds one word per row");
    const int lane = (int)lane_id();
    const i64 khi = k_hi < n_rows ? k_hi : n_rows;                // rows past `len` are dropped
    const i64 kpr = k_prev < khi ? k_prev : khi;
    const int wpr = (int)kpr * NW, whi = (int)khi * NW;           // relative words reached before / now
    const int sub = lane / LPS, q0 = lane % LPS;
    const i64 row_words = n_rows * NW;
    T* tb = base + (warp_first + sub) * row_words;                // first word of this lane's trajectory
    const int rw = (int)(row_words & (SW - 1));
    int phi = (int)((((unsigned long long)base / sizeof(T)) + (unsigned long long)((warp_first + sub) * row_words)) & (unsigned long long)(SW - 1));
    const T* bj = buf0 + (size_t)sub * stride;
    if (!final && kpr >= SW) {
        // steady state: both ends of the window are sector boundaries (lockstep_window reduces to two roundings)
        for (int j = sub; j < nstrips; j += PER_IT, tb += PER_IT * row_words, bj += PER_IT * stride, phi = (phi + PER_IT * rw) & (SW - 1)) {
            const int F = ((wpr + phi) & ~(SW - 1)) - phi, E = ((whi + phi) & ~(SW - 1)) - phi;
            const int nq = (E - F) / QW;
            T* dst = tb + F + q0 * QW;
            _Pragma("unroll 1")
            for (int q = q0; q < nq; q += LPS, dst += LPS * QW) {
                uint4 pk;
                if constexpr (LINEAR) {
                    pk = *reinterpret_cast<const uint4*>(bj + q * QW);
                } else {
                    const int r = F + q * QW;
                    T v[QW];
                    DEGK_UNROLL for (int i = 0; i < QW; ++i) v[i] = bj[(r + i) & (RING - 1)];
                    if constexpr (sizeof(T) == 4) {
                        pk = make_uint4(__float_as_uint((float)v[0]), __float_as_uint((float)v[1]), __float_as_uint((float)v[QW - 2]), __float_as_uint((float)v[QW - 1]));
                    } else {
                        const unsigned long long l0 = (unsigned long long)__double_as_longlong((double)v[0]), l1 = (unsigned long long)__double_as_longlong((double)v[QW - 1]);
                        pk = make_uint4((unsigned)l0, (unsigned)(l0 >> 32), (unsigned)l1, (unsigned)(l1 >> 32));
                    }
                }
                *reinterpret_cast<uint4*>(dst) = pk;
            }
        }
        return;
    }
    for (int j = sub; j < nstrips; j += PER_IT, tb += PER_IT * row_words, bj += PER_IT * stride, phi = (phi + PER_IT * rw) & (SW - 1)) {
        int F, Fa, Ea, E;
        lockstep_window<SW>(phi, wpr, whi, final, F, Fa, Ea, E);
        auto word = [&](int r) -> T { return LINEAR ? bj[r - F] : bj[(r / NW) & (RING - 1)]; };   // relative word r
        const int nq = (Ea - Fa) / QW;
        T* dst = tb + Fa + q0 * QW;
        _Pragma("unroll 1")                                       // at most one trip per lane except at the first flush
        for (int q = q0; q < nq; q += LPS, dst += LPS * QW) {
            const int r = Fa + q * QW;
            uint4 pk;
            if (LINEAR && Fa == F) {                              // steady state: the buffer starts on the boundary
                pk = *reinterpret_cast<const uint4*>(bj + q * QW);
            } else {
                T v[QW];
                DEGK_UNROLL for (int i = 0; i < QW; ++i) v[i] = word(r + i);
                if constexpr (sizeof(T) == 4) {
                    pk = make_uint4(__float_as_uint((float)v[0]), __float_as_uint((float)v[1]), __float_as_uint((float)v[QW - 2]), __float_as_uint((float)v[QW - 1]));
                } else {
                    const unsigned long long l0 = (unsigned long long)__double_as_longlong((double)v[0]), l1 = (unsigned long long)__double_as_longlong((double)v[QW - 1]);
                    pk = make_uint4((unsigned)l0, (unsigned)(l0 >> 32), (unsigned)l1, (unsigned)(l1 >> 32));
                }
            }
            *reinterpret_cast<uint4*>(dst) = pk;
        }
        // ragged head [F, Fa) of a trajectory's first piece and tail [Ea, E) of its last: fewer than SW words each
        _Pragma("unroll 1")
        for (int r = F + q0; r < Fa; r += LPS) tb[r] = word(r);
        _Pragma("unroll 1")
        for (int r = Ea + q0; r < E; r += LPS) tb[r] = word(r);
    }
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
        if constexpr (!STAGED) {
            DEGK_UNROLL for (int s = 0; s < W; ++s)
                if (valid[s]) { store_u<T, N>(a, traj[s], 0, us_[s]); store_t<T>(a, traj[s], 0, t0); } // row 0 = prob.u0
        }
    }
    const T dt = (T)a.dt;
    const V dtv = V(dt);

    // ---- staged saves (reference layout): rings in shared memory, sector-aligned flushes ----
    constexpr int RING = lockstep_ring_rows(W), PERIOD = RING - 8;
    constexpr int S = lockstep_buf_words(N, W);                   // words between the buffers of two trajectories
    constexpr int SW = 32 / (int)sizeof(T);                       // words per 32-byte sector
    T* wu = (T*)smem_raw + (size_t)(threadIdx.x >> 5) * ((size_t)32 * W * S + RING);
    T* wt = wu + (size_t)32 * W * S;
    i64 k_prev = 0, k_hi = 0;                                     // rows [0, k_prev) were offered to a flush, [k_prev, k_hi) are new
    // Every warp of the launch runs the same loop in the same rhythm; the first flush of warp w comes after
    // 7 + (w mod PERIOD) rows (row 0 included that is at least 8 rows, so every trajectory has reached a sector
    // boundary, and at most RING) and then every PERIOD: the warps store at different times.
    int room = STAGED ? 7 + (int)((warp_first / (32 * W)) % PERIOD) : 0;   // rows until the next flush
    T* fill[W];                                                   // where this lane stages the next row of its trajectories
    int phis[W];                                                  // sector phase of their first words in `us`
    DEGK_UNROLL for (int s = 0; s < W; ++s) {
        fill[s] = wu + (size_t)(lane + 32 * s) * S;
        phis[s] = (int)((((unsigned long long)a.us / sizeof(T)) + (unsigned long long)(traj[s] * a.n_rows * N)) & (unsigned long long)(SW - 1));
    }

    auto flush = [&](bool final) {
        __syncwarp();
        lockstep_flush_array<T, N, 8, true, RING>((T*)a.us, wu, S, warp_first, nstrips, a.n_rows, k_prev, k_hi, final);
        // (the save times of a steady-state flush are PERIOD rows = PERIOD / 4 sixteen-byte stores per trajectory)
        if (a.ts != nullptr) lockstep_flush_array<T, 1, (PERIOD * (int)sizeof(T) / 16 >= 8 ? 8 : (PERIOD * (int)sizeof(T) / 16 >= 4 ? 4 : 2)), false, RING>((T*)a.ts, wt, 0, warp_first, nstrips, a.n_rows, k_prev, k_hi, final);
        __syncwarp();
        if (!final) {
            // every lane moves what its trajectories keep (the words behind their last sector boundary, fewer than
            // a sector) to the front of their buffers
            const i64 khi = k_hi < a.n_rows ? k_hi : a.n_rows, kpr = k_prev < khi ? k_prev : khi;
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                int F, Fa, Ea, E;
                if (kpr >= SW) {                                  // steady state: two roundings (see lockstep_flush_array)
                    F = (((int)kpr * N + phis[s]) & ~(SW - 1)) - phis[s];
                    E = (((int)khi * N + phis[s]) & ~(SW - 1)) - phis[s];
                } else {
                    lockstep_window<SW>(phis[s], (int)kpr * N, (int)khi * N, false, F, Fa, Ea, E);
                }
                T* b = wu + (size_t)(lane + 32 * s) * S;
                const int keep = (int)khi * N - E, off = E - F;
                if (off > 0) {
                    _Pragma("unroll 1")
                    for (int i = 0; i < keep; ++i) b[i] = b[off + i];
                }
                fill[s] = b + keep;
            }
            __syncwarp();
        }
        k_prev = k_hi;
    };
    // stage row k_hi (state `v`, time `tt`) of every trajectory of this lane
    auto stage_row = [&](const V (&v)[N], T tt) {
        if (k_hi < a.n_rows) {                                    // rows past `len` are dropped
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                DEGK_UNROLL for (int c = 0; c < N; ++c) fill[s][c] = PO::get(v[c], s);
                fill[s] += N;
            }
            if (lane == 0) wt[(int)(k_hi & (RING - 1))] = tt;
        }
        ++k_hi;
    };

    // ---- unstaged saves (trajectory-major layout): per-slot output pointers that advance by one row per step ----
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
    if constexpr (STAGED) stage_row(ua, t0);                      // row 0 = prob.u0

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
            stage_row(uout, t);
            --room;
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
            bool more = step(ua, ub);
            if (more) { last_in_a = false; more = step(ub, ua); }
            if (!more) break;
            if constexpr (STAGED) {                               // one flush site per two steps: 8 or 9 new rows + <= 7 left over
                if (room <= 0) { flush(false); room += PERIOD; }
            }
        }
    } else {
        DEGK_UNROLL for (int c = 0; c < N; ++c) ub[c] = ua[c];    // no step: u = u0
    }
    if constexpr (STAGED) flush(true);

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
