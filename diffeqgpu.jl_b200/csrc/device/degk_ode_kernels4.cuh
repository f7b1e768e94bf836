// degk_ode_kernels4.cuh -- fourth-generation adaptive ensemble kernel: ONE body for both fp modes.
//
// Same job as the earlier generations (reference kernels.jl:74-152 + the adaptive step! of each solver):
// persistent warps, a per-warp problem pool fed from a global work queue, W trajectories ("slots") per
// thread, deferred batched saves.  What is new:
//
//  (1) ONE CONTROL FLOW, TWO ARITHMETIC POLICIES.  The strict build (un-fused FMUL/FADD, IEEE div/sqrt, the
//      reference's PI controller with Float32 `^` through FP64) and the fast build (FFMA2 pairs, MUFU,
//      log-domain controller) instantiate the same loop, service path, save queue and replay; they differ
//      only in StepMath<> below.  The strict instantiation is compared bit for bit with the CPU oracle
//      (tests/test_gpu_parity.py), which pins the queue / push / multi-crossing / retire logic of the kernel
//      that produces the headline number.
//  (2) The cost model changed.  tools/pipe_probe.cu (profiles/r2_pipe_probe*.jsonl) shows that on sm_100 the
//      costs of the instructions of one warp ADD UP at the dispatch port: ~2.08 cycles per packed FP32
//      instruction with two register-pair operands (3.06 with three), ~0.8-1.1 per two-source ALU instruction,
//      ~2 per three-source integer instruction, ~4 per POPC -- nothing hides "in the shadow" of the FMA pipe.
//      So: (a) stage sums use the h-scaled form of the generated steppers (attempt_hk: immediates + two
//      register pairs), (b) everything that is not stage arithmetic is counted in instructions:
//        - the save queue is lane-private (degk_ode_saves.cuh): a push is two predicated stores, two predicated
//          adds and one predicated load of the next save time; ranks, compaction and the cursor fix-up after a
//          step across several save points live in the service path;
//        - the fast build derives the landing on tf from the exact remainder tf - (t + h) and lets a finished
//          slot die by itself (its next step is < dtmin): no stop / park selects in the loop;
//        - the maximum-iteration test only raises the service flag, the service path parks the slot;
//        - trajectory start-up work that does not depend on the integration (row 0, the t0 pre-fill of ts,
//          empty or invalid time spans) is done by the full warp when it loads 32 problems into its pool;
//          retiring a slot is a handful of scattered 4-byte stores;
//        - the service path tests "is a retire batch due" first and leaves in ~12 instructions otherwise.
#pragma once
#include "degk_ode_saves.cuh"
#include "degk_dae_init.cuh"

#ifndef DEGK4_HK
#define DEGK4_HK 1          // h-scaled stage sums in the fast build
#endif
#ifndef DEGK4_SERVICE_PERIOD
#define DEGK4_SERVICE_PERIOD 8   // passes between two looks at the stopped slots (power of two)
#endif

namespace degk {

// ------------------------------------------------------------------------------------------
// Sc2: two independent scalar trajectories per thread (the strict build's W = 2, and Float64).
// Every operation is the scalar operation on each half, so rounding is exactly the one-slot kernel's.
template <class T>
struct Sc2 {
    T a, b;
    DEGK_DEV Sc2() {}
    DEGK_DEV Sc2(T x, T y) : a(x), b(y) {}
    DEGK_DEV explicit Sc2(T x) : a(x), b(x) {}
    template <class U> DEGK_DEV explicit Sc2(U x) : a((T)x), b((T)x) {}
    DEGK_DEV T get(int s) const { return s ? b : a; }
    friend DEGK_DEV Sc2 operator*(Sc2 x, Sc2 y) { return Sc2(x.a * y.a, x.b * y.b); }
    friend DEGK_DEV Sc2 operator+(Sc2 x, Sc2 y) { return Sc2(x.a + y.a, x.b + y.b); }
    friend DEGK_DEV Sc2 operator-(Sc2 x, Sc2 y) { return Sc2(x.a - y.a, x.b - y.b); }
    friend DEGK_DEV Sc2 operator/(Sc2 x, Sc2 y) { return Sc2(x.a / y.a, x.b / y.b); }
    friend DEGK_DEV Sc2 operator-(Sc2 x) { return Sc2(-x.a, -x.b); }
    friend DEGK_DEV Sc2 fma_(Sc2 x, Sc2 y, Sc2 z) { return Sc2(degk::fma_(x.a, y.a, z.a), degk::fma_(x.b, y.b, z.b)); }
    friend DEGK_DEV Sc2 abs_(Sc2 x) { return Sc2(degk::abs_(x.a), degk::abs_(x.b)); }
    friend DEGK_DEV Sc2 sqrt_(Sc2 x) { return Sc2(degk::sqrt_(x.a), degk::sqrt_(x.b)); }
    friend DEGK_DEV Sc2 sqrt(Sc2 x) { return Sc2(::sqrt(x.a), ::sqrt(x.b)); }
    friend DEGK_DEV Sc2 sin(Sc2 x) { return Sc2(::sin(x.a), ::sin(x.b)); }
    friend DEGK_DEV Sc2 cos(Sc2 x) { return Sc2(::cos(x.a), ::cos(x.b)); }
    friend DEGK_DEV Sc2 exp(Sc2 x) { return Sc2(::exp(x.a), ::exp(x.b)); }
    friend DEGK_DEV Sc2 log(Sc2 x) { return Sc2(::log(x.a), ::log(x.b)); }
};
template <class T> DEGK_DEV void assign_if(const bool* f, Sc2<T>& x, Sc2<T> y) { x.a = f[0] ? y.a : x.a; x.b = f[1] ? y.b : x.b; }
template <class T> DEGK_DEV Sc2<T> blendm(unsigned m, Sc2<T> x, Sc2<T> y) { return Sc2<T>((m & 1u) ? x.a : y.a, (m & 2u) ? x.b : y.b); }

// value type that carries W trajectories in this generation: packed pairs where fused arithmetic is allowed
template <class T, int W, bool PACKED> struct Slots4 : PackOf<T, W> {};
template <class T> struct Slots4<T, 2, false> {
    typedef Sc2<T> type;
    static DEGK_DEV T get(Sc2<T> v, int s) { return s ? v.b : v.a; }
    static DEGK_DEV Sc2<T> make(const T (&s)[2]) { return Sc2<T>(s[0], s[1]); }
    static DEGK_DEV Sc2<T> set(Sc2<T> v, int s, T x) { return s == 0 ? Sc2<T>(x, v.b) : Sc2<T>(v.a, x); }
};
template <class T, int W> __host__ __device__ constexpr bool slots4_packed() { return !DEGK_STRICT && W == 2 && sizeof(T) == 4; }

// save-queue capacity per warp in records: 32 lanes x depth, plus the records' worth of bytes that hold the
// flush directory (one u16 per record)
template <class T, int N, int W> __host__ __device__ constexpr int asolve4_qcap() {
    return 32 * save_queue_depth<T, N, W>() + (int)((64 * save_queue_depth<T, N, W>() + sizeof(SaveRec<T, N>) - 1) / sizeof(SaveRec<T, N>));
}
template <class T, int N, int NP> __host__ __device__ constexpr int asolve4_pool_words() { return N + NP + 3; }   // u0, p, t0, tf, first cursor
template <class T, int N, int NP, int W>
__host__ __device__ constexpr size_t asolve4_smem_bytes(int nwarps, int nsv) {
    return (size_t)nwarps * asolve4_qcap<T, N, W>() * sizeof(SaveRec<T, N>) + (size_t)nwarps * 32 * asolve4_pool_words<T, N, NP>() * sizeof(T) +
           ((size_t)nsv + 2) * sizeof(T);
}

template <class M, class = void> struct has_hk_of { static constexpr bool value = false; };
template <class M> struct has_hk_of<M, typename replay_void_<decltype(M::HAS_HK)>::type> { static constexpr bool value = M::HAS_HK; };

// resident blocks (of DEGK_BLOCK2 = 128 threads) per SM the kernel is compiled for.  Measured on C2 (Lorenz,
// GPUTsit5, 8.4 M trajectories; tools/c2_probe.cu, profiles/r2_c2_lanequeue.md): packed fast build 4 (127 registers,
// no spills): 134.8, 5 (96 registers, ~100 B of loop-carried state in local memory): 131.5 G steps/s; one-slot
// strict build 4: 38.2, 5: 40.3, 6 (80 registers): 41.4 (47.8 with the shared controller logarithm).
// Steppers with more live stage vectors (Vern7/9, the Rosenbrock family) keep the 128-register budget.
template <class T, class MS> __host__ __device__ constexpr int asolve4_minblocks() {
    if (sizeof(T) != 4) return 1;
    if (!MS::ALWAYS_SOLVED) return 4;
    if (has_hk_of<MS>::value) return DEGK_STRICT ? 6 : 4;        // FSAL, <= 7 stages
    return 4;
}

// does the stepper offer the h-scaled attempt / dense output (generated explicit RK, FSAL, no extra stages)?
template <class M> __host__ __device__ constexpr bool use_hk() { return !DEGK_STRICT && DEGK4_HK && has_hk_of<M>::value; }

template <bool HK, class M, class V, class KeepT, int N>
DEGK_DEV bool attempt4(KeepT& K, const V (&u)[N], const V* p, V t, V h, V (&unew)[N], V (&err)[N]) {
    if constexpr (HK) return M::template attempt_hk<true>(K, u, p, t, h, unew, err);
    else return M::template attempt<true>(K, u, p, t, h, unew, err);
}

// ------------------------------------------------------------------------------------------
// StepMath: the arithmetic that differs between the fp modes (error norm, PI controller).
//   strict: gpu_tsit5_perform_step.jl:121-137 operation by operation (same as ode_asolve2_body / the oracle)
//   fast:   log-domain PI controller.  With L = log2(N * EEst^2) (no mean, no square root) the accept and the reject
//           branch share one exponent: fac = 2^clamp(-b1*L/2 + b2*lq'/2 + log2(gamma)), lq' = 0 on reject
//           (dt / min(1/qmin, q11/gamma) = dt * max(qmin, fac); the upper clamp is inactive there) -- one MUFU.LG2 and
//           one MUFU.EX2 replace the square root, two powers and three divisions
// The controller memory `lq` holds qold^beta2 (strict) or log2(N * qold^2) (fast).
template <class T, int ORDER, int N, bool FAST> struct StepMath;

// EEst^beta1 and max(EEst, qoldinit)^beta2 for the strict controller.  Float32: Julia's `^` is
// Float32(exp2(log2(Float64 x) * Float64 y)) (degk_common.cuh pow_), so the two powers of one base share the
// Float64 logarithm -- the same double feeds both exp2, bit for bit what two separate pow_ calls return, for one
// FP64 log2 less per attempt (the controller memory holds qold^beta2 instead of qold for that reason).
template <int ORDER, class T>
DEGK_DEV void ctl_powers(T EEst, T lq0, T& q11, T& lq_acc) {
    typedef Ctl<T, ORDER> C;
    q11 = pow_(EEst, C::beta1());
    lq_acc = EEst < C::qoldinit() ? lq0 : pow_(EEst, C::beta2());       // NaN propagates like jl_max
}
#if DEGK_STRICT
template <int ORDER>
DEGK_DEV void ctl_powers(float EEst, float lq0, float& q11, float& lq_acc) {
    typedef Ctl<float, ORDER> C;
    const double lg = log2((double)EEst);
    const float p1 = (float)exp2(lg * (double)C::beta1()), p2 = (float)exp2(lg * (double)C::beta2());
    q11 = EEst == 1.0f ? 1.0f : p1;
    lq_acc = EEst < C::qoldinit() ? lq0 : (EEst == 1.0f ? 1.0f : p2);
}
#endif

template <class T, int ORDER, int N>
struct StepMath<T, ORDER, N, false> {
    typedef Ctl<T, ORDER> C;
    static DEGK_DEV T lq_init() { return pow_(C::qoldinit(), C::beta2()); }
    // per slot: sum of squares of tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol)
    static DEGK_DEV T scaled_sq(T uo, T un, T e, T abstol, T reltol) {
        const T sc = abstol + jl_max(abs_(uo), abs_(un)) * reltol;
        const T v = e / sc;
        return v * v;
    }
    // lq = qold^beta2, lq0 = qoldinit^beta2
    // -> reject?, candidate step (before the tf clamp), controller memory after an accepted step
    static DEGK_DEV void control(T accn, T lq, T lq0, T h, bool& reject, T& hf, T& lq_acc) {
        const T EEst = sqrt_(mean_<T, N>(accn));
        T q11;
        ctl_powers<ORDER>(EEst, lq0, q11, lq_acc);       // lq_acc = max(EEst, qoldinit)^beta2
        T q;
        if (EEst == (T)0) {
            q = (T)1 / C::qmax();
            q11 = (T)0;
        } else {
            q = q11 / lq;
        }
        reject = EEst > (T)1;
        if (reject) {
            hf = h / jl_min((T)1 / C::qmin(), q11 / C::gamma());
        } else {
            q = jl_max((T)1 / C::qmax(), jl_min((T)1 / C::qmin(), q / C::gamma()));
            hf = h / q;
        }
    }
    static DEGK_DEV T next_h_accept(T hf, T rem) { return jl_min(abs_(hf), abs_(rem)); }
};

// ------------------------------------------------------------------------------------------
template <class T, class Model, template <class, class> class MethodT, int W>
DEGK_DEV void ode_asolve4_body(const KArgs& a, unsigned char* smem_raw) {
    constexpr bool FAST = !DEGK_STRICT;
    constexpr bool PACKED = slots4_packed<T, W>();
    typedef Slots4<T, W, PACKED> PO;
    typedef typename PO::type V;
    typedef MethodT<V, Model> MethodV;       // stepping (packed when PACKED)
    typedef MethodT<T, Model> MethodS;       // scalar: deferred saves of the non-packed builds, constants
    typedef StepMath<T, MethodS::ORDER, Model::N, false> SM;
    constexpr int N = Model::N;
    constexpr int NPA = Model::NP > 0 ? Model::NP : 1;
    constexpr int QCAP = asolve4_qcap<T, N, W>();
    constexpr int QDEPTH = save_queue_depth<T, N, W>();             // records per lane of the lane-private save queue
    static_assert(QDEPTH > W, "a lane must be able to queue one record per slot and pass");
    constexpr bool HK = use_hk<MethodV>();
    // fast build: a slot that lands on tf gets a next step < dtmin from the exact remainder and so stops by itself;
    // the strict build follows the reference's formulas, where the slot has to be parked explicitly
    constexpr bool PARK_NATURAL = FAST;
    typedef SaveRec<T, N> Rec;
    constexpr int QBATCH = 32;                                   // records replayed at a time, one per lane
    constexpr u32 QSTRIDE = 32u * (u32)sizeof(Rec);              // bytes between two records of one lane

    // tolerances in the working precision, pinned in registers (left to itself the compiler re-converts the
    // Float64 kernel arguments in every pass: two quarter-rate F2F per pass)
    const T abstol = opaque_f((T)a.abstol), reltol = opaque_f((T)a.reltol);
    const bool has_saveat = a.saveat != nullptr;
    const int nsv = (int)opaque((u32)(has_saveat ? a.n_saveat : 0));
    const u32 lane = lane_id();
    const u32 lt_mask = (1u << lane) - 1u;                    // (service path only)
    const int warp_in_block = (int)(threadIdx.x >> 5);
    const int nwarps = (int)(blockDim.x >> 5);
    const u32 max_it = a.max_iters > 0x3fffffffLL ? 0x3fffffffu : (u32)a.max_iters;
    const T kInf = (T)__longlong_as_double(0x7ff0000000000000LL);   // +inf: "no further save point"
    const T dtmin = MethodS::dtmin();
    // A slot integrates while its step is not below dtmin.  Strict build: exactly the reference's `dt < dtmin` test, so a
    // NaN step (the controller of a trajectory that went non-finite) is still attempted: that attempt is accepted (NaN
    // > 1 is false), t becomes NaN, `while t < tf` ends and the end-point row / retcode come from the NaN state -- as
    // in the reference and the oracle.  Fast build: a NaN step is not live (its slots stop by themselves through
    // h < dtmin, which a NaN never satisfies).
    auto is_live = [&](T hh) { return FAST ? (hh >= dtmin) : !(hh < dtmin); };
    const T kDead = (T)-1;                   // h of a slot that is not integrating

    // fast controller constants in the L = log2(N * EEst^2) representation
    const double lgN = log2((double)N), lgGamma = log2(9.0 / 10.0);
    const T b1h = (T)(0.5 * 7.0 / (10.0 * MethodS::ORDER));
    const T b2h = (T)(0.5 * 2.0 / (5.0 * MethodS::ORDER));
    const T k0 = (T)((0.5 * 7.0 / (10.0 * MethodS::ORDER) - 0.5 * 2.0 / (5.0 * MethodS::ORDER)) * lgN + lgGamma);
    const T lqZero = (T)lgN;                                  // qold = 1 (the reject branch ignores qold)
    const T lqInitF = (T)(2.0 * log2(1.0e-4) + lgN);          // qoldinit = 1e-4
    const T exLo = (T)-2.321928094887362, exHi = (T)3.321928094887362;   // fac in [qmin, qmax] = [1/5, 10]
    const T lqInit = FAST ? lqInitF : SM::lq_init();

    // shared memory: [per-warp save queues][per-warp problem pools][saveat copy + 2 x inf]
    constexpr int PW = asolve4_pool_words<T, N, Model::NP>();    // u0, p, t0, tf, first save cursor (0: nothing to integrate)
    // save queue of this warp: record k of lane l at index l + 32 k; behind the 32 * QDEPTH records the flush directory
    Rec* queue = (Rec*)smem_raw + (size_t)warp_in_block * QCAP;
    unsigned short* qdir = (unsigned short*)(queue + 32 * QDEPTH);
    const u32 queue_saddr = (u32)__cvta_generic_to_shared(queue);
    const u32 q0 = queue_saddr + lane * (u32)sizeof(Rec);                   // this lane's first record
    // qaddr >= qtrig: fewer than W free records (pinned in a register: the compiler would re-derive it from %tid in every pass)
    const u32 qtrig = opaque(queue_saddr + (u32)(QDEPTH - W + 1) * QSTRIDE);
    u32 qaddr = opaque(q0);                  // this lane's next free record
    T* pool = (T*)(smem_raw + (size_t)nwarps * QCAP * sizeof(Rec)) + (size_t)warp_in_block * 32 * PW;
    T* sv_s = (T*)(smem_raw + (size_t)nwarps * QCAP * sizeof(Rec)) + (size_t)nwarps * 32 * PW;
    // saveat is staged in shared memory as a 1-based array with two +inf sentinels behind the last entry (the host
    // launches the first-generation kernel for grids longer than DEGK_SAVEAT_STAGE_MAX)
    for (int i = (int)threadIdx.x; i < nsv; i += (int)blockDim.x) sv_s[i] = ((const T*)a.saveat)[i];
    if (threadIdx.x < 2) sv_s[nsv + (int)threadIdx.x] = kInf;
    __syncthreads();
    const u32 sv_saddr = opaque((u32)__cvta_generic_to_shared(sv_s) - (u32)sizeof(T));   // 1-based
    auto save_time = [&](int c) -> T { return lds_(sv_saddr + (u32)c * (u32)sizeof(T), (T)0); };   // c <= nsv + 2
    auto cursor_of = [&](u32 addr) -> int { return (int)((addr - sv_saddr) / (u32)sizeof(T)); };

    // per-thread state: W trajectories ("slots")
    V u[N], unew[N], err[N], p[NPA];
    typename MethodV::Keep K;
    T t[W], h[W], tf[W], next_save[W], lq[W];
    u32 ca[W];                               // save cursor as the shared-memory address of its saveat entry
    int traj[W];
    u32 natt[W], nacc[W];
    u32 singm = 0;                           // bit s: W was singular
    DEGK_UNROLL for (int s = 0; s < W; ++s) {
        traj[s] = -1; ca[s] = sv_saddr + (u32)sizeof(T); natt[s] = 0; nacc[s] = 0;
        t[s] = (T)0; h[s] = kDead; tf[s] = (T)0; next_save[s] = kInf; lq[s] = lqInit;
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
    int pool_base = 0;                       // warp-uniform
    int pool_n = 0, pool_pos = 0;

    // Load up to 32 problems [base, base + n) into the pool, one per lane, and do everything that does not
    // depend on the integration with the full warp: row 0 (kernels.jl:116-126), the t0 pre-fill of ts
    // (lowerlevel_solve.jl:318 fill!; the rows saved later overwrite it -- same warp, after the __syncwarp),
    // and trajectories with nothing to integrate (empty span, non-finite data, dt0 that cannot start).
    auto load_pool = [&](i64 base, int n) {
        if ((int)lane < n) {
            const i64 claim = a.order ? (i64)a.order[base + lane] : base + lane;
            T us_[N], ps_[NPA], t0_, tf_;
            load_problem<T, Model>(a, claim, us_, ps_, t0_, tf_);
            int c1 = 1;
            bool init_ok = true;
            if (a.reserved & 2) {                            // DAE initialisation, kernels.jl:93-99 (tolerances of the solve)
                T ug[N];
                DEGK_UNROLL for (int c = 0; c < N; ++c) ug[c] = us_[c];
                init_ok = dae_initialize<T, Model>(us_, ps_, t0_, abstol, reltol);
                if (!init_ok) {                              // kernels.jl:143-150: store the initial values and bail out
                    store_u<T, N>(a, claim, 0, ug); store_t<T>(a, claim, 0, t0_);
                    const bool ends = !has_saveat && !a.save_everystep;
                    if (ends) { store_u<T, N>(a, claim, 1, ug); store_t<T>(a, claim, 1, t0_); }
                    if (a.ts != nullptr) for (i64 k = ends ? 2 : 1; k < a.n_rows; ++k) store_t<T>(a, claim, k, t0_);
                    if (a.retcode) a.retcode[claim] = RC_INIT_FAILURE;
                    if (a.naccept) a.naccept[claim] = 0;
                    if (a.nreject) a.nreject[claim] = 0;
                    if (has_saveat && a.nsaved) a.nsaved[claim] = 1;
                    ++tot_fail;
                }
            }
            if (!init_ok) {
                c1 = 0;
            } else if (has_saveat) {
                if (t0_ == save_time(1)) { c1 = 2; store_u<T, N>(a, claim, 0, us_); store_t<T>(a, claim, 0, t0_); }
                if (a.ts != nullptr) for (i64 k = c1 - 1; k < a.n_rows; ++k) store_t<T>(a, claim, k, t0_);
            } else {
                store_u<T, N>(a, claim, 0, us_);
                if (a.ts != nullptr) for (i64 k = 0; k < a.n_rows; ++k) store_t<T>(a, claim, k, t0_);
            }
            if (init_ok && !(t0_ < tf_)) {               // empty time span: nothing to integrate
                if (!has_saveat && !a.save_everystep) { store_u<T, N>(a, claim, 1, us_); store_t<T>(a, claim, 1, t0_); }
                if (a.retcode) a.retcode[claim] = RC_SUCCESS;
                if (a.naccept) a.naccept[claim] = 0;
                if (a.nreject) a.nreject[claim] = 0;
                if (has_saveat && a.nsaved) a.nsaved[claim] = c1 - 1;
                c1 = 0;
            }
            T* e = pool + lane * PW;
            DEGK_UNROLL for (int c = 0; c < N; ++c) e[c] = us_[c];
            DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) e[N + c] = ps_[c];
            e[N + Model::NP] = t0_; e[N + Model::NP + 1] = tf_;
            ((int*)e)[(N + Model::NP + 2) * (int)(sizeof(T) / sizeof(int))] = c1;
        }
        __syncwarp();
    };

    // start the pooled trajectory `ei` in slot s of this lane
    auto start_slot = [&](int s, int ei, u32& freshm) {
        const T* e = pool + ei * PW;
        const int c1 = ((const int*)e)[(N + Model::NP + 2) * (int)(sizeof(T) / sizeof(int))];
        if (c1 == 0) return;                              // finished at load time; the slot stays free
        DEGK_UNROLL for (int c = 0; c < N; ++c) u[c] = PO::set(u[c], s, e[c]);
        DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) p[c] = PO::set(p[c], s, e[N + c]);
        const T t0_ = e[N + Model::NP], tf_ = e[N + Model::NP + 1];
        t[s] = t0_; tf[s] = tf_;
        lq[s] = lqInit;
        natt[s] = 0; nacc[s] = 0;
        ca[s] = sv_saddr + (u32)c1 * (u32)sizeof(T);
        next_save[s] = save_time(c1);
        traj[s] = a.order ? a.order[pool_base + ei] : pool_base + ei;
        const T h0 = (T)a.dt;
        // dt0 < dtmin errors at the first attempt; non-finite time data cannot be integrated:
        // both park the slot (dead), the retire path derives the return code
        const bool valid = finite_(t0_) & finite_(tf_) & finite_(h0);
        h[s] = valid ? fmax_(h0, (T)0) : kDead;           // dt0 <= 0 fails like dt0 < dtmin
        if (h[s] >= dtmin) freshm |= (1u << s);
    };

    // Flush the lane-private save queue (warp-collective).  Full batches of 32 records are processed one per lane;
    // the < 32 left-over records are re-homed one per lane (which also levels the per-lane counts), unless `final`.
    // (A packed replay -- two records per lane through the packed stepper -- halves the issue cost per record but
    //  was measured slower: inlined it pushes loop-carried state of the attempt loop into local memory, as a
    //  separate function the call does; profiles/r2_c2_lanequeue.md.)
    auto flush_saves = [&](bool final) {
        __syncwarp();
        const int n = (int)((qaddr - q0) / QSTRIDE);
        int incl = n;
        DEGK_UNROLL for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - n;
        DEGK_UNROLL for (int k = 0; k < QDEPTH; ++k)
            if (k < n) qdir[excl + k] = (unsigned short)(lane + 32u * (u32)k);
        __syncwarp();
        int done = 0;
        while (total - done >= QBATCH || (final && done < total)) {
            const int cnt = total - done < QBATCH ? total - done : QBATCH;
            process_saves<T, Model, MethodS>(a, queue, qdir, done, cnt, sv_saddr);
            done += cnt;
        }
        const int left = total - done;       // < QBATCH
        Rec r[QBATCH / 32];
        DEGK_UNROLL for (int j = 0; j < QBATCH / 32; ++j)
            if ((int)lane + 32 * j < left) rec_copy(&r[j], queue + qdir[done + (int)lane + 32 * j]);
        __syncwarp();
        qaddr = q0;
        DEGK_UNROLL for (int j = 0; j < QBATCH / 32; ++j)
            if ((int)lane + 32 * j < left) { rec_copy(queue + lane + 32 * j, &r[j]); qaddr += QSTRIDE; }
        __syncwarp();
    };

    bool service = true;                     // warp-uniform: a slot needs attention (or start of the kernel)
    u32 iter = 0;                            // warp-uniform pass counter
    bool all_done = false;
    bool started = false;                    // warp-uniform: the first service pass (initial fill) ran
    for (;;) {
        if (service) {
            u32 freshm = 0;
            // some lane's part of the save queue is nearly full
            if (__any_sync(0xffffffffu, qaddr >= qtrig)) flush_saves(false);
            // maximum number of attempts reached: park the slot, the retire path reports MaxIters
            DEGK_UNROLL for (int s = 0; s < W; ++s)
                if (natt[s] >= max_it && is_live(h[s])) h[s] = kDead;
            for (;;) {
                // slot states: integrating (h >= dtmin) / stopped, waiting to retire / free
                u32 havem = 0, donem = 0;
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    const bool hv = is_live(h[s]);
                    havem |= (u32)hv << s;
                    donem |= (u32)(!hv & (traj[s] >= 0)) << s;
                }
                // several save points inside one accepted step: the queued record covers all of them
                // (the replay loops), skip the cursor past them
                // (after the push next_save is the save time behind the pushed one and t the end of the step,
                //  so `next_save <= t` after at least one accepted step identifies exactly those slots;
                //  before the first accepted step a save point may legitimately lie before t0 -- it
                //  is extrapolated from the first step like integrator_utils.jl:34-47 does)
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    if (next_save[s] <= t[s] && nacc[s] != 0u) {
                        do { ca[s] += (u32)sizeof(T); next_save[s] = lds_(ca[s], (T)0); } while (next_save[s] <= t[s]);   // ends at the +inf sentinel
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
                            const u32 natt_ = natt[s];
                            if (!(t[s] < tf[s])) {                   // reached tf -- or t is NaN (`while t < tf` ended)
                                T uf[N];
                                DEGK_UNROLL for (int c = 0; c < N; ++c) uf[c] = PO::get(u[c], s);
                                if (!has_saveat && !a.save_everystep) {  // kernels.jl:139-142
                                    store_u<T, N>(a, traj[s], 1, uf);
                                    store_t<T>(a, traj[s], 1, t[s]);
                                }
                                bool fin = true;
                                DEGK_UNROLL for (int c = 0; c < N; ++c) fin = fin && finite_(uf[c]);
                                if (!fin) rc = RC_UNSTABLE;
                            } else if ((singm >> s) & 1u) rc = RC_SINGULAR;
                            else if (natt_ >= max_it) rc = RC_MAXITERS;
                            else if (h[s] >= (T)0) rc = RC_DT_LESS_THAN_MIN;
                            else rc = RC_UNSTABLE;
                            if (has_saveat && a.nsaved) a.nsaved[traj[s]] = cursor_of(ca[s]) - 1;
                            if (a.retcode) a.retcode[traj[s]] = rc;
                            if (a.naccept) a.naccept[traj[s]] = (int)nacc[s];
                            if (a.nreject) a.nreject[traj[s]] = (int)(natt_ - nacc[s]);
                            tot_acc += nacc[s]; tot_rej += natt_ - nacc[s];
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
                                load_pool(base, n);
                                pool_base = (int)base; pool_n = n; pool_pos = 0;
                                if (n == 0) break;
                            }
                            const int avail = pool_n - pool_pos;
                            const int take = avail < cnt - served ? avail : cnt - served;
                            if (mine && rank >= served && rank < served + take) start_slot(s, pool_pos + rank - served, freshm);
                            pool_pos += take;
                            served += take;
                        }
                        __syncwarp();
                    } else if (!static_done) {
                        // static schedule: slot s of warp w owns trajectories (w * W + s) * 32 + lane
                        const i64 base = (warp_global * W + s) * 32;
                        const i64 left = a.n_traj - base;
                        const int n = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
                        load_pool(base, n);
                        pool_base = (int)base;
                        if ((int)lane < n) start_slot(s, (int)lane, freshm);
                        __syncwarp();
                    }
                }
                if (!queue_sched) { static_done = true; exhausted = true; pool_n = pool_pos = 0; }
                // anything integrating now?
                bool mine_live = false, mine_wait = false;
                DEGK_UNROLL for (int s = 0; s < W; ++s) {
                    const bool hv = is_live(h[s]);
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
        const bool solved = attempt4<HK, MethodV>(K, u, p, tv, hv, unew, err);

        // ---------------- error norm and step-size control ----------------
        bool rej[W];
        T hf_[W], lqa_[W], rem_[W], tsum_[W];
        if constexpr (FAST && sizeof(T) == 8) {
            // Float64 fast build: the error norm decides accept / reject and feeds the step-size factor -- a few digits
            // are enough, so the scaled errors are formed in Float32 with the MUFU reciprocal / log2 / exp2 (the
            // Float64 division, log2 and exp2 were ~190 of the ~650 instructions of a Vern9 attempt)
            static_assert(W == 1, "Float64 runs one trajectory per thread");
            float accf = 0.f;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                const float sc = (float)fma_(vmaxabs(u[c], unew[c]), reltol, abstol);
                const float v = (float)err[c] * rcp_(sc);
                accf = fmaf(v, v, accf);
            }
            const float Lf = log2_(accf);                            // log2(N * EEst^2)
            rej[0] = !(accf <= (float)N);                            // EEst > 1 -- or NaN: see the packed branch below
            const float lqe = rej[0] ? (float)lqZero : (float)lq[0];
            const float ex = vclamp(fmaf(-(float)b1h, Lf, fmaf((float)b2h, lqe, (float)k0)), (float)exLo, (float)exHi);
            hf_[0] = h[0] * (T)exp2_(ex);                            // dt * fac
            tsum_[0] = t[0] + h[0];
            rem_[0] = tf[0] - tsum_[0];                              // what is left after this step (exact when small)
            lqa_[0] = (T)fmaxf(Lf, (float)lqInit);
        } else if constexpr (FAST) {
            // tmp ./ (abstol .+ max.(abs.(uprev), abs.(u)) * reltol), sum of squares (ODE_DEFAULT_NORM), packed
            V accn;
            DEGK_UNROLL for (int c = 0; c < N; ++c) {
                const V sc = fma_(vmaxabs(u[c], unew[c]), V(reltol), V(abstol));
                const V v = err[c] * vrcp(sc);
                accn = (c == 0) ? v * v : fma_(v, v, accn);
            }
            const V L = vlog2(accn);                                 // log2(N * EEst^2)
            T lqe[W];
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                // EEst > 1.  A NaN estimate (an attempt that overflowed: a start step far beyond the span, say, where the
                // fused arithmetic of this build meets inf - inf sooner than the reference's) is REJECTED here and the
                // step shrinks by the largest factor (the clamp below maps NaN to its lower bound); the reference, and
                // the strict build, accept it (`NaN > 1` is false) and lose the trajectory
                rej[s] = !(PO::get(accn, s) <= (T)N);
                lqe[s] = rej[s] ? lqZero : lq[s];
            }
            const V ex = vclamp(fma_(V(-b1h), L, fma_(V(b2h), PO::make(lqe), V(k0))), exLo, exHi);
            const V hf = hv * vexp2(ex);                             // dt * fac
            const V tsum = tv + hv;
            const V rem = tfv - tsum;                                // what is left after this step (exact when small)
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                hf_[s] = PO::get(hf, s); rem_[s] = PO::get(rem, s); tsum_[s] = PO::get(tsum, s);
                lqa_[s] = fmax_(PO::get(L, s), lqInit);
            }
        } else {
            DEGK_UNROLL for (int s = 0; s < W; ++s) {
                T accn = (T)0;
                DEGK_UNROLL for (int c = 0; c < N; ++c) {
                    const T sq = SM::scaled_sq(PO::get(u[c], s), PO::get(unew[c], s), PO::get(err[c], s), abstol, reltol);
                    accn = (c == 0) ? sq : accn + sq;
                }
                SM::control(accn, lq[s], lqInit, h[s], rej[s], hf_[s], lqa_[s]);
                rem_[s] = tf[s] - t[s] - h[s];
                tsum_[s] = t[s] + h[s];
            }
        }

        // ---------------- per-slot flags ----------------
        bool push[W], acc_[W];
        T tnew_[W], hnext_[W];
        bool any_evt = false;
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            bool land;
            T hacc;
            if constexpr (FAST) {
                // rem = tf - (t + dt): landing (gpu_tsit5_perform_step.jl:155-156) leaves a next step
                // min(dt * fac, rem) < dtmin, i.e. the slot stops integrating by itself
                land = rem_[s] < dtmin;
                hacc = fmin_(hf_[s], rem_[s]);
            } else {
                // rem = tf - t - dt as the reference computes it; a step that cannot advance t (remaining span
                // below ulp(t)) lands too -- the reference would loop forever
                land = (rem_[s] < MethodS::land()) | ((tsum_[s] == t[s]) & (rem_[s] <= h[s]));
                hacc = SM::next_h_accept(hf_[s], rem_[s]);
            }
            const T tn = land ? tf[s] : tsum_[s];
            const T hn = rej[s] ? hf_[s] : hacc;
            const bool live = is_live(h[s]);                     // dead slots carry h < dtmin
            const bool ok = live & solved;                       // W factorised
            const bool accept = ok & !rej[s];
            inc_if(ok, natt[s]);
            inc_if(accept, nacc[s]);
            // (a step size below dtmin needs no test here: the slot is simply not live any more
            //  in the next iteration -- `dt < dtmin && error(...)` -- and retires as DtLessThanMin)
            push[s] = accept & (next_save[s] <= tn);
            acc_[s] = accept;
            tnew_[s] = tn;
            // a slot that reached tf just idles until the service path next looks (every DEGK4_SERVICE_PERIOD
            // passes); the attempt limit, a step across several save points (below) and a nearly full save
            // queue need the service path at once
            any_evt |= ok & (natt[s] >= max_it);
            bool park = false;
            if constexpr (!PARK_NATURAL) park = accept & !(tn < tf[s]);
            if (!MethodS::ALWAYS_SOLVED) {
                const bool sing = live & !solved;
                singm |= (u32)sing << s;
                park |= sing;
            }
            hnext_[s] = live ? (park ? kDead : hn) : h[s];
        }

        // ---------------- queue the deferred saves (branch-free, lane-private) ----------------
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            Rec r;
            r.traj = traj[s];
            r.cur = ca[s];
            r.tprev = t[s];
            r.h = h[s];
            r.tnew = tnew_[s];
            DEGK_UNROLL for (int c = 0; c < N; ++c) r.u[c] = PO::get(u[c], s);
            rec_store_if(push[s], qaddr, r);
            add_if<QSTRIDE>(push[s], qaddr);
            add_if<(u32)sizeof(T)>(push[s], ca[s]);
            lds_if(push[s], ca[s], next_save[s]);
            // several save points inside this step: the queued record covers all of them (the replay loops),
            // the service path moves the cursor past them before the next attempt
            any_evt |= push[s] & (next_save[s] <= tnew_[s]);
        }
        any_evt |= qaddr >= qtrig;

        // ---------------- state update (selects only) ----------------
        DEGK_UNROLL for (int s = 0; s < W; ++s) {
            h[s] = hnext_[s];
            lq[s] = acc_[s] ? lqa_[s] : lq[s];
            t[s] = acc_[s] ? tnew_[s] : t[s];
        }

        // ---------------- commit accepted steps ----------------
        DEGK_UNROLL for (int c = 0; c < N; ++c) assign_if(acc_, u[c], unew[c]);
        MethodV::accepted_if(K, acc_);

        ++iter;
        service = __any_sync(0xffffffffu, any_evt) | ((iter & (u32)(DEGK4_SERVICE_PERIOD - 1)) == 0u);
    }
    // the remaining deferred saves
    flush_saves(true);
    add_totals<T>(a, tot_acc, tot_rej, tot_fail);
}

}  // namespace degk
