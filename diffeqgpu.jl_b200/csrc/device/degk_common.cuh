// degk_common.cuh -- shared device-side definitions of the B200 ensemble ODE/SDE engine.
//
// This header (and everything under csrc/device/) is compiled two ways from the same text:
//   * ahead of time by nvcc into libdegk.so for the built-in models, and
//   * at run time by NVRTC (degk_jit.cpp) with a user model spliced in,
// so it must not include any host/standard header when __CUDACC_RTC__ is defined.
//
// Floating-point modes (one translation unit each, selected with -DDEGK_STRICT=0/1 together
// with nvcc/NVRTC --fmad=false/true):
//   strict: stage arithmetic is separate FMUL/FADD in the reference's association order,
//           IEEE div/sqrt, Float32 pow through FP64 log2/exp2  -> bit-parity with the oracle
//           (which restates what Julia emits on the reference's CPU backend).
//   fast:   nvcc contracts a*b+c to FFMA, the error-norm division and the controller powers
//           use MUFU approximations.  Step-size control differs in the last bits only.
#pragma once

#ifndef DEGK_STRICT
#define DEGK_STRICT 0
#endif

#define DEGK_DEV __device__ __forceinline__
#define DEGK_UNROLL _Pragma("unroll")

namespace degk {

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

// ---- return codes per trajectory (numbering shared with include/degk.h) ----
enum { RC_DEFAULT = 0, RC_SUCCESS = 1, RC_DT_LESS_THAN_MIN = 2, RC_UNSTABLE = 3, RC_MAXITERS = 4,
       RC_SINGULAR = 5 };
enum { LAYOUT_REF = 0, LAYOUT_SOA = 1 };
enum { SCHED_STATIC = 0, SCHED_QUEUE = 1 };

// Kernel argument block (plain data; mirrored field-for-field by degk_api.cu).
struct KArgs {
    i64 n_traj;          // trajectories in this launch (this shard)
    i64 traj_offset;     // global index of local trajectory 0 (RNG keys, multi-GPU shards)
    const void* u0;  i64 u0_stride;      // elements between trajectories, 0 = broadcast
    const void* p;   i64 p_stride;
    const void* tspan; i64 tspan_stride; // (t0, tf) pairs
    const void* saveat; int n_saveat;    // device array of T, or null
    int save_everystep;
    i64 n_rows;          // `len` of the reference's (len x N) outputs
    void* us; void* ts;  // ts may be null
    int out_layout; int schedule;
    int* retcode; int* naccept; int* nreject;   // optional, per trajectory
    double dt, abstol, reltol;
    u64 seed;
    double* reduce;      // optional [n_rows][N][2] : sum(u), sum(u^2) over trajectories
    u64* totals;         // optional [4]: accepted, rejected, failed, (unused)
    u64* work_counter;   // SCHED_QUEUE: next unclaimed trajectory
    i64 max_iters;
    int retire_batch;    // v2 adaptive kernel: retire/refill when this many slots of a warp have finished
    int reserved;
    int* nsaved;         // optional, per trajectory (saveat runs of the adaptive kernels): rows written
    const void* tstops;  // event-capable kernels (degk_ode_events.cuh): times the steppers must hit
    int n_tstops;
    int stage_rows;      // fixed-dt kernel: rows staged in shared memory per lane before a coalesced flush (0 = off)
    i64 saveat_stride;   // elements between the saveat grids of two trajectories; 0 = one grid shared by all
                         // (per-problem `saveat` of the reference, kernels.jl:15-17, 89-91: same length everywhere)
    const int* order;    // optional: the k-th trajectory the adaptive kernel starts is order[k] (a permutation of
                         // 0..n_traj-1, e.g. sorted by a parameter so that a warp's trajectories take similar numbers
                         // of steps); outputs stay at the trajectory's own index
};

// ---- fused multiply-add that stays fused in both fp modes (reference: @muladd / muladd) ----
// tableau coefficient `v`, entry `i` of the method's __constant__ table `tab` (gen_erk_*.cuh).  Default: the literal for
// every element type.  With DEGK_COEF_TABLE 1 the Float64 instantiations read the table instead: a 64-bit literal costs
// two UMOVs each time it is used (266 UMOV next to 235 DFMA in one Vern9 attempt), a table entry half an LDCU.128 --
// 37 % fewer SASS instructions in the Vern9 Float64 kernels, but the FP64 pipe, not the issue slots, bounds them:
// measured C3 Henon-Heiles +6.6 %, C3 Lorenz -2 %, lock-step Float64 +2 %, one-thread-per-trajectory Vern9 Float64
// -12 %, so it stays off.
#ifndef DEGK_COEF_TABLE
#define DEGK_COEF_TABLE 0
#endif
template <class T> DEGK_DEV T coef_(const double* tab, int i, double v) { (void)tab; (void)i; return (T)v; }
#if DEGK_COEF_TABLE
template <> DEGK_DEV double coef_<double>(const double* tab, int i, double v) { (void)v; return tab[i]; }
#endif
DEGK_DEV float  fma_(float a, float b, float c)    { return fmaf(a, b, c); }
DEGK_DEV double fma_(double a, double b, double c) { return fma(a, b, c); }

// Julia max/min propagate NaN; CUDA fmaxf/fminf return the non-NaN operand.
template <class T> DEGK_DEV T jl_max(T a, T b) {
#if DEGK_STRICT
    return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
#else
    return a > b ? a : b;
#endif
}
template <class T> DEGK_DEV T jl_min(T a, T b) {
#if DEGK_STRICT
    return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
#else
    return a < b ? a : b;
#endif
}
#if DEGK_STRICT
// Float32: the NaN-propagating forms are single instructions (FMNMX.NAN) instead of three compares and three
// selects.  Same value as the generic form for every operand pair the steppers produce (the operands are
// magnitudes or positive step-size factors, so the +0 / -0 ordering of max.NaN never matters); a NaN result may
// carry the other operand's payload, which no output ever shows.
template <> DEGK_DEV float jl_max<float>(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
template <> DEGK_DEV float jl_min<float>(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
#endif
DEGK_DEV float  abs_(float x)  { return fabsf(x); }
DEGK_DEV double abs_(double x) { return fabs(x); }
DEGK_DEV float  sqrt_(float x)  { return sqrtf(x); }     // IEEE (-prec-sqrt=true default)
DEGK_DEV double sqrt_(double x) { return sqrt(x); }

// x^y for the PI controller (integrator_utils.jl:1-11 exponents; gpu_tsit5_perform_step.jl:127).
// strict/Float32: Julia's Base `^` = Float32(exp2(log2(Float64 x) * Float64 y)).
DEGK_DEV float pow_(float x, float y) {
#if DEGK_STRICT
    if (x == 1.0f) return 1.0f;
    return (float)exp2(log2((double)x) * (double)y);
#else
    return exp2f(y * __log2f(x));    // MUFU.LG2 + FMUL + MUFU.EX2
#endif
}
DEGK_DEV double pow_(double x, double y) { return pow(x, y); }

// MUFU log2 / exp2 for the fast-mode controller (log-domain PI controller)
DEGK_DEV float log2_(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DEGK_DEV float exp2_(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DEGK_DEV double log2_(double x) { return log2(x); }
DEGK_DEV double exp2_(double x) { return exp2(x); }

DEGK_DEV float rcp_(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DEGK_DEV double rcp_(double x) { return 1.0 / x; }
DEGK_DEV float fmax_(float a, float b) { return fmaxf(a, b); }
DEGK_DEV float fmin_(float a, float b) { return fminf(a, b); }
DEGK_DEV double fmax_(double a, double b) { return fmax(a, b); }
DEGK_DEV double fmin_(double a, double b) { return fmin(a, b); }

// division used only for step-size control quantities (error scaling, q factors)
DEGK_DEV float ctl_div(float a, float b) {
#if DEGK_STRICT
    return a / b;
#else
    return __fdividef(a, b);
#endif
}
DEGK_DEV double ctl_div(double a, double b) { return a / b; }

// sum/length of ODE_DEFAULT_NORM
template <class T, int N> DEGK_DEV T mean_(T acc) {
#if DEGK_STRICT
    return acc / (T)N;
#else
    return acc * (T)(1.0 / N);
#endif
}

// ---- output stores ------------------------------------------------------------------
// REF layout = the reference's (len x N) column-major arrays of SVector{n,T}
// (lowerlevel_solve.jl:81-83): us[(i*len + k)*n + c], ts[i*len + k].
// SOA layout: us[(k*n + c)*N + i], ts[k*N + i]  (coalesced when lanes save in lock step).
template <class T, int N>
DEGK_DEV void store_u(const KArgs& a, i64 traj, i64 k, const T (&u)[N]) {
    if (k < 0 || k >= a.n_rows) return;   // the reference would write out of bounds here
    T* us = (T*)a.us;
    if (a.out_layout == LAYOUT_REF) {
        T* d = us + (traj * a.n_rows + k) * N;
        DEGK_UNROLL for (int c = 0; c < N; ++c) d[c] = u[c];
    } else {
        DEGK_UNROLL for (int c = 0; c < N; ++c) us[(k * N + c) * a.n_traj + traj] = u[c];
    }
}
template <class T>
DEGK_DEV void store_t(const KArgs& a, i64 traj, i64 k, T t) {
    if (a.ts == nullptr || k < 0 || k >= a.n_rows) return;
    T* ts = (T*)a.ts;
    if (a.out_layout == LAYOUT_REF) ts[traj * a.n_rows + k] = t;
    else ts[k * a.n_traj + traj] = t;
}

// PI controller constants, integrator_utils.jl:1-11 (computed in Float64, converted to T)
template <class T, int ORDER>
struct Ctl {
    static DEGK_DEV T beta1()    { return (T)(7.0 / (10.0 * ORDER)); }
    static DEGK_DEV T beta2()    { return (T)(2.0 / (5.0 * ORDER)); }
    static DEGK_DEV T qmax()     { return (T)10.0; }
    static DEGK_DEV T qmin()     { return (T)(1.0 / 5.0); }
    static DEGK_DEV T gamma()    { return (T)(9.0 / 10.0); }
    static DEGK_DEV T qoldinit() { return (T)1.0e-4; }
};

// stage hook of the generated explicit RK steppers (see gen_erk_*.cuh::attempt): nothing by default
struct NoHook { template <int J, int NH> DEGK_DEV void at() {} };

DEGK_DEV u32 lane_id() { u32 r; asm volatile("mov.u32 %0, %%laneid;" : "=r"(r)); return r; }

template <class T> DEGK_DEV bool finite_(T x) { return (x - x) == (T)0; }

// per-slot select used by the packed kernel (W = 1: plain scalar select)
// m: bit s set => take `a` for slot s
template <class V> DEGK_DEV V blendm(unsigned m, V a, V b) { return (m & 1u) ? a : b; }

// load u0 / p / tspan of one trajectory (AoS inputs like the reference's probs[i])
template <class T, class Model>
DEGK_DEV void load_problem(const KArgs& a, i64 traj, T (&u)[Model::N], T* p, T& t0, T& tf) {
    const T* u0 = (const T*)a.u0 + traj * a.u0_stride;
    DEGK_UNROLL for (int c = 0; c < Model::N; ++c) u[c] = u0[c];
    if (Model::NP > 0) {
        const T* pp = (const T*)a.p + traj * a.p_stride;
        DEGK_UNROLL for (int c = 0; c < Model::NP; ++c) p[c] = pp[c];
    }
    const T* ts = (const T*)a.tspan + traj * a.tspan_stride;
    t0 = ts[0];
    tf = ts[1];
}

// unwritten ts slots keep tspan[1] (lowerlevel_solve.jl:82 `fill!(ts, prob.tspan[1])`): the host
// infers early termination from them (src/solve.jl:260-277).  Instead of a separate fill pass
// over ts we write t0 into the rows this trajectory did not reach.
template <class T>
DEGK_DEV void fill_unwritten_ts(const KArgs& a, i64 traj, i64 first, T t0) {
    if (a.ts == nullptr) return;
    for (i64 k = first; k < a.n_rows; ++k) store_t<T>(a, traj, k, t0);
}

template <class T>
DEGK_DEV void add_totals(const KArgs& a, u32 nacc, u32 nrej, u32 nfail) {
    if (a.totals == nullptr) return;
    DEGK_UNROLL for (int o = 16; o > 0; o >>= 1) {
        nacc += __shfl_xor_sync(0xffffffffu, nacc, o);
        nrej += __shfl_xor_sync(0xffffffffu, nrej, o);
        nfail += __shfl_xor_sync(0xffffffffu, nfail, o);
    }
    if (lane_id() == 0) {
        atomicAdd(a.totals + 0, (u64)nacc);
        atomicAdd(a.totals + 1, (u64)nrej);
        if (nfail) atomicAdd(a.totals + 2, (u64)nfail);
    }
}

}  // namespace degk
