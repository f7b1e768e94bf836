/* degk.h -- C ABI of libdegk: the B200-native engine behind DiffEqGPU.jl's EnsembleGPUKernel path.
 *
 * Drop-in boundary.  Each entry point states the reference interface it replaces;
 * paths are relative to the reference repository root.
 *
 *   degk_program_build   replaces the GPUCompiler/KernelAbstractions specialisation of
 *                        `ode_solve_kernel`/`ode_asolve_kernel`/`em_kernel`/`siea_kernel` on
 *                        (f, alg, eltype) that happens at first call of
 *                        src/ensemblegpukernel/lowerlevel_solve.jl:113 / :182-188 / :333
 *   degk_solve           replaces the kernel launches of `vectorized_solve` (ODE)
 *                        src/ensemblegpukernel/lowerlevel_solve.jl:119-123, `vectorized_solve`
 *                        (SDE) :194-197 and `vectorized_asolve` :339-343, i.e. the device code
 *                        in src/ensemblegpukernel/kernels.jl:1-152 and
 *                        perform_step/gpu_{em,siea}_perform_step.jl
 *   degk_solve_host      replaces `batch_solve_up_kernel` src/solve.jl:382-419 (H2D `adapt` of
 *                        probs :399-400, the solve :403-415, D2H `Array(us)`, `Array(ts)` :416-417)
 *
 * Conventions: every function returns a degk_status (0 = ok) and never throws or aborts;
 * degk_last_error() gives the message.  All buffers are caller owned.  degk_solve() takes
 * DEVICE pointers and only enqueues work on the caller's stream; degk_solve_host() takes HOST
 * pointers and returns when the results are in host memory.  No torch types appear here.
 */
#ifndef DEGK_H
#define DEGK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEGK_VERSION 100

#if defined(__GNUC__)
#define DEGK_API __attribute__((visibility("default")))
#else
#define DEGK_API
#endif

typedef struct degk_ctx degk_ctx;         /* one per (process, device) */
typedef struct degk_program degk_program; /* immutable after build: (model, alg, dtype, fp mode) */

typedef enum {
    DEGK_OK = 0,
    DEGK_ERR_INVALID = 1,     /* bad argument */
    DEGK_ERR_CUDA = 2,        /* CUDA runtime/driver error (no GPU, launch failure, ...) */
    DEGK_ERR_NVRTC = 3,       /* JIT compilation failed; log in degk_last_error */
    DEGK_ERR_UNSUPPORTED = 4, /* combination not available (e.g. SIEA with non-diagonal noise) */
    DEGK_ERR_NOMEM = 5
} degk_status;

typedef enum { DEGK_F32 = 0, DEGK_F64 = 1 } degk_dtype;

/* reference selector structs: src/ensemblegpukernel/gpukernel_algorithms.jl:30-266 */
typedef enum {
    DEGK_ALG_TSIT5 = 0,        /* GPUTsit5        */
    DEGK_ALG_VERN7 = 1,        /* GPUVern7        */
    DEGK_ALG_VERN9 = 2,        /* GPUVern9        */
    DEGK_ALG_ROSENBROCK23 = 3, /* GPURosenbrock23 */
    DEGK_ALG_RODAS4 = 4,       /* GPURodas4       */
    DEGK_ALG_RODAS5P = 5,      /* GPURodas5P      */
    DEGK_ALG_EM = 6,           /* GPUEM           */
    DEGK_ALG_SIEA = 7,         /* GPUSIEA         */
    DEGK_ALG_KVAERNO3 = 8,     /* GPUKvaerno3 (ESDIRK + Newton; saveat through the default Hermite interpolant,
                                  nonstiff/interpolants.jl:1-21) */
    DEGK_ALG_KVAERNO5 = 9      /* GPUKvaerno5     */
} degk_alg;

/* strict: un-fused stage arithmetic like Julia emits -> bit-parity with the reference's CPU
 * backend (as restated by oracle/); fast: FMA-contracted, MUFU step-size control. */
typedef enum { DEGK_FP_STRICT = 0, DEGK_FP_FAST = 1 } degk_fp_mode;

/* DEGK_LAYOUT_REF: the reference's (len x N) column-major arrays of SVector{n,T}
 *   (lowerlevel_solve.jl:81-83): us[(i*len + k)*n + c], ts[i*len + k].
 * DEGK_LAYOUT_SOA: us[(k*n + c)*N + i], ts[k*N + i]. */
typedef enum { DEGK_LAYOUT_REF = 0, DEGK_LAYOUT_SOA = 1 } degk_layout;

/* DEGK_SCHED_STATIC: thread i integrates trajectory i (the reference's ndrange = length(probs)).
 * DEGK_SCHED_QUEUE : persistent kernel + atomic work queue with per-lane refill (adaptive only). */
typedef enum { DEGK_SCHED_STATIC = 0, DEGK_SCHED_QUEUE = 1, DEGK_SCHED_AUTO = 2 } degk_schedule;

/* per-trajectory return codes written to degk_solve_args.retcode */
typedef enum {
    DEGK_RC_DEFAULT = 0,
    DEGK_RC_SUCCESS = 1,
    DEGK_RC_DT_LESS_THAN_MIN = 2, /* reference: device error("dt<dtmin"), gpu_tsit5_perform_step.jl:102 */
    DEGK_RC_UNSTABLE = 3,         /* non-finite state or step size */
    DEGK_RC_MAXITERS = 4,
    DEGK_RC_SINGULAR = 5,         /* reference: SingularException, linalg/lu.jl:41-44 */
    DEGK_RC_TERMINATED = 6,       /* terminate!(integrator) in a callback affect, integrator_utils.jl:52-66 */
    DEGK_RC_INIT_FAILURE = 7      /* DAE initialisation did not converge (dae_init; kernels.jl:63-70, 143-150) */
} degk_retcode;

/* DEGK_ENGINE_AUTO picks the persistent adaptive kernel (batched deferred saves) when the program has
 * one, and for fixed-dt runs the lock-step kernel whenever its preconditions hold (one tspan and dt for
 * the launch, every-step saves, explicit RK stepper; a state too large for the staged flush of the
 * reference layout keeps it only for launches that do not fill the GPU).  In the Float32 fast mode both
 * carry two trajectories per thread in FFMA2 register pairs for launches that fill the GPU and one per
 * thread below (degk_api.cu::launch); DEGK_ENGINE_V1 forces the first-generation kernels (one thread per
 * trajectory; kept for A/B measurements); DEGK_ENGINE_LOCKSTEP takes the lock-step kernel whenever
 * its preconditions hold, whatever the layout and size. */
typedef enum { DEGK_ENGINE_AUTO = 0, DEGK_ENGINE_V1 = 1, DEGK_ENGINE_LOCKSTEP = 2 } degk_engine;

typedef enum { DEGK_NOISE_NONE = 0, DEGK_NOISE_DIAGONAL = 1, DEGK_NOISE_GENERAL = 2 } degk_noise;

/* Model description.  Either `builtin` names a model compiled into the library
 * ("lorenz", "henon_heiles", "rober", "decay", "linear15", "gbm", "scalar_sde", "osc_t",
 * "gbm_nd"), or the *_src fields hold CUDA C++ function BODIES that are JIT-compiled by NVRTC
 * for sm_100a and inlined into the stepper kernels.  Inside a body the names
 *   T (scalar type), u[i], p[i], t   and the output  du[i] / J[i][j] / dT[i] / g[i] / G[i][j]
 * are in scope (0-based).  This is the lowering target for the Julia side's
 * Symbolics `build_function(..., target = CTarget())` output. */
typedef struct {
    const char* builtin;    /* or NULL */
    const char* rhs_src;    /* du[...] = f(u, p, t)            (required when builtin == NULL) */
    const char* jac_src;    /* J[i][j] = d f_i / d u_j         (optional: see jac_mode) */
    const char* tgrad_src;  /* dT[i] = d f_i / d t             (NULL => zero) */
    const char* noise_src;  /* diagonal: g[i]; general: G[i][j] (SDE only) */
    int32_t n_state, n_param, n_noise;
    int32_t noise_kind;     /* degk_noise */
    int32_t dtype;          /* degk_dtype */
    int32_t alg;            /* degk_alg */
    int32_t fp_mode;        /* degk_fp_mode */
    int32_t force_jit;      /* JIT-compile even when `builtin` exists ahead of time */
    /* Events (reference: tstops + GPUDiscreteCallback, callbacks.jl:1-36, integrator_utils.jl:69-150).
     * events != 0 builds the event-capable kernel pair (always through NVRTC) that honours
     * degk_solve_args.tstops and the callbacks below (all ODE solvers; not the SDE ones).
     * Callback c: cb_condition_src[c] is the BODY of `bool condition(u, p, t)` (must `return`),
     * cb_affect_src[c] the body of `affect!(integrator)`: it may assign u[i] and p[i] and call
     * terminate().  Callbacks run in order after every step; save_positions is (false, false)
     * like the reference requires. */
    int32_t events;
    int32_t n_callbacks;
    const char* const* cb_condition_src;
    const char* const* cb_affect_src;
    /* Stiff solvers, model without an analytic Jacobian body (reference nlsolve/type.jl:129-157):
     * 0 = the default: jac_src / the built-in's Jacobian when there is one, else forward-mode duals
     *     (ForwardDiff.jacobian, alg autodiff = true);
     * 1 = finite differences (autodiff = false: finite_diff_jac, alg_utils.jl:17-27);
     * 2 = forward-mode duals.  On a built-in model 1 and 2 ignore its analytic Jacobian. */
    int32_t jac_mode;
    int32_t reserved;
    /* constant mass matrix M of  M u' = f(u, p, t)  (ODEFunction(f; mass_matrix = M), src/utils.jl:42-57):
     * body assigning Mm[i][j] (zero-initialised), NULL = identity.  Implicit solvers only (GPURosenbrock23,
     * GPURodas4, GPURodas5P: perform_step reads f.mass_matrix; GPUKvaerno3/5: nlsolve does); u0 must be
     * consistent (no DAE initialisation). */
    const char* mass_src;
    /* Continuous callbacks (GPUContinuousCallback, callbacks.jl:38-124; same kernels as `events`).
     * cc_condition_src[c]: body RETURNING the value of condition(u, t, integrator) (a root function);
     * cc_affect_src[c] / cc_affect_neg_src[c]: bodies of affect! (sign - -> +) and affect_neg! (+ -> -),
     * a NULL entry = `nothing`; cc_rootfind[c]: 0 LeftRootFind, 1 RightRootFind, 2 NoRootFind;
     * cc_abstol / cc_repeat_nudge / cc_dtrelax: the callback's fields (defaults 10eps(Float32), 1//100, 1). */
    int32_t n_ccallbacks;
    int32_t reserved3;
    const char* const* cc_condition_src;
    const char* const* cc_affect_src;
    const char* const* cc_affect_neg_src;
    const int32_t* cc_rootfind;
    const double* cc_abstol;
    const double* cc_repeat_nudge;
    const double* cc_dtrelax;
} degk_model_desc;

typedef struct {
    int32_t n_state, n_param, n_noise, noise_kind, dtype, alg, fp_mode;
    int32_t is_jit;             /* 1 when produced by NVRTC */
    int32_t regs_fixed, regs_adaptive;     /* registers per thread of the two kernels (0 = n/a) */
    int32_t local_bytes_fixed, local_bytes_adaptive; /* local-memory (spill) bytes per thread */
    int32_t max_blocks_per_sm;  /* occupancy of the adaptive (or only) kernel at 256 threads */
    /* second-generation adaptive kernel (deferred saves; packed pairs when slots_per_thread2 == 2) */
    int32_t regs_adaptive2, local_bytes_adaptive2, slots_per_thread2, max_blocks_per_sm2;
    double jit_seconds;         /* NVRTC compile + module load time */
} degk_program_info;

/* Arguments of one batched solve.  `probs` of the reference (a Vector of ImmutableODEProblem,
 * each carrying u0, p, tspan) arrive as three strided arrays; a stride of 0 broadcasts one
 * value to every trajectory.  All arrays have element type `dtype` of the program. */
typedef struct {
    int64_t n_traj;        /* length(probs) */
    int64_t traj_offset;   /* global index of trajectory 0 (keeps RNG streams shard-invariant) */
    const void* u0;    int64_t u0_stride;    /* n_state values per trajectory */
    const void* p;     int64_t p_stride;     /* n_param values per trajectory (may be NULL if 0) */
    const void* tspan; int64_t tspan_stride; /* (t0, tf) */
    double dt;             /* fixed step, or initial step when adaptive (kw `dt`) */
    int32_t adaptive;      /* 0: vectorized_solve, 1: vectorized_asolve */
    double abstol, reltol; /* kw `abstol`, `reltol` (adaptive only) */
    const void* saveat;    /* kw `saveat` already converted to a vector of T, or NULL */
    int32_t n_saveat;
    int32_t save_everystep;/* kw `save_everystep` */
    int64_t n_rows;        /* `len`: rows of ts/us, computed by the caller exactly as
                              lowerlevel_solve.jl:71-109 / :311-324 do (see degk_output_rows) */
    void* us;              /* out: n_rows * n_traj * n_state values */
    void* ts;              /* out: n_rows * n_traj values; NULL to skip (saveat is shared) */
    int32_t out_layout;    /* degk_layout */
    int32_t schedule;      /* degk_schedule */
    int32_t* retcode;      /* optional out, per trajectory */
    int32_t* naccept;      /* optional out: accepted steps (the reference keeps no counters) */
    int32_t* nreject;      /* optional out: rejected attempts */
    uint64_t seed;         /* SDE: prob.seed */
    double* reduce;        /* optional inout [n_rows][n_state][2]: += sum(u), sum(u^2) over the
                              trajectories of this call (SDE kernels; needs tspan_stride == 0) */
    uint64_t* totals;      /* optional inout [4]: += accepted, rejected, failed, 0 */
    int64_t max_iters;     /* attempts per trajectory before DEGK_RC_MAXITERS; 0 => 1e7 for adaptive runs, no cap for fixed dt */
    int32_t engine;        /* degk_engine: which adaptive kernel generation to run */
    int32_t dae_init;      /* 1: before the first step, solve the algebraic equations of a mass-matrix DAE for consistent
                              initial values of the algebraic states (reference gpu_initialization_solve,
                              nlsolve/initialization.jl:1-54; trust-region Newton, tolerances 1e-6 for fixed dt, the
                              solve's abstol / reltol when adaptive).  Trajectories whose initialisation fails are not
                              integrated (DEGK_RC_INIT_FAILURE).  Not available with tstops / callbacks */
    const void* tstops;    /* kw `tstops`: n_tstops ascending times of the program's dtype (device pointer in
                              degk_solve, host pointer in degk_solve_host); needs a program built with
                              events != 0 */
    int32_t n_tstops;
    int32_t reserved2;
    int32_t* nsaved;       /* optional out, per trajectory, saveat runs of the adaptive kernels
                              (generation 2/3): number of leading rows written.  Rows k < nsaved
                              hold ts = saveat[k], the others keep t0 -- which lets a caller pass
                              ts = NULL and rebuild the reference's ts array from 4 bytes per
                              trajectory instead of transferring it (degk_solve_host does) */
    int64_t saveat_stride; /* elements between the saveat grids of consecutive trajectories; 0 = one grid for all.
                              Per-problem `saveat` (reference kernels.jl:15-17, 89-91, src/solve.jl:226-250): every
                              trajectory brings its own grid of the SAME length n_saveat; `saveat` then points at
                              n_traj * saveat_stride values */
    const int32_t* order;  /* optional (adaptive kernel): permutation of 0..n_traj-1; trajectories are STARTED in this
                              order (e.g. sorted by a parameter that predicts the step count, so that the lanes of a
                              warp finish together), results are written at each trajectory's own index.  Device
                              pointer in degk_solve; ignored by degk_solve_host, the fixed-dt and the SDE kernels */
} degk_solve_args;

DEGK_API int degk_version(void);

/* device < 0 => current device.  Fails with DEGK_ERR_CUDA when no GPU is present:
 * there is no CPU fallback in this library. */
DEGK_API int degk_ctx_create(int device, degk_ctx** out);
DEGK_API void degk_ctx_destroy(degk_ctx* ctx);
DEGK_API const char* degk_last_error(degk_ctx* ctx); /* ctx may be NULL: last error of failed ctx_create */

DEGK_API int degk_program_build(degk_ctx* ctx, const degk_model_desc* desc, degk_program** out);
DEGK_API int degk_program_get_info(const degk_program* prog, degk_program_info* info);
DEGK_API void degk_program_destroy(degk_program* prog);

/* Number of built-in (ahead-of-time) kernels and their names, for introspection/tests. */
DEGK_API int degk_builtin_count(void);
DEGK_API const char* degk_builtin_name(int i);

/* `len` of the outputs as the reference computes it on the host:
 *   saveat given            -> n_saveat
 *   !save_everystep         -> 2
 *   fixed,  save_everystep  -> length(t0:dt:tf)            (lowerlevel_solve.jl:64-73)
 *   adaptive, save_everystep-> ceil((tf-t0)/dt) + 1        (lowerlevel_solve.jl:312-313)
 * Range length follows Julia's float `:` (rational lifting with rounding fallback). */
DEGK_API int64_t degk_output_rows(int dtype, double t0, double tf, double dt, int adaptive,
                         int save_everystep, int n_saveat);

/* Enqueue one batched solve on `stream` (a cudaStream_t, NULL = default stream).
 * All pointers in `args` are device pointers.  Asynchronous. */
DEGK_API int degk_solve(degk_program* prog, const degk_solve_args* args, void* stream);

/* Same solve with HOST pointers: uploads u0/p/tspan/saveat, solves in chunks that are
 * double-buffered across streams, downloads us/ts(/retcode/naccept/nreject) and returns when
 * everything is in host memory.  `reduce`/`totals` are host arrays here too.
 * chunk_traj <= 0 picks a default. */
DEGK_API int degk_solve_host(degk_program* prog, const degk_solve_args* args, int64_t chunk_traj);

/* Raw RNG access for parity tests: fills out[4*n] with Philox4x32-10 blocks for counters
 * (c0 + i, c1, 0, 0), key (k0, k1), computed on the device. */
DEGK_API int degk_debug_philox(degk_ctx* ctx, uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1,
                      int64_t n, uint32_t* out_host);

/* Compile a model description with NVRTC without loading it (works without a GPU): returns
 * the status, the cubin size and the compiler log / error text. */
DEGK_API int degk_jit_compile_check(const degk_model_desc* desc, int64_t* cubin_bytes, char* msg,
                                    int64_t msg_cap);

#ifdef __cplusplus
}
#endif
#endif /* DEGK_H */
