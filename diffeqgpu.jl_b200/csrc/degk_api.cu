// degk_api.cu -- host side of libdegk: contexts, programs, launch logic, the host-buffer
// (end-to-end) solve, and small host helpers.  See include/degk.h for the contract and the
// reference interfaces each entry point replaces.
//
// There is deliberately no CPU execution path in this library: every solve is a CUDA kernel
// launch and degk_ctx_create fails when no device is present.
#include <cuda_runtime.h>
#include <cuda.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/degk.h"
#include "degk_internal.h"
#include "degk_host.h"
#include "device/degk_common.cuh"
#include "device/degk_sde_kernels.cuh"

// ---- AOT tables (one per fp mode x model group, see degk_aot.cu) ----
#define DECL_TABLE(S) extern "C" const degk_aot_entry* degk_aot_table_##S(int* n);
DECL_TABLE(fast_0) DECL_TABLE(fast_1) DECL_TABLE(fast_2) DECL_TABLE(fast_3) DECL_TABLE(fast_4)
DECL_TABLE(strict_0) DECL_TABLE(strict_1) DECL_TABLE(strict_2) DECL_TABLE(strict_3) DECL_TABLE(strict_4)

typedef const degk_aot_entry* (*table_fn)(int*);
static const table_fn g_tables[2][5] = {
    {degk_aot_table_strict_0, degk_aot_table_strict_1, degk_aot_table_strict_2,
     degk_aot_table_strict_3, degk_aot_table_strict_4},
    {degk_aot_table_fast_0, degk_aot_table_fast_1, degk_aot_table_fast_2, degk_aot_table_fast_3,
     degk_aot_table_fast_4}};

static thread_local std::string g_create_error;

void degk_set_error(degk_ctx* ctx, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) {
        std::lock_guard<std::mutex> g(ctx->mu);
        ctx->err = buf;
    } else {
        g_create_error = buf;
    }
}

#define CK(ctx, call)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            degk_set_error(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                           __LINE__);                                                           \
            return DEGK_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

static size_t dtype_size(int dtype) { return dtype == DEGK_F64 ? 8 : 4; }

// dynamic shared memory of the adaptive kernel (degk_ode_kernels4.cuh, asolve4_smem_bytes):
// per-warp save queues + per-warp problem pools (32 x (n + np + 3) values) + saveat copy
size_t degk_smem2_bytes(const degk_program* prog, int n_saveat_staged, int qcap) {
    if (qcap <= 0) qcap = prog->qcap2;
    const size_t es = dtype_size(prog->info.dtype);
    const size_t nw = DEGK_BLOCK2 / 32;
    return nw * qcap * prog->rec_bytes2 + nw * 32 * (size_t)(prog->info.n_state + prog->info.n_param + 3) * es +
           ((size_t)n_saveat_staged + 2) * es;     // + two +inf sentinels
}

// dynamic shared memory of the lock-step kernel in the reference layout (degk_ode_lockstep.cuh, lockstep_smem_bytes):
// per warp 32 w buffers of 32 / w rows (rounded up to a multiple of four values, plus four) and the ring of save times
// largest staging area the lock-step kernel may ask for (bytes per block)
size_t degk_lockstep_smem_max() {
    static const size_t v = [] { const char* e = getenv("DEGK_LOCKSTEP_SMEM_MAX"); return e ? (size_t)atoll(e) : (size_t)DEGK_LOCKSTEP_SMEM_MAX; }();
    return v;
}
size_t degk_lockstep_smem_bytes(const degk_program* prog, int w) {
    const size_t es = dtype_size(prog->info.dtype);
    const int rows = 32 / std::max(1, w);                         // lockstep_ring_rows
    return (size_t)(DEGK_BLOCK2 / 32) * ((size_t)32 * w * (size_t)(((prog->info.n_state * rows + 3) & ~3) + 4) + (size_t)rows) * es;
}

extern "C" int degk_version(void) { return DEGK_VERSION; }

extern "C" const char* degk_last_error(degk_ctx* ctx) {
    if (!ctx) return g_create_error.c_str();
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->err_copy = ctx->err;
    return ctx->err_copy.c_str();
}

extern "C" int degk_ctx_create(int device, degk_ctx** out) {
    if (!out) return DEGK_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        degk_set_error(nullptr, "no CUDA device available (%s); libdegk has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return DEGK_ERR_CUDA;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= ndev) {
        degk_set_error(nullptr, "device %d out of range (%d devices)", device, ndev);
        return DEGK_ERR_INVALID;
    }
    degk_ctx* ctx = new degk_ctx();
    ctx->device = device;
    CK(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(nullptr, cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    CK(nullptr, cudaMalloc(&ctx->d_counters, sizeof(unsigned long long) * DEGK_NCOUNTERS));
    CK(nullptr, cudaMemset(ctx->d_counters, 0, sizeof(unsigned long long) * DEGK_NCOUNTERS));
    for (int i = 0; i < DEGK_NSTREAMS; ++i)
        CK(nullptr, cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking));
    *out = ctx;
    return DEGK_OK;
}

extern "C" void degk_ctx_destroy(degk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < DEGK_NSTREAMS; ++i) {
        if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
        for (auto& w : ctx->work[i].bufs)
            if (w.ptr) cudaFree(w.ptr);
    }
    if (ctx->h_nsaved) cudaFreeHost(ctx->h_nsaved);
    for (auto e : ctx->chunk_done) cudaEventDestroy(e);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_saveat) cudaFree(ctx->d_saveat);
    delete ctx;
}

// ------------------------------------------------------------------------------------------
static int g_builtin_n = -1;
static std::vector<degk_aot_entry> g_builtin;   // fast-mode table, used for names only
static void load_builtin() {
    if (g_builtin_n >= 0) return;
    for (int g = 0; g < 5; ++g) {
        int n = 0;
        const degk_aot_entry* t = g_tables[1][g](&n);
        for (int i = 0; i < n; ++i) g_builtin.push_back(t[i]);
    }
    g_builtin_n = (int)g_builtin.size();
}
extern "C" int degk_builtin_count(void) { load_builtin(); return g_builtin_n; }
extern "C" const char* degk_builtin_name(int i) {
    load_builtin();
    if (i < 0 || i >= g_builtin_n) return nullptr;
    static thread_local char buf[128];
    const degk_aot_entry& e = g_builtin[i];
    snprintf(buf, sizeof buf, "%s/alg%d/%s/%s", e.model, e.alg, e.dtype ? "f64" : "f32",
             e.adaptive ? "adaptive" : "fixed");
    return buf;
}

static const degk_aot_entry* find_aot(int fp_mode, const char* model, int alg, int dtype,
                                      int adaptive) {
    for (int g = 0; g < 5; ++g) {
        int n = 0;
        const degk_aot_entry* t = g_tables[fp_mode ? 1 : 0][g](&n);
        for (int i = 0; i < n; ++i)
            if (t[i].alg == alg && t[i].dtype == dtype && t[i].adaptive == adaptive &&
                strcmp(t[i].model, model) == 0)
                return &t[i];
    }
    return nullptr;
}

extern "C" int degk_program_build(degk_ctx* ctx, const degk_model_desc* d, degk_program** out) {
    if (!ctx || !d || !out) return DEGK_ERR_INVALID;
    *out = nullptr;
    if (d->alg < DEGK_ALG_TSIT5 || d->alg > DEGK_ALG_KVAERNO5 || (d->dtype != DEGK_F32 && d->dtype != DEGK_F64) ||
        (d->fp_mode != DEGK_FP_STRICT && d->fp_mode != DEGK_FP_FAST)) {
        degk_set_error(ctx, "invalid alg/dtype/fp_mode in model description");
        return DEGK_ERR_INVALID;
    }
    CK(ctx, cudaSetDevice(ctx->device));
    const bool is_sde = d->alg == DEGK_ALG_EM || d->alg == DEGK_ALG_SIEA;
    degk_program* prog = new degk_program();
    prog->ctx = ctx;
    prog->is_sde = is_sde;
    memset(&prog->info, 0, sizeof prog->info);
    prog->info.dtype = d->dtype; prog->info.alg = d->alg; prog->info.fp_mode = d->fp_mode;

    bool use_aot = d->builtin && !d->force_jit && !d->events && d->n_callbacks == 0 && d->n_ccallbacks == 0 &&   // event kernels are JIT-built
                   d->jac_mode == 0;
    if (use_aot) {
        const degk_aot_entry* e0 = find_aot(d->fp_mode, d->builtin, d->alg, d->dtype, 0);
        const degk_aot_entry* e1 = is_sde ? nullptr : find_aot(d->fp_mode, d->builtin, d->alg, d->dtype, 1);
        if (!e0 || (!is_sde && !e1)) {
            use_aot = false;          // not every (model, solver) pair is compiled ahead of time: NVRTC builds
                                      // the built-in's struct from the embedded degk_models.cuh instead
        } else {
            prog->fn[0] = e0->fn;
            prog->fn[1] = e1 ? e1->fn : nullptr;
            if (e0->fn3) { prog->fn[3] = e0->fn3; prog->w3 = e0->w3; prog->fn[4] = e0->fn4; }
            if (e1 && e1->fn2) {
                prog->fn[2] = e1->fn2;
                prog->w2 = e1->w2; prog->qcap2 = e1->qcap2; prog->rec_bytes2 = e1->rec_bytes2;
                prog->fn[5] = e1->fn2b; prog->qcap2b = e1->qcap2b;
            }
            prog->info.n_state = e0->n_state; prog->info.n_param = e0->n_param;
            prog->info.n_noise = e0->n_noise; prog->info.noise_kind = e0->noise_kind;
            for (int k = 0; k < 2; ++k) {
                if (!prog->fn[k]) continue;
                cudaFuncAttributes fa;
                cudaError_t e = cudaFuncGetAttributes(&fa, prog->fn[k]);
                if (e != cudaSuccess) {
                    degk_set_error(ctx, "cudaFuncGetAttributes failed: %s -- the library holds sm_100a code "
                                        "only; device is sm_%d%d", cudaGetErrorString(e), ctx->cc_major, ctx->cc_minor);
                    delete prog;
                    return DEGK_ERR_CUDA;
                }
                (k == 0 ? prog->info.regs_fixed : prog->info.regs_adaptive) = fa.numRegs;
                (k == 0 ? prog->info.local_bytes_fixed : prog->info.local_bytes_adaptive) = (int)fa.localSizeBytes;
            }
            int occ = 0;
            const void* fo = prog->fn[1] ? prog->fn[1] : prog->fn[0];
            CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fo, DEGK_BLOCK, 0));
            prog->info.max_blocks_per_sm = occ;
            for (int k = 3; k <= 4; ++k) {
                if (!prog->fn[k]) continue;
                const size_t ls_smem = degk_lockstep_smem_bytes(prog, k == 3 ? prog->w3 : 1);
                if (ls_smem > 48 * 1024 && ls_smem <= degk_lockstep_smem_max())
                    CK(ctx, cudaFuncSetAttribute(prog->fn[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ls_smem));
            }
            if (prog->fn[2]) {
                cudaFuncAttributes fa;
                CK(ctx, cudaFuncGetAttributes(&fa, prog->fn[2]));
                prog->info.regs_adaptive2 = fa.numRegs;
                prog->info.local_bytes_adaptive2 = (int)fa.localSizeBytes;
                prog->info.slots_per_thread2 = prog->w2;
                const size_t smem = degk_smem2_bytes(prog, 1024), smem_max = degk_smem2_bytes(prog, DEGK_SAVEAT_STAGE_MAX);
                if (smem_max > 48 * 1024) CK(ctx, cudaFuncSetAttribute(prog->fn[2], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
                CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, prog->fn[2], DEGK_BLOCK2, smem));
                prog->info.max_blocks_per_sm2 = occ;
                if (prog->fn[5]) {
                    const size_t smem_b = degk_smem2_bytes(prog, 1024, prog->qcap2b), smem_b_max = degk_smem2_bytes(prog, DEGK_SAVEAT_STAGE_MAX, prog->qcap2b);
                    if (smem_b_max > 48 * 1024) CK(ctx, cudaFuncSetAttribute(prog->fn[5], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b_max));
                    CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, prog->fn[5], DEGK_BLOCK2, smem_b));
                    prog->max_blocks_per_sm2b = occ;
                }
            }
        }
    }
    if (!use_aot) {
        int rc = degk_jit_build(ctx, d, prog);
        if (rc != DEGK_OK) { delete prog; return rc; }
    }
    *out = prog;
    return DEGK_OK;
}

extern "C" int degk_program_get_info(const degk_program* prog, degk_program_info* info) {
    if (!prog || !info) return DEGK_ERR_INVALID;
    *info = prog->info;
    return DEGK_OK;
}

extern "C" void degk_program_destroy(degk_program* prog) {
    if (!prog) return;
    degk_jit_release(prog);
    delete prog;
}

// ------------------------------------------------------------------------------------------
// Julia's `length(t0:dt:tf)` for IEEE floats (base/twiceprecision.jl `(:)(start, step, stop)`):
// lift the three numbers to exact rationals when they have short continued fractions, else
// fall back to round((stop-start)/step)+1 with an overshoot correction.
namespace {
template <class T>
bool rat(T x, long long& num, long long& den) {
    // Base.rat: continued fraction until it reproduces x (or the terms leave the exact range)
    T y = x;
    long long a = 1, d = 1, b = 0, c = 0;
    const T m = (sizeof(T) == 4) ? (T)16777216.0 : (T)9007199254740992.0;   // maxintfloat
    // narrow(): Julia uses the next narrower float's maxintfloat for the bound on the terms
    const double mn = (sizeof(T) == 4) ? 2048.0 : 16777216.0;
    (void)m;
    while (std::fabs((double)y) <= mn) {
        long long f = (long long)std::trunc((double)y);
        y -= (T)f;
        long long na = f * a + c, nb = f * b + d;
        c = a; d = b; a = na; b = nb;
        if (std::max(std::llabs(a), std::llabs(b)) > (long long)mn) { num = c; den = d; return true; }
        if ((T)a / (T)b == x) break;
        y = (T)1 / y;
    }
    num = a; den = b;
    return true;
}
long long gcdll(long long a, long long b) { a = std::llabs(a); b = std::llabs(b); while (b) { long long t = a % b; a = b; b = t; } return a; }

template <class T>
long long range_length(T start, T step, T stop) {
    if (step == (T)0) return 0;
    long long step_n, step_d;
    rat<T>(step, step_n, step_d);
    if (step_d != 0 && (T)((T)step_n / (T)step_d) == step) {
        long long start_n, start_d, stop_n, stop_d;
        rat<T>(start, start_n, start_d);
        rat<T>(stop, stop_n, stop_d);
        if (start_d != 0 && stop_d != 0 && (T)((T)start_n / (T)start_d) == start &&
            (T)((T)stop_n / (T)stop_d) == stop) {
            long long den = start_d / gcdll(start_d, step_d) * step_d;   // lcm
            const double m = (sizeof(T) == 4) ? 16777216.0 : 9007199254740992.0;
            if (den != 0 && std::fabs((double)start * den) <= m && std::fabs((double)step * den) <= m &&
                den % start_d == 0 && den % step_d == 0) {
                long long sn = (long long)std::llround((double)start * den);
                long long pn = (long long)std::llround((double)step * den);
                long long q = (den * stop_n - stop_d * sn) / (pn * stop_d);   // Julia div: truncation
                long long len = std::max(0LL, q + 1);
                // isbetween checks
                auto between = [](T a, T x, T b) { return (a <= x && x <= b) || (a >= x && x >= b); };
                T last = start + (T)(len - 1) * step;
                T next = start + (T)len * step;
                if (between(start, last, stop + step / (T)2) && !between(start, next, stop)) return len;
            }
        }
    }
    T lf = (stop - start) / step;
    if (lf < 0) return 0;
    if (lf == 0) return 1;
    long long len = (long long)std::llround((double)lf) + 1;   // round-half-even vs llround: ties are measure zero
    T stop2 = start + (T)(len - 1) * step;
    len -= ((start < stop && stop < stop2) ? 1 : 0) + ((start > stop && stop > stop2) ? 1 : 0);
    return len;
}
}  // namespace

extern "C" int64_t degk_output_rows(int dtype, double t0, double tf, double dt, int adaptive,
                                    int save_everystep, int n_saveat) {
    if (n_saveat > 0) return n_saveat;
    if (!save_everystep) return 2;
    if (adaptive) {   // ceil(Int, (tf - t0) / dt) + 1 in T arithmetic
        if (dtype == DEGK_F32) return (int64_t)std::ceil(((float)tf - (float)t0) / (float)dt) + 1;
        return (int64_t)std::ceil((tf - t0) / dt) + 1;
    }
    if (dtype == DEGK_F32) return range_length<float>((float)t0, (float)dt, (float)tf);
    return range_length<double>(t0, dt, tf);
}

// ------------------------------------------------------------------------------------------
static int validate(degk_program* prog, const degk_solve_args* a) {
    degk_ctx* ctx = prog->ctx;
    if (a->n_traj < 0 || a->n_rows <= 0) { degk_set_error(ctx, "n_traj/n_rows invalid"); return DEGK_ERR_INVALID; }
    if (a->n_traj == 0) return DEGK_OK;      // empty batch: nothing to check, nothing to launch
    if (!a->u0 || !a->tspan || !a->us) { degk_set_error(ctx, "u0, tspan and us are required"); return DEGK_ERR_INVALID; }
    if (prog->info.n_param > 0 && !a->p) { degk_set_error(ctx, "model has %d parameters but p is NULL", prog->info.n_param); return DEGK_ERR_INVALID; }
    if (prog->is_sde && a->adaptive) {
        // lowerlevel_solve.jl:348-356
        degk_set_error(ctx, "Adaptive time-stepping is not supported yet with GPUEM.");
        return DEGK_ERR_UNSUPPORTED;
    }
    if (!(a->dt > 0) && !(a->dt < 0)) { degk_set_error(ctx, "dt must be non-zero"); return DEGK_ERR_INVALID; }
    if (a->saveat && a->n_saveat <= 0) { degk_set_error(ctx, "saveat given with n_saveat <= 0"); return DEGK_ERR_INVALID; }
    if (a->saveat && a->n_rows != a->n_saveat) { degk_set_error(ctx, "n_rows must equal n_saveat when saveat is given"); return DEGK_ERR_INVALID; }
    if (a->saveat && a->saveat_stride != 0 && a->saveat_stride < a->n_saveat) { degk_set_error(ctx, "saveat_stride must be 0 or >= n_saveat"); return DEGK_ERR_INVALID; }
    if (a->reduce && (!prog->is_sde || a->tspan_stride != 0)) {
        degk_set_error(ctx, "reduce needs an SDE program and a broadcast tspan (tspan_stride == 0)");
        return DEGK_ERR_UNSUPPORTED;
    }
    if (a->adaptive && a->n_traj > 2147483000LL) {
        degk_set_error(ctx, "adaptive launches take at most 2^31 - 648 trajectories; split the batch");
        return DEGK_ERR_INVALID;
    }
    if (a->out_layout != DEGK_LAYOUT_REF && a->out_layout != DEGK_LAYOUT_SOA) { degk_set_error(ctx, "bad out_layout"); return DEGK_ERR_INVALID; }
    if (a->dae_init && prog->has_events) {
        degk_set_error(ctx, "dae_init is not available for programs built with tstops / callbacks");
        return DEGK_ERR_UNSUPPORTED;
    }
    if (a->tstops && a->n_tstops > 0 && !prog->has_events) {
        degk_set_error(ctx, "tstops need a program built with degk_model_desc.events = 1");
        return DEGK_ERR_UNSUPPORTED;
    }
    return DEGK_OK;
}

static int launch(degk_program* prog, const degk_solve_args* a, cudaStream_t stream) {
    degk_ctx* ctx = prog->ctx;
    if (a->n_traj == 0) return DEGK_OK;
    degk::KArgs k;
    memset(&k, 0, sizeof k);
    k.n_traj = a->n_traj; k.traj_offset = a->traj_offset;
    k.u0 = a->u0; k.u0_stride = a->u0_stride;
    k.p = a->p; k.p_stride = a->p_stride;
    k.tspan = a->tspan; k.tspan_stride = a->tspan_stride;
    k.saveat = a->saveat; k.n_saveat = a->saveat ? a->n_saveat : 0;
    k.saveat_stride = a->saveat ? a->saveat_stride : 0;
    k.order = a->order;
    if (a->dae_init) k.reserved |= 2;
    k.save_everystep = a->save_everystep ? 1 : 0;
    k.n_rows = a->n_rows; k.us = a->us; k.ts = a->ts;
    k.out_layout = a->out_layout;
    k.retcode = a->retcode; k.naccept = a->naccept; k.nreject = a->nreject; k.nsaved = a->nsaved;
    k.tstops = a->n_tstops > 0 ? a->tstops : nullptr; k.n_tstops = a->tstops ? a->n_tstops : 0;
    k.dt = a->dt; k.abstol = a->abstol; k.reltol = a->reltol;
    k.seed = a->seed; k.reduce = a->reduce; k.totals = (unsigned long long*)a->totals;
    // adaptive attempts per trajectory are capped (the reference would spin forever on a stalled controller); a
    // fixed-dt run makes exactly (tf - t0) / dt steps, so it is never cut short unless the caller asks for a cap
    k.max_iters = a->max_iters > 0 ? a->max_iters : (a->adaptive ? 10000000LL : 0x7fffffffffffffffLL);
    {
        static const int rb = getenv("DEGK_RETIRE_BATCH") ? atoi(getenv("DEGK_RETIRE_BATCH")) : 0;   // tuning knob
        k.retire_batch = rb;
    }

    const int which = (a->adaptive && !prog->is_sde) ? 1 : 0;
    int sched = a->schedule;
    if (sched == DEGK_SCHED_AUTO) sched = which ? DEGK_SCHED_QUEUE : DEGK_SCHED_STATIC;
    if (!which) sched = DEGK_SCHED_STATIC;
    k.schedule = sched;

    // the persistent adaptive kernel (degk_ode_kernels4.cuh) unless the caller pins the first generation.  It stages
    // ONE saveat grid in shared memory: per-problem grids (saveat_stride != 0) and grids longer than
    // DEGK_SAVEAT_STAGE_MAX run on the first-generation kernel, which reads the grid of its trajectory from global memory
    const bool stageable = !a->saveat || (a->n_saveat <= DEGK_SAVEAT_STAGE_MAX && a->saveat_stride == 0);
    const bool v2 = which == 1 && a->engine != DEGK_ENGINE_V1 && prog->info.slots_per_thread2 > 0 && !prog->has_events &&
                    stageable && (prog->info.is_jit ? prog->jit_fn[2] != nullptr : prog->fn[2] != nullptr);
    if (prog->has_events) sched = k.schedule = DEGK_SCHED_STATIC;   // one thread per trajectory
    // fixed dt, one (t0, tf, dt) for the whole launch, every-step saves, explicit RK stepper: the lock-step kernel
    // (degk_ode_lockstep.cuh) at every ensemble size -- its warp-uniform loop is also the shorter dependent chain of a
    // latency-bound small launch (C1 at N = 10^4: 26 us against 60 us, profiles/r2g_c1_sizes.md).
    const bool has_ls = prog->info.is_jit ? prog->jit_fn[3] != nullptr : prog->fn[3] != nullptr;
    const bool ls = which == 0 && has_ls && !prog->is_sde && !prog->has_events && a->engine != DEGK_ENGINE_V1 && !a->saveat &&
                    a->save_everystep && a->tspan_stride == 0 && !a->dae_init && a->n_tstops == 0 && !a->reduce &&
                    // (reference layout with a state too large for the staging area: its scattered stores lose to the
                    //  one-thread-per-trajectory kernel with staged saves once the launch fills the GPU)
                    (a->engine == DEGK_ENGINE_LOCKSTEP || a->out_layout != DEGK_LAYOUT_REF ||
                     degk_lockstep_smem_bytes(prog, 1) <= degk_lockstep_smem_max() ||
                     a->n_traj < (long long)ctx->sm_count * DEGK_BLOCK * 3) &&
                    !getenv("DEGK_NO_LOCKSTEP");
    // Where the build carries two trajectories per thread (fast Float32) there is a one-per-thread twin: a launch that
    // does not fill the GPU is latency-bound, and a dependent chain of packed FMAs runs at half the rate of a scalar
    // one (trajectory-major layout: packed pairs win from ~2 * 10^5 trajectories); with the staged flush of the
    // reference layout the twin is faster at every size (10^6: 138.8 against 133.9 G steps/s)
    const bool ls_staged = ls && a->out_layout == DEGK_LAYOUT_REF && degk_lockstep_smem_bytes(prog, 1) <= degk_lockstep_smem_max();
    long long w1_below = ls_staged ? (1LL << 62) : (long long)ctx->sm_count * DEGK_BLOCK2 * 2 * DEGK_LOCKSTEP_W1_FILL;
    if (const char* e = getenv("DEGK_LOCKSTEP_W1_BELOW")) w1_below = atoll(e);
    const bool ls1 = ls && !prog->info.is_jit && prog->fn[4] && prog->w3 > 1 && a->n_traj < w1_below;
    const int wls = ls1 ? 1 : prog->w3;
    const int block = (v2 || ls) ? DEGK_BLOCK2 : DEGK_BLOCK;
    // the adaptive kernel has the same twin: below DEGK_ADAPTIVE_W1_FILL resident two-trajectory launches' worth of work
    // the one-per-thread form is faster (profiles/r2g_c2_sizes.md)
    long long v2b_below = (long long)ctx->sm_count * DEGK_BLOCK2 * 2 * std::max(1, prog->info.max_blocks_per_sm2) * DEGK_ADAPTIVE_W1_FILL;
    if (const char* e = getenv("DEGK_ADAPTIVE_W1_BELOW")) v2b_below = atoll(e);
    const bool v2b = v2 && !prog->info.is_jit && prog->fn[5] && prog->w2 > 1 && a->n_traj < v2b_below;
    const int per_block = v2 ? DEGK_BLOCK2 * (v2b ? 1 : prog->info.slots_per_thread2) : (ls ? DEGK_BLOCK2 * wls : DEGK_BLOCK);
    size_t smem = 0;
    if (v2) smem = degk_smem2_bytes(prog, a->saveat ? a->n_saveat : 0, v2b ? prog->qcap2b : 0);
    if (ls) {
        // reference layout: a 32 / w-row buffer per trajectory in shared memory, flushed in sector-aligned pieces
        // (degk_ode_lockstep.cuh); the trajectory-major layout needs no staging (lanes already write consecutive
        // addresses), and states too large for the rings go out unstaged as well
        const size_t ring_bytes = degk_lockstep_smem_bytes(prog, wls);
        if (a->out_layout == DEGK_LAYOUT_REF && ring_bytes <= degk_lockstep_smem_max()) {
            k.stage_rows = 16;
            smem = ring_bytes;
        }
    }
    // fixed-dt kernel, every-step saves in the reference layout: stage R rows per lane in shared
    // memory (<= 48 KB per block, so no opt-in attribute is needed) and flush them coalesced
    // (only when the launch fills the GPU: with a few warps per SM the kernel is latency-bound and the
    //  serial flush costs more than the scattered stores -- C1 at N = 10^4: 0.15 ms direct, 0.22 ms staged)
    if (!ls && which == 0 && !prog->is_sde && !prog->has_events && !a->saveat && a->save_everystep &&
        a->out_layout == DEGK_LAYOUT_REF && a->n_traj >= (long long)ctx->sm_count * DEGK_BLOCK * 3 &&
        !getenv("DEGK_NO_STAGED_SAVES")) {
        const size_t es = dtype_size(prog->info.dtype);
        const int n = prog->info.n_state;
        const int words_per_lane = (int)(48 * 1024 / (DEGK_BLOCK * es));
        int R = (words_per_lane - 2) / (n + 1);
        if (R > 16) R = 16;
        if (R >= 4) {
            k.stage_rows = R;
            smem = (size_t)DEGK_BLOCK * ((size_t)n * R + 1 + R + 1) * es;
        }
    }
    long long blocks = (a->n_traj + per_block - 1) / per_block;
    if (sched == DEGK_SCHED_QUEUE) {
        long long resident = (long long)ctx->sm_count * std::max(1, v2 ? (v2b ? prog->max_blocks_per_sm2b : prog->info.max_blocks_per_sm2) : prog->info.max_blocks_per_sm);
        if (blocks > resident) blocks = resident;
        unsigned slot = ctx->next_counter.fetch_add(1) % DEGK_NCOUNTERS;
        k.work_counter = ctx->d_counters + slot;
        CK(ctx, cudaMemsetAsync(k.work_counter, 0, sizeof(unsigned long long), stream));
    }
    if (blocks > 2147483647LL) { degk_set_error(ctx, "too many blocks"); return DEGK_ERR_INVALID; }

    const int kidx = v2 ? (v2b ? 5 : 2) : (ls ? (ls1 ? 4 : 3) : which);
    int rc = DEGK_OK;
    if (prog->info.is_jit) {
        rc = degk_jit_launch(prog, kidx, (unsigned)blocks, (unsigned)block, (unsigned)smem, &k, stream);
    } else {
        void* params[1] = {(void*)&k};
        cudaError_t e = cudaLaunchKernel(prog->fn[kidx], dim3((unsigned)blocks), dim3((unsigned)block), params, smem, stream);
        if (e != cudaSuccess) { degk_set_error(ctx, "kernel launch failed: %s", cudaGetErrorString(e)); rc = DEGK_ERR_CUDA; }
    }
    return rc;
}

extern "C" int degk_solve(degk_program* prog, const degk_solve_args* a, void* stream) {
    if (!prog || !a) return DEGK_ERR_INVALID;
    int rc = validate(prog, a);
    if (rc != DEGK_OK) return rc;
    CK(prog->ctx, cudaSetDevice(prog->ctx->device));
    return launch(prog, a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// host-buffer path
static int ws_get(degk_ctx* ctx, int s, int slot, size_t bytes, void** out) {
    auto& w = ctx->work[s].bufs[slot];
    if (w.cap < bytes) {
        if (w.ptr) CK(ctx, cudaFree(w.ptr));
        w.ptr = nullptr; w.cap = 0;
        size_t cap = bytes + bytes / 8 + 256;
        CK(ctx, cudaMalloc(&w.ptr, cap));
        w.cap = cap;
    }
    *out = w.ptr;
    return DEGK_OK;
}

enum { WS_U0 = 0, WS_P, WS_TSPAN, WS_US, WS_TS, WS_RC, WS_NA, WS_NR, WS_REDUCE, WS_TOTALS, WS_NSAVED, WS_SAVEAT, WS_COUNT };
static_assert(WS_COUNT <= DEGK_NWSBUF, "workspace slots");

// Host-side rebuild of the reference's ts array for saveat runs (lowerlevel_solve.jl:318 fill! +
// integrator_utils.jl:34-47): row k of trajectory i is saveat[k] when k < nsaved[i], else t0_i.
// Transferring 4 bytes per trajectory instead of n_rows values cuts the D2H volume of the C2
// sweep by 25 % (the path is PCIe-bound end to end).
template <class T>
static void rebuild_ts_rows(T* ts, const int32_t* nsaved, const T* saveat, const T* tspan, int64_t tspan_stride,
                            int64_t i0, int64_t i1, int64_t rows) {
    for (int64_t i = i0; i < i1; ++i) {
        T* d = ts + i * rows;
        const int64_t ns = nsaved[i - i0] < 0 ? 0 : (nsaved[i - i0] > rows ? rows : nsaved[i - i0]);
        if (ns == rows) { memcpy(d, saveat, sizeof(T) * (size_t)rows); continue; }
        const T t0 = tspan[i * tspan_stride];
        for (int64_t k = 0; k < ns; ++k) d[k] = saveat[k];
        for (int64_t k = ns; k < rows; ++k) d[k] = t0;
    }
}
static void rebuild_ts(const degk_solve_args* a, size_t es, const int32_t* nsaved, int64_t c0, int64_t cn) {
    // host threads: at most 8, and the host cores are shared by the ranks of a multi-GPU job
    // (torchrun exports LOCAL_WORLD_SIZE)
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* lw = getenv("LOCAL_WORLD_SIZE")) { const int w = atoi(lw); if (w > 1 && hw) hw = std::max(1u, hw / (unsigned)w); }
    if (const char* ht = getenv("DEGK_HOST_THREADS")) { const int w = atoi(ht); if (w > 0) hw = (unsigned)w; }
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw ? hw : 4, 8), cn / 65536 + 1));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const int64_t i0 = c0 + cn * t / nt, i1 = c0 + cn * (t + 1) / nt;
        auto job = [=]() {
            if (es == 4) rebuild_ts_rows<float>((float*)a->ts, nsaved + (i0 - c0), (const float*)a->saveat, (const float*)a->tspan, a->tspan_stride, i0, i1, a->n_rows);
            else rebuild_ts_rows<double>((double*)a->ts, nsaved + (i0 - c0), (const double*)a->saveat, (const double*)a->tspan, a->tspan_stride, i0, i1, a->n_rows);
        };
        if (t + 1 < nt) th.emplace_back(job); else job();
    }
    for (auto& x : th) x.join();
}

extern "C" int degk_solve_host(degk_program* prog, const degk_solve_args* a, int64_t chunk_traj) {
    if (!prog || !a) return DEGK_ERR_INVALID;
    int rc = validate(prog, a);
    if (rc != DEGK_OK) return rc;
    degk_ctx* ctx = prog->ctx;
    std::lock_guard<std::mutex> hostlock(ctx->host_mu);   // workspaces are per ctx
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t es = dtype_size(prog->info.dtype);
    const int n = prog->info.n_state, np = prog->info.n_param;
    const int64_t N = a->n_traj;
    if (N == 0) return DEGK_OK;
    if (chunk_traj <= 0) chunk_traj = 1 << 21;
    if (chunk_traj > N) chunk_traj = N;
    // broadcast inputs and saveat are uploaded once (stream 0), others per chunk
    cudaStream_t s0 = ctx->streams[0];
    void* d_saveat = nullptr;
    if (a->saveat && a->saveat_stride == 0) {
        size_t b = es * a->n_saveat;
        if (ctx->saveat_cap < b) {
            if (ctx->d_saveat) CK(ctx, cudaFree(ctx->d_saveat));
            ctx->d_saveat = nullptr; ctx->saveat_cap = 0;
            CK(ctx, cudaMalloc(&ctx->d_saveat, b + 256));
            ctx->saveat_cap = b + 256;
        }
        d_saveat = ctx->d_saveat;
        CK(ctx, cudaMemcpyAsync(d_saveat, a->saveat, b, cudaMemcpyHostToDevice, s0));
    }
    CK(ctx, cudaStreamSynchronize(s0));

    const size_t red_bytes = a->reduce ? sizeof(double) * a->n_rows * n * 2 : 0;
    const int nchunks = (int)((N + chunk_traj - 1) / chunk_traj);
    const int ns = std::min(DEGK_NSTREAMS, nchunks);
    // saveat runs of the adaptive generation-2/3 kernels: ts is rebuilt on the host from the
    // per-trajectory row counts instead of being transferred
    const bool compact_ts = a->ts && a->saveat && a->saveat_stride == 0 && a->n_saveat <= DEGK_SAVEAT_STAGE_MAX &&
                            a->adaptive && !prog->is_sde && a->engine != DEGK_ENGINE_V1 &&
                            prog->info.slots_per_thread2 > 0 && a->out_layout == DEGK_LAYOUT_REF &&
                            !getenv("DEGK_NO_COMPACT_TS");
    std::thread rebuild_worker;
    std::atomic<int> rebuild_rc{0};
    struct Joiner {                              // every return path stops and joins the worker
        std::thread& t; std::atomic<int>& rc; bool ok = false;
        ~Joiner() { if (t.joinable()) { if (!ok) rc.store(2); t.join(); } }
    } joiner{rebuild_worker, rebuild_rc};
    ctx->chunks_enqueued.store(0);
    if (compact_ts) {
        const size_t need = sizeof(int32_t) * (size_t)N;
        if (ctx->h_nsaved_cap < need) {
            if (ctx->h_nsaved) CK(ctx, cudaFreeHost(ctx->h_nsaved));
            ctx->h_nsaved = nullptr; ctx->h_nsaved_cap = 0;
            CK(ctx, cudaMallocHost((void**)&ctx->h_nsaved, need + need / 8));
            ctx->h_nsaved_cap = need + need / 8;
        }
        while ((int)ctx->chunk_done.size() < nchunks) {
            cudaEvent_t e;
            CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->chunk_done.push_back(e);
        }
    }
    std::vector<double> red_host;
    std::vector<unsigned long long> tot_host;
    if (a->reduce) red_host.assign((size_t)ns * a->n_rows * n * 2, 0.0);
    if (a->totals) tot_host.assign((size_t)ns * 4, 0ULL);
    // per-stream accumulators live on the device for the whole call
    for (int s = 0; s < ns; ++s) {
        void* q;
        if (a->reduce) { rc = ws_get(ctx, s, WS_REDUCE, red_bytes, &q); if (rc) return rc; CK(ctx, cudaMemsetAsync(q, 0, red_bytes, ctx->streams[s])); }
        if (a->totals) { rc = ws_get(ctx, s, WS_TOTALS, 32, &q); if (rc) return rc; CK(ctx, cudaMemsetAsync(q, 0, 32, ctx->streams[s])); }
    }
    for (int c = 0; c < nchunks; ++c) {
        const int s = c % ns;
        cudaStream_t st = ctx->streams[s];
        const int64_t c0 = (int64_t)c * chunk_traj;
        const int64_t cn = std::min<int64_t>(chunk_traj, N - c0);
        degk_solve_args k = *a;
        k.order = nullptr;                     // (a start order refers to device memory of the whole batch)
        k.n_traj = cn;
        k.traj_offset = a->traj_offset + c0;
        k.saveat = d_saveat;
        if (a->saveat && a->saveat_stride != 0) {            // per-problem grids travel with their chunk
            void* dsv;
            const size_t b = es * ((size_t)a->saveat_stride * (cn - 1) + a->n_saveat);
            rc = ws_get(ctx, s, WS_SAVEAT, b, &dsv); if (rc) return rc;
            CK(ctx, cudaMemcpyAsync(dsv, (const char*)a->saveat + es * a->saveat_stride * c0, b, cudaMemcpyHostToDevice, st));
            k.saveat = dsv;
        }
        void *du0, *dp = nullptr, *dts_in, *dus, *dts = nullptr, *drc = nullptr, *dna = nullptr, *dnr = nullptr;
        // inputs
        {
            const int64_t cnt = a->u0_stride ? cn : 1;
            const size_t b = es * (a->u0_stride ? (size_t)a->u0_stride * (cnt - 1) + n : (size_t)n);
            rc = ws_get(ctx, s, WS_U0, b, &du0); if (rc) return rc;
            CK(ctx, cudaMemcpyAsync(du0, (const char*)a->u0 + es * a->u0_stride * c0, b, cudaMemcpyHostToDevice, st));
            k.u0 = du0;
        }
        if (np > 0) {
            const size_t b = es * (a->p_stride ? (size_t)a->p_stride * (cn - 1) + np : (size_t)np);
            rc = ws_get(ctx, s, WS_P, b, &dp); if (rc) return rc;
            CK(ctx, cudaMemcpyAsync(dp, (const char*)a->p + es * a->p_stride * c0, b, cudaMemcpyHostToDevice, st));
            k.p = dp;
        }
        {
            const size_t b = es * (a->tspan_stride ? (size_t)a->tspan_stride * (cn - 1) + 2 : (size_t)2);
            rc = ws_get(ctx, s, WS_TSPAN, b, &dts_in); if (rc) return rc;
            CK(ctx, cudaMemcpyAsync(dts_in, (const char*)a->tspan + es * a->tspan_stride * c0, b, cudaMemcpyHostToDevice, st));
            k.tspan = dts_in;
        }
        // outputs
        const size_t us_b = es * (size_t)cn * a->n_rows * n, ts_b = es * (size_t)cn * a->n_rows;
        rc = ws_get(ctx, s, WS_US, us_b, &dus); if (rc) return rc;
        k.us = dus;
        void* dnsv = nullptr;
        if (compact_ts) { rc = ws_get(ctx, s, WS_NSAVED, 4 * (size_t)cn, &dnsv); if (rc) return rc; }
        else if (a->ts) { rc = ws_get(ctx, s, WS_TS, ts_b, &dts); if (rc) return rc; }
        k.ts = dts;
        k.nsaved = (int32_t*)dnsv;
        if (a->retcode) { rc = ws_get(ctx, s, WS_RC, 4 * (size_t)cn, &drc); if (rc) return rc; }
        if (a->naccept) { rc = ws_get(ctx, s, WS_NA, 4 * (size_t)cn, &dna); if (rc) return rc; }
        if (a->nreject) { rc = ws_get(ctx, s, WS_NR, 4 * (size_t)cn, &dnr); if (rc) return rc; }
        k.retcode = (int32_t*)drc; k.naccept = (int32_t*)dna; k.nreject = (int32_t*)dnr;
        k.reduce = a->reduce ? (double*)ctx->work[s].bufs[WS_REDUCE].ptr : nullptr;
        k.totals = a->totals ? (uint64_t*)ctx->work[s].bufs[WS_TOTALS].ptr : nullptr;
        rc = launch(prog, &k, st);
        if (rc) return rc;
        // download
        if (a->out_layout == DEGK_LAYOUT_REF) {
            CK(ctx, cudaMemcpyAsync((char*)a->us + es * (size_t)c0 * a->n_rows * n, dus, us_b, cudaMemcpyDeviceToHost, st));
            if (compact_ts) CK(ctx, cudaMemcpyAsync(ctx->h_nsaved + c0, dnsv, 4 * (size_t)cn, cudaMemcpyDeviceToHost, st));
            else if (a->ts) CK(ctx, cudaMemcpyAsync((char*)a->ts + es * (size_t)c0 * a->n_rows, dts, ts_b, cudaMemcpyDeviceToHost, st));
        } else {
            CK(ctx, cudaMemcpy2DAsync((char*)a->us + es * c0, es * N, dus, es * cn, es * cn, (size_t)a->n_rows * n, cudaMemcpyDeviceToHost, st));
            if (a->ts) CK(ctx, cudaMemcpy2DAsync((char*)a->ts + es * c0, es * N, dts, es * cn, es * cn, (size_t)a->n_rows, cudaMemcpyDeviceToHost, st));
        }
        if (a->retcode) CK(ctx, cudaMemcpyAsync(a->retcode + c0, drc, 4 * (size_t)cn, cudaMemcpyDeviceToHost, st));
        if (a->naccept) CK(ctx, cudaMemcpyAsync(a->naccept + c0, dna, 4 * (size_t)cn, cudaMemcpyDeviceToHost, st));
        if (a->nreject) CK(ctx, cudaMemcpyAsync(a->nreject + c0, dnr, 4 * (size_t)cn, cudaMemcpyDeviceToHost, st));
        if (compact_ts) {
            CK(ctx, cudaEventRecord(ctx->chunk_done[c], st));
            // the rebuild runs beside the transfers: a worker waits for the chunks in order
            if (c == 0) {
                const int dev = ctx->device;
                rebuild_worker = std::thread([=, &rebuild_rc]() {
                    cudaSetDevice(dev);
                    for (int q = 0; q < nchunks; ++q) {
                        // chunk q's event is recorded by the enqueueing thread before it moves on to
                        // chunk q + 1; spin until it exists in the stream
                        while (ctx->chunks_enqueued.load(std::memory_order_acquire) <= q && !rebuild_rc.load()) std::this_thread::yield();
                        if (rebuild_rc.load()) return;
                        if (cudaEventSynchronize(ctx->chunk_done[q]) != cudaSuccess) { rebuild_rc.store(1); return; }
                        const int64_t p0 = (int64_t)q * chunk_traj;
                        rebuild_ts(a, es, ctx->h_nsaved + p0, p0, std::min<int64_t>(chunk_traj, N - p0));
                    }
                });
            }
            ctx->chunks_enqueued.store(c + 1, std::memory_order_release);
        }
    }
    for (int s = 0; s < ns; ++s) {
        if (a->reduce) CK(ctx, cudaMemcpyAsync(red_host.data() + (size_t)s * a->n_rows * n * 2, ctx->work[s].bufs[WS_REDUCE].ptr, red_bytes, cudaMemcpyDeviceToHost, ctx->streams[s]));
        if (a->totals) CK(ctx, cudaMemcpyAsync(tot_host.data() + (size_t)s * 4, ctx->work[s].bufs[WS_TOTALS].ptr, 32, cudaMemcpyDeviceToHost, ctx->streams[s]));
    }
    for (int s = 0; s < ns; ++s) CK(ctx, cudaStreamSynchronize(ctx->streams[s]));
    if (rebuild_worker.joinable()) {
        joiner.ok = true;
        rebuild_worker.join();
        if (rebuild_rc.load()) { degk_set_error(ctx, "rebuilding ts on the host failed (CUDA event wait)"); return DEGK_ERR_CUDA; }
    }
    if (a->reduce)
        for (int s = 0; s < ns; ++s)
            for (int64_t i = 0; i < a->n_rows * n * 2; ++i) a->reduce[i] += red_host[(size_t)s * a->n_rows * n * 2 + i];
    if (a->totals)
        for (int s = 0; s < ns; ++s)
            for (int i = 0; i < 4; ++i) a->totals[i] += tot_host[(size_t)s * 4 + i];
    return DEGK_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void k_debug_philox(unsigned c0, unsigned c1, unsigned k0, unsigned k1, long long n, unsigned* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned r[4];
    degk::philox4x32_10(c0 + (unsigned)i, c1, 0u, 0u, k0, k1, r);
    for (int q = 0; q < 4; ++q) out[4 * i + q] = r[q];
}

extern "C" int degk_debug_philox(degk_ctx* ctx, uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1,
                                 int64_t n, uint32_t* out_host) {
    if (!ctx || !out_host || n <= 0) return DEGK_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    unsigned* d = nullptr;
    CK(ctx, cudaMalloc(&d, 16 * (size_t)n));
    k_debug_philox<<<(unsigned)((n + 255) / 256), 256>>>(c0, c1, k0, k1, n, d);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpy(out_host, d, 16 * (size_t)n, cudaMemcpyDeviceToHost));
    CK(ctx, cudaFree(d));
    return DEGK_OK;
}
