// degk_jit.cpp -- run-time specialisation of the stepper kernels on a user model via NVRTC.
//
// Replaces what GPUCompiler does for the reference at the first call of a kernel
// (src/ensemblegpukernel/lowerlevel_solve.jl:113, :182-188, :333): the user's RHS / Jacobian /
// tgrad / noise function bodies are spliced into a model struct, the same device headers that
// libdegk's ahead-of-time kernels are built from are handed to NVRTC as in-memory includes, and
// the result is compiled straight to sm_100a SASS so state, parameters and stage vectors stay in
// registers and the tableau coefficients become immediates.
//
// libnvrtc and the driver API are resolved lazily (dlopen / cudaGetDriverEntryPoint) so that
// libdegk.so itself loads on machines without a GPU driver.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "degk_host.h"
#include "degk_internal.h"
#include "degk_embedded.inc"   // generated: degk_embedded_names[], degk_embedded_sources[], degk_embedded_count

namespace {

// ---- NVRTC, loaded on demand ----
typedef void* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    int (*CompileProgram)(nvrtcProgram, int, const char* const*);
    int (*GetProgramLogSize)(nvrtcProgram, size_t*);
    int (*GetProgramLog)(nvrtcProgram, char*);
    int (*GetCUBINSize)(nvrtcProgram, size_t*);
    int (*GetCUBIN)(nvrtcProgram, char*);
    int (*DestroyProgram)(nvrtcProgram*);
    const char* (*GetErrorString)(int);
    int (*Version)(int*, int*);
    std::string load_error;
};
Nvrtc g_nvrtc;
std::once_flag g_nvrtc_once;

void load_nvrtc() {
    const char* cands[] = {getenv("DEGK_NVRTC_LIB"), "libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* c : cands) {
        if (!c || !*c) continue;
        g_nvrtc.h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
        if (g_nvrtc.h) break;
        g_nvrtc.load_error += std::string(c) + ": " + dlerror() + "; ";
    }
    if (!g_nvrtc.h) return;
#define SYM(field, name)                                                        \
    *(void**)(&g_nvrtc.field) = dlsym(g_nvrtc.h, name);                         \
    if (!g_nvrtc.field) { g_nvrtc.load_error = std::string("missing symbol ") + name; dlclose(g_nvrtc.h); g_nvrtc.h = nullptr; return; }
    SYM(CreateProgram, "nvrtcCreateProgram")
    SYM(CompileProgram, "nvrtcCompileProgram")
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SYM(GetProgramLog, "nvrtcGetProgramLog")
    SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SYM(GetCUBIN, "nvrtcGetCUBIN")
    SYM(DestroyProgram, "nvrtcDestroyProgram")
    SYM(GetErrorString, "nvrtcGetErrorString")
    SYM(Version, "nvrtcVersion")
#undef SYM
}

// ---- driver API through the runtime (no link-time dependency on libcuda) ----
struct Driver {
    bool ok = false;
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void**, void**);
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction);
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t);
    CUresult (*GetErrorString)(CUresult, const char**);
    std::string load_error;
};
Driver g_drv;
std::once_flag g_drv_once;

void load_driver() {
    cudaFree(nullptr);   // make sure the primary context exists
#define DSYM(field, name)                                                                         \
    {                                                                                             \
        void* fp = nullptr;                                                                       \
        cudaDriverEntryPointQueryResult qr;                                                       \
        cudaError_t e = cudaGetDriverEntryPoint(name, &fp, cudaEnableDefault, &qr);               \
        if (e != cudaSuccess || !fp) { g_drv.load_error = std::string("driver symbol ") + name + " unavailable"; return; } \
        *(void**)(&g_drv.field) = fp;                                                             \
    }
    DSYM(ModuleLoadData, "cuModuleLoadData")
    DSYM(ModuleGetFunction, "cuModuleGetFunction")
    DSYM(ModuleUnload, "cuModuleUnload")
    DSYM(LaunchKernel, "cuLaunchKernel")
    DSYM(FuncGetAttribute, "cuFuncGetAttribute")
    DSYM(FuncSetAttribute, "cuFuncSetAttribute")
    DSYM(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
    DSYM(GetErrorString, "cuGetErrorString")
#undef DSYM
    g_drv.ok = true;
}

const char* method_header(int alg) {
    switch (alg) {
    case DEGK_ALG_TSIT5: return "gen_erk_tsit5.cuh";
    case DEGK_ALG_VERN7: return "gen_erk_vern7.cuh";
    case DEGK_ALG_VERN9: return "gen_erk_vern9.cuh";
    case DEGK_ALG_ROSENBROCK23: case DEGK_ALG_RODAS4: case DEGK_ALG_RODAS5P: return "degk_rosenbrock.cuh";
    case DEGK_ALG_KVAERNO3: case DEGK_ALG_KVAERNO5: return "degk_kvaerno.cuh";
    default: return "degk_sde_kernels.cuh";
    }
}
const char* method_type(int alg) {
    switch (alg) {
    case DEGK_ALG_TSIT5: return "degk::ErkTsit5<REAL, MODEL>";
    case DEGK_ALG_VERN7: return "degk::ErkVern7<REAL, MODEL>";
    case DEGK_ALG_VERN9: return "degk::ErkVern9<REAL, MODEL>";
    case DEGK_ALG_ROSENBROCK23: return "degk::Rosenbrock23<REAL, MODEL>";
    case DEGK_ALG_RODAS4: return "degk::Rodas<REAL, MODEL, false>";
    case DEGK_ALG_RODAS5P: return "degk::Rodas<REAL, MODEL, true>";
    case DEGK_ALG_KVAERNO3: return "degk::Kvaerno<REAL, MODEL, false>";
    case DEGK_ALG_KVAERNO5: return "degk::Kvaerno<REAL, MODEL, true>";
    default: return "";
    }
}
const char* method_template(int alg) {
    switch (alg) {
    case DEGK_ALG_TSIT5: return "degk::ErkTsit5<T_, M_>";
    case DEGK_ALG_VERN7: return "degk::ErkVern7<T_, M_>";
    case DEGK_ALG_VERN9: return "degk::ErkVern9<T_, M_>";
    case DEGK_ALG_ROSENBROCK23: return "degk::Rosenbrock23<T_, M_>";
    case DEGK_ALG_RODAS4: return "degk::Rodas<T_, M_, false>";
    case DEGK_ALG_RODAS5P: return "degk::Rodas<T_, M_, true>";
    case DEGK_ALG_KVAERNO3: return "degk::Kvaerno<T_, M_, false>";
    case DEGK_ALG_KVAERNO5: return "degk::Kvaerno<T_, M_, true>";
    default: return "";
    }
}
int builtin_n_state(const char* name) {
    static const struct { const char* n; int N; } dims[] = {{"lorenz", 3}, {"henon_heiles", 4}, {"rober", 3}, {"rober_dae", 3}, {"decay", 1},
        {"linear15", 15}, {"gbm", 3}, {"scalar_sde", 1}, {"osc_t", 2}, {"gbm_nd", 2}};
    for (auto& m : dims) if (name && strcmp(m.n, name) == 0) return m.N;
    return 0;
}
// packed pairs (2 trajectories per thread) for Float32 explicit RK in fast mode; a model whose
// body does not compile for the packed type (e.g. it calls cos()) falls back to 1 slot
int default_slots(const degk_model_desc* d) {
    if (d->dtype != DEGK_F32 || d->fp_mode != DEGK_FP_FAST) return 1;
    if (d->alg <= DEGK_ALG_VERN9) return 2;
    // Rosenbrock steppers: packed while the linear solve is the closed form (n <= 3) and the Jacobian is a body
    // (the dual-number / finite-difference paths are scalar)
    const int n = d->rhs_src ? d->n_state : builtin_n_state(d->builtin);
    const bool has_jac = d->rhs_src ? d->jac_src != nullptr : d->jac_mode == 0;
    if (d->alg >= DEGK_ALG_ROSENBROCK23 && d->alg <= DEGK_ALG_RODAS5P && n >= 1 && n <= 3 && has_jac) return 2;
    return 1;
}
const char* builtin_struct(const char* name) {
    static const char* map[][2] = {{"lorenz", "Lorenz"}, {"henon_heiles", "HenonHeiles"}, {"rober", "Rober"}, {"rober_dae", "RoberDae"},
                                   {"decay", "Decay"}, {"linear15", "Linear15"}, {"gbm", "Gbm"},
                                   {"scalar_sde", "ScalarSde"}, {"osc_t", "OscT"}, {"gbm_nd", "GbmNd"}};
    for (auto& m : map) if (strcmp(m[0], name) == 0) return m[1];
    return nullptr;
}

}  // namespace

// Build the translation unit handed to NVRTC.
// host-side mirror of sizeof(degk::SaveRec<T, N>) (checked by a static_assert in the JIT source)
static int save_rec_bytes(int dtype, int n_state) {
    const int es = dtype == DEGK_F64 ? 8 : 4;
    int b = 4 + 4;                       // traj, cur
    b += 3 * es + n_state * es;          // tprev, h, tnew, u[N]
    return (b + 15) / 16 * 16;           // __align__(16)
}

// host-side mirror of degk::asolve4_qcap<T, N, W>() (checked by a static_assert in the JIT source): 32 lanes x
// save_queue_depth records, plus the records' worth of bytes that hold the flush directory
static int save_queue_cap(int rec_bytes, int slots) {
    const int depth = rec_bytes <= 32 ? 4 * slots : (rec_bytes <= 64 ? 2 * slots : slots + 2);
    return 32 * depth + (64 * depth + rec_bytes - 1) / rec_bytes;
}

// `slots` = trajectories per thread of the second-generation adaptive kernel (1 or 2)
static int make_source(degk_ctx* ctx, const degk_model_desc* d, int slots, std::string& src) {
    const bool is_sde = d->alg == DEGK_ALG_EM || d->alg == DEGK_ALG_SIEA;
    const bool kvaerno = d->alg == DEGK_ALG_KVAERNO3 || d->alg == DEGK_ALG_KVAERNO5;
    const bool stiff = (d->alg >= DEGK_ALG_ROSENBROCK23 && d->alg <= DEGK_ALG_RODAS5P) || kvaerno;
    char buf[512];
    src += "#include \"degk_common.cuh\"\n#include \"degk_pack.cuh\"\n";
    src += std::string("#include \"") + method_header(d->alg) + "\"\n";
    const bool events = d->events != 0 || d->n_callbacks > 0 || d->n_ccallbacks > 0;
    if (events && is_sde) {
        degk_set_error(ctx, "tstops / callbacks are available for the ODE solvers only");
        return DEGK_ERR_UNSUPPORTED;
    }
    if (d->n_callbacks < 0 || d->n_callbacks > 16 || (d->n_callbacks > 0 && (!d->cb_condition_src || !d->cb_affect_src))) {
        degk_set_error(ctx, "n_callbacks must be in 0..16 with condition and affect sources");
        return DEGK_ERR_INVALID;
    }
    src += is_sde ? "#include \"degk_sde_kernels.cuh\"\n"
           : events ? "#include \"degk_ode_events.cuh\"\n"
                    : "#include \"degk_ode_kernels.cuh\"\n#include \"degk_ode_kernels4.cuh\"\n#include \"degk_ode_lockstep.cuh\"\n";
    snprintf(buf, sizeof buf, "typedef %s REAL;\n", d->dtype == DEGK_F64 ? "double" : "float");
    src += buf;
    if (d->rhs_src) {
        if (d->n_state <= 0 || d->n_state > 64 || d->n_param < 0 || d->n_noise < 0) {
            degk_set_error(ctx, "n_state must be in 1..64 and n_param, n_noise >= 0");
            return DEGK_ERR_INVALID;
        }
        if (d->jac_mode < 0 || d->jac_mode > 2) { degk_set_error(ctx, "jac_mode must be 0, 1 or 2"); return DEGK_ERR_INVALID; }
        if (is_sde && !d->noise_src) { degk_set_error(ctx, "SDE solvers need noise_src"); return DEGK_ERR_INVALID; }
        const int noise = is_sde ? d->noise_kind : 0;
        const int m = noise == DEGK_NOISE_GENERAL ? d->n_noise : d->n_state;
        src += "namespace degk {\nstruct UserModel {\n";
        // no jac body: nlsolve/type.jl:131-137 -- ForwardDiff (jac_mode 0/2, the reference default
        // autodiff = true) or finite differences (jac_mode 1); see degk_dual.cuh
        const bool has_tgrad = d->jac_src || d->tgrad_src;
        snprintf(buf, sizeof buf,
                 "    static constexpr int N = %d, NP = %d, M = %d, NOISE = %d;\n"
                 "    static constexpr bool HAS_JAC = %s, HAS_TGRAD = %s;\n"
                 "    static constexpr int JAC_MODE = %d;\n",
                 d->n_state, d->n_param, m, noise, d->jac_src ? "true" : "false", has_tgrad ? "true" : "false",
                 d->jac_mode == 1 ? 1 : 2);
        src += buf;
        // a Jacobian body without a time-gradient body declares dT == 0 (autonomous right-hand side): the fast build
        // leaves the dT terms of the Rosenbrock stages out (degk_dual.cuh, tgrad_zero_of)
        if (d->jac_src && !d->tgrad_src) src += "    static constexpr bool TGRAD_ZERO = true;\n";
        src += "    template <class T> static DEGK_DEV void f(T (&du)[N], const T (&u)[N], const T* p, T t) {\n";
        src += d->rhs_src;
        src += "\n    }\n";
        if (d->jac_src) {
            src += "    template <class T> static DEGK_DEV void jac(T (&J)[N][N], const T (&u)[N], const T* p, T t) {\n"
                   "        DEGK_UNROLL for (int i_ = 0; i_ < N; ++i_) DEGK_UNROLL for (int j_ = 0; j_ < N; ++j_) J[i_][j_] = (T)0;\n";
            src += d->jac_src;
            src += "\n    }\n";
        }
        if (has_tgrad) {
            src += "    template <class T> static DEGK_DEV void tgrad(T (&dT)[N], const T (&u)[N], const T* p, T t) {\n"
                   "        DEGK_UNROLL for (int i_ = 0; i_ < N; ++i_) dT[i_] = (T)0;\n";
            if (d->tgrad_src) src += d->tgrad_src;
            src += "\n    }\n";
        }
        if (d->mass_src) {
            if (!stiff) {
                degk_set_error(ctx, "mass matrices need an implicit solver (GPURosenbrock23, GPURodas4, GPURodas5P, GPUKvaerno3, GPUKvaerno5)");
                return DEGK_ERR_UNSUPPORTED;
            }
            src += "    static constexpr bool HAS_MASS = true;\n"
                   "    template <class T> static DEGK_DEV void mass(T (&Mm)[N][N]) {\n"
                   "        DEGK_UNROLL for (int i_ = 0; i_ < N; ++i_) DEGK_UNROLL for (int j_ = 0; j_ < N; ++j_) Mm[i_][j_] = (T)0;\n";
            src += d->mass_src;
            src += "\n    }\n";
        }
        if (noise == DEGK_NOISE_DIAGONAL) {
            src += "    template <class T> static DEGK_DEV void g(T (&g)[N], const T (&u)[N], const T* p, T t) {\n";
            src += d->noise_src;
            src += "\n    }\n";
        } else if (noise == DEGK_NOISE_GENERAL) {
            src += "    template <class T> static DEGK_DEV void G(T (&G)[N][M], const T (&u)[N], const T* p, T t) {\n"
                   "        DEGK_UNROLL for (int i_ = 0; i_ < N; ++i_) DEGK_UNROLL for (int j_ = 0; j_ < M; ++j_) G[i_][j_] = (T)0;\n";
            src += d->noise_src;
            src += "\n    }\n";
        }
        src += "};\n}\ntypedef degk::UserModel MODEL;\n";
    } else {
        const char* st = d->builtin ? builtin_struct(d->builtin) : nullptr;
        if (!st) { degk_set_error(ctx, "unknown built-in model '%s'", d->builtin ? d->builtin : "(null)"); return DEGK_ERR_INVALID; }
        if (strcmp(d->builtin, "rober_dae") == 0 && !stiff) {
            degk_set_error(ctx, "mass matrices need an implicit solver (GPURosenbrock23, GPURodas4, GPURodas5P, GPUKvaerno3, GPUKvaerno5)");
            return DEGK_ERR_UNSUPPORTED;
        }
        src += "#include \"degk_models.cuh\"\n";
        if (d->jac_mode != 0 && stiff) {
            // a built-in model asked to ignore its analytic Jacobian (jac_mode 1: finite differences,
            // 2: forward-mode duals) -- used to validate those paths against the analytic one
            snprintf(buf, sizeof buf,
                     "namespace degk { struct ModelNoJac : %s {\n"
                     "    static constexpr bool HAS_JAC = false, HAS_TGRAD = false;\n"
                     "    static constexpr int JAC_MODE = %d;\n}; }\ntypedef degk::ModelNoJac MODEL;\n", st, d->jac_mode);
            src += buf;
        } else {
            src += std::string("typedef degk::") + st + " MODEL;\n";
        }
    }
    snprintf(buf, sizeof buf, "#define DEGK_JIT_BLOCK %d\n", DEGK_BLOCK);
    src += buf;
    if (is_sde) {
        snprintf(buf, sizeof buf,
                 "extern \"C\" __global__ void __launch_bounds__(DEGK_JIT_BLOCK) degk_jit_fixed(const degk::KArgs a) {\n"
                 "    degk::sde_solve_body<REAL, MODEL, %s>(a);\n}\n",
                 d->alg == DEGK_ALG_EM ? "degk::ALG_EM" : "degk::ALG_SIEA");
        src += buf;
    } else if (events) {
        // callbacks: condition / affect bodies spliced into one struct (see degk_ode_events.cuh)
        src += std::string("typedef ") + method_type(d->alg) + " METHOD;\n";
        src += "namespace degk {\nstruct UserCallbacks {\n";
        snprintf(buf, sizeof buf, "    static constexpr int NCB = %d;\n    static constexpr int N = MODEL::N;\n", d->n_callbacks);
        src += buf;
        for (int c = 0; c < d->n_callbacks; ++c) {
            if (!d->cb_condition_src[c] || !d->cb_affect_src[c]) { degk_set_error(ctx, "callback %d: NULL source", c); return DEGK_ERR_INVALID; }
            snprintf(buf, sizeof buf, "    template <class T> static DEGK_DEV bool condition%d(const T (&u)[N], const T* p, T t) {\n", c);
            src += buf; src += d->cb_condition_src[c]; src += "\n    }\n";
            snprintf(buf, sizeof buf, "    template <class T> static DEGK_DEV void affect%d(T (&u)[N], T* p, T t, bool& terminate_) {\n"
                                      "#define terminate() (terminate_ = true)\n", c);
            src += buf; src += d->cb_affect_src[c]; src += "\n#undef terminate\n    }\n";
        }
        src += "    template <class T> static DEGK_DEV bool condition(int c, const T (&u)[N], const T* p, T t) {\n        switch (c) {\n";
        for (int c = 0; c < d->n_callbacks; ++c) { snprintf(buf, sizeof buf, "        case %d: return condition%d<T>(u, p, t);\n", c, c); src += buf; }
        src += "        default: return false;\n        }\n    }\n";
        src += "    template <class T> static DEGK_DEV void affect(int c, T (&u)[N], T* p, T t, bool& terminate_) {\n        switch (c) {\n";
        for (int c = 0; c < d->n_callbacks; ++c) { snprintf(buf, sizeof buf, "        case %d: affect%d<T>(u, p, t, terminate_); break;\n", c, c); src += buf; }
        src += "        default: break;\n        }\n    }\n";
        // continuous callbacks
        if (d->n_ccallbacks < 0 || d->n_ccallbacks > 8 || (d->n_ccallbacks > 0 && !d->cc_condition_src)) {
            degk_set_error(ctx, "n_ccallbacks must be in 0..8 with condition sources");
            return DEGK_ERR_INVALID;
        }
        snprintf(buf, sizeof buf, "    static constexpr int NCC = %d;\n", d->n_ccallbacks);
        src += buf;
        for (int c = 0; c < d->n_ccallbacks; ++c) {
            if (!d->cc_condition_src[c]) { degk_set_error(ctx, "continuous callback %d: NULL condition", c); return DEGK_ERR_INVALID; }
            snprintf(buf, sizeof buf, "    template <class T> static DEGK_DEV T ccondition%d(const T (&u)[N], const T* p, T t) {\n", c);
            src += buf; src += d->cc_condition_src[c]; src += "\n    }\n";
            for (int neg = 0; neg < 2; ++neg) {
                const char* body = neg ? (d->cc_affect_neg_src ? d->cc_affect_neg_src[c] : nullptr) : (d->cc_affect_src ? d->cc_affect_src[c] : nullptr);
                snprintf(buf, sizeof buf, "    template <class T> static DEGK_DEV void caffect%d_%d(T (&u)[N], T* p, T t, bool& terminate_) {\n"
                                          "#define terminate() (terminate_ = true)\n", c, neg);
                src += buf; if (body) src += body; src += "\n#undef terminate\n    }\n";
            }
        }
        src += "    template <class T> static DEGK_DEV T ccondition(int c, const T (&u)[N], const T* p, T t) {\n        switch (c) {\n";
        for (int c = 0; c < d->n_ccallbacks; ++c) { snprintf(buf, sizeof buf, "        case %d: return ccondition%d<T>(u, p, t);\n", c, c); src += buf; }
        src += "        default: return (T)1;\n        }\n    }\n";
        src += "    template <class T> static DEGK_DEV void caffect(int c, bool neg, T (&u)[N], T* p, T t, bool& terminate_) {\n        switch (2 * c + (neg ? 1 : 0)) {\n";
        for (int c = 0; c < d->n_ccallbacks; ++c)
            for (int neg = 0; neg < 2; ++neg) { snprintf(buf, sizeof buf, "        case %d: caffect%d_%d<T>(u, p, t, terminate_); break;\n", 2 * c + neg, c, neg); src += buf; }
        src += "        default: break;\n        }\n    }\n";
        auto table = [&](const char* sig, const char* deflt, auto value) {
            src += std::string("    static DEGK_DEV ") + sig + " {\n        switch (c) {\n";
            for (int c = 0; c < d->n_ccallbacks; ++c) { char b2[160]; snprintf(b2, sizeof b2, "        case %d: return %s;\n", c, value(c).c_str()); src += b2; }
            src += std::string("        default: return ") + deflt + ";\n        }\n    }\n";
        };
        auto dbl = [](double v) { char b2[64]; snprintf(b2, sizeof b2, "%.17g", v); return std::string(b2); };
        src += "    static DEGK_DEV bool has_caffect(int c, bool neg) {\n        switch (2 * c + (neg ? 1 : 0)) {\n";
        for (int c = 0; c < d->n_ccallbacks; ++c)
            for (int neg = 0; neg < 2; ++neg) {
                const bool has = neg ? (d->cc_affect_neg_src && d->cc_affect_neg_src[c]) : (d->cc_affect_src && d->cc_affect_src[c]);
                snprintf(buf, sizeof buf, "        case %d: return %s;\n", 2 * c + neg, has ? "true" : "false"); src += buf;
            }
        src += "        default: return false;\n        }\n    }\n";
        table("int rootfind(int c)", "0", [&](int c) { return std::to_string(d->cc_rootfind ? d->cc_rootfind[c] : 0); });
        table("double cc_abstol(int c)", "0.0", [&](int c) { return dbl(d->cc_abstol ? d->cc_abstol[c] : 10 * 1.1920928955078125e-7); });
        table("double cc_repeat_nudge(int c)", "0.0", [&](int c) { return dbl(d->cc_repeat_nudge ? d->cc_repeat_nudge[c] : 0.01); });
        table("double cc_dtrelax(int c)", "1.0", [&](int c) { return dbl(d->cc_dtrelax ? d->cc_dtrelax[c] : 1.0); });
        src += "};\n}\n";
        src += "extern \"C\" __global__ void __launch_bounds__(DEGK_JIT_BLOCK) degk_jit_fixed(const degk::KArgs a) {\n"
               "    degk::ode_solve_events_body<REAL, MODEL, METHOD, degk::UserCallbacks>(a);\n}\n"
               "extern \"C\" __global__ void __launch_bounds__(DEGK_JIT_BLOCK) degk_jit_adaptive(const degk::KArgs a) {\n"
               "    degk::ode_asolve_events_body<REAL, MODEL, METHOD, degk::UserCallbacks>(a);\n}\n";
    } else {
        src += std::string("typedef ") + method_type(d->alg) + " METHOD;\n";
        src += std::string("template <class T_, class M_> using METHODT = ") + method_template(d->alg) + ";\n";
        src += "extern \"C\" __global__ void __launch_bounds__(DEGK_JIT_BLOCK) degk_jit_fixed(const degk::KArgs a) {\n"
               "    extern __shared__ __align__(16) unsigned char degk_smem[];\n"
               "    degk::ode_solve_body<REAL, MODEL, METHOD>(a, degk_smem);\n}\n"
               // first-generation adaptive kernel (one thread per trajectory, saveat read from global memory): runs
               // per-problem saveat grids and grids too long for shared memory
               "extern \"C\" __global__ void __launch_bounds__(DEGK_JIT_BLOCK) degk_jit_adaptive(const degk::KArgs a) {\n"
               "    degk::ode_asolve_body<REAL, MODEL, METHOD>(a);\n}\n";
        const int rec_bytes = save_rec_bytes(d->dtype, d->rhs_src ? d->n_state : builtin_n_state(d->builtin));
        snprintf(buf, sizeof buf,
                 "static_assert(sizeof(degk::SaveRec<REAL, MODEL::N>) == %d, \"host/device SaveRec size mismatch\");\n"
                 "static_assert(degk::asolve4_qcap<REAL, MODEL::N, %d>() == %d, \"host/device save-queue capacity mismatch\");\n"
                 "extern \"C\" __global__ void __launch_bounds__(%d, (degk::asolve4_minblocks<REAL, METHOD>())) degk_jit_adaptive2(const degk::KArgs a) {\n"
                 "    extern __shared__ __align__(16) unsigned char degk_smem[];\n"
                 "    degk::ode_asolve4_body<REAL, MODEL, METHODT, %d>(a, degk_smem);\n}\n",
                 rec_bytes, slots, save_queue_cap(rec_bytes, slots), DEGK_BLOCK2, slots);
        src += buf;
        if (d->alg == DEGK_ALG_TSIT5 || d->alg == DEGK_ALG_VERN7 || d->alg == DEGK_ALG_VERN9) {
            // lock-step fixed-dt kernel (uniform tspan and dt, every-step saves): the explicit RK steppers only
            snprintf(buf, sizeof buf,
                     "extern \"C\" __global__ void __launch_bounds__(%d, (degk::lockstep_minblocks<REAL>())) degk_jit_lockstep(const degk::KArgs a) {\n"
                     "    extern __shared__ __align__(16) unsigned char degk_smem[];\n"
                     "    degk::ode_solve_lockstep_body<REAL, MODEL, METHODT, 1>(a, degk_smem);\n}\n",
                     DEGK_BLOCK2);     // one trajectory per thread: faster in the reference layout at every size (degk_api.cu)
            src += buf;
        }
    }
    return DEGK_OK;
}

// Compile to a cubin.  Usable without a GPU (NVRTC is a pure compiler).
static int compile_cubin(degk_ctx* ctx, const degk_model_desc* d, int slots, std::vector<char>& cubin,
                         std::string& log) {
    std::call_once(g_nvrtc_once, load_nvrtc);
    if (!g_nvrtc.h) {
        degk_set_error(ctx, "cannot load libnvrtc (%s) -- set DEGK_NVRTC_LIB", g_nvrtc.load_error.c_str());
        return DEGK_ERR_NVRTC;
    }
    std::string src;
    int rc = make_source(ctx, d, slots, src);
    if (rc != DEGK_OK) return rc;
    nvrtcProgram prog = nullptr;
    int e = g_nvrtc.CreateProgram(&prog, src.c_str(), "degk_jit.cu", degk_embedded_count,
                                  degk_embedded_sources, degk_embedded_names);
    if (e != 0) { degk_set_error(ctx, "nvrtcCreateProgram: %s", g_nvrtc.GetErrorString(e)); return DEGK_ERR_NVRTC; }
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo",
                                     d->fp_mode == DEGK_FP_STRICT ? "-DDEGK_STRICT=1" : "-DDEGK_STRICT=0",
                                     d->fp_mode == DEGK_FP_STRICT ? "--fmad=false" : "--fmad=true"};
    e = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
    size_t ls = 0;
    g_nvrtc.GetProgramLogSize(prog, &ls);
    if (ls > 1) { log.resize(ls); g_nvrtc.GetProgramLog(prog, &log[0]); }
    if (e != 0) {
        degk_set_error(ctx, "NVRTC compilation failed (%s):\n%.3500s", g_nvrtc.GetErrorString(e), log.c_str());
        g_nvrtc.DestroyProgram(&prog);
        return DEGK_ERR_NVRTC;
    }
    size_t cs = 0;
    g_nvrtc.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    g_nvrtc.GetCUBIN(prog, cubin.data());
    g_nvrtc.DestroyProgram(&prog);
    return DEGK_OK;
}

extern "C" int degk_jit_compile_check(const degk_model_desc* d, int64_t* cubin_bytes, char* msg, int64_t msg_cap) {
    std::vector<char> cubin;
    std::string log;
    degk_ctx tmp;
    int slots = default_slots(d);
    int rc = compile_cubin(&tmp, d, slots, cubin, log);
    if (rc == DEGK_ERR_NVRTC && slots == 2) rc = compile_cubin(&tmp, d, 1, cubin, log);
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    if (msg && msg_cap > 0) {
        const std::string& m = rc == DEGK_OK ? log : tmp.err;
        snprintf(msg, (size_t)msg_cap, "%s", m.c_str());
    }
    return rc;
}

#define DRV(ctx, call)                                                                        \
    do {                                                                                      \
        CUresult r_ = (call);                                                                 \
        if (r_ != CUDA_SUCCESS) {                                                             \
            const char* s_ = nullptr;                                                         \
            g_drv.GetErrorString(r_, &s_);                                                    \
            degk_set_error(ctx, "%s failed: %s", #call, s_ ? s_ : "?");                       \
            return DEGK_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

int degk_jit_build(degk_ctx* ctx, const degk_model_desc* d, degk_program* prog) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<char> cubin;
    std::string log;
    int slots = default_slots(d);
    int rc = compile_cubin(ctx, d, slots, cubin, log);
    if (rc == DEGK_ERR_NVRTC && slots == 2) { slots = 1; rc = compile_cubin(ctx, d, slots, cubin, log); }
    if (rc != DEGK_OK) return rc;
    std::call_once(g_drv_once, load_driver);
    if (!g_drv.ok) { degk_set_error(ctx, "CUDA driver API unavailable: %s", g_drv.load_error.c_str()); return DEGK_ERR_CUDA; }
    CUmodule mod = nullptr;
    DRV(ctx, g_drv.ModuleLoadData(&mod, cubin.data()));
    prog->jit_module = mod;
    const bool is_sde = prog->is_sde;
    CUfunction f0 = nullptr, f1 = nullptr;
    DRV(ctx, g_drv.ModuleGetFunction(&f0, mod, "degk_jit_fixed"));
    prog->jit_fn[0] = f0;
    prog->jit_fn[1] = f1;
    const bool events = d->events != 0 || d->n_callbacks > 0 || d->n_ccallbacks > 0;
    prog->has_events = events;
    CUfunction f2 = nullptr;
    if (events) {
        DRV(ctx, g_drv.ModuleGetFunction(&f1, mod, "degk_jit_adaptive"));
        prog->jit_fn[1] = f1;
    } else if (!is_sde) {
        DRV(ctx, g_drv.ModuleGetFunction(&f1, mod, "degk_jit_adaptive"));
        prog->jit_fn[1] = f1;
        DRV(ctx, g_drv.ModuleGetFunction(&f2, mod, "degk_jit_adaptive2"));
    }
    prog->jit_fn[2] = f2;
    if (!events && !is_sde && (d->alg == DEGK_ALG_TSIT5 || d->alg == DEGK_ALG_VERN7 || d->alg == DEGK_ALG_VERN9)) {
        CUfunction f3 = nullptr;
        DRV(ctx, g_drv.ModuleGetFunction(&f3, mod, "degk_jit_lockstep"));
        prog->jit_fn[3] = f3;
        prog->w3 = 1;
        prog->info.dtype = d->dtype;
        prog->info.n_state = d->rhs_src ? d->n_state : builtin_n_state(d->builtin);
        const size_t ls_smem = degk_lockstep_smem_bytes(prog, prog->w3);
        if (ls_smem > 48 * 1024 && ls_smem <= degk_lockstep_smem_max())
            DRV(ctx, g_drv.FuncSetAttribute(f3, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)ls_smem));
    }
    prog->info.is_jit = 1;
    if (d->rhs_src) {
        const int noise = is_sde ? d->noise_kind : 0;
        prog->info.n_state = d->n_state; prog->info.n_param = d->n_param;
        prog->info.n_noise = noise == DEGK_NOISE_GENERAL ? d->n_noise : d->n_state;
        prog->info.noise_kind = noise;
    } else {
        // dims of a JIT-compiled built-in: mirror of degk_models.cuh
        static const struct { const char* n; int N, NP, M, K; } dims[] = {
            {"lorenz", 3, 3, 3, 1}, {"henon_heiles", 4, 0, 0, 0}, {"rober", 3, 3, 0, 0}, {"rober_dae", 3, 3, 0, 0}, {"decay", 1, 1, 0, 0},
            {"linear15", 15, 0, 0, 0}, {"gbm", 3, 2, 3, 1}, {"scalar_sde", 1, 2, 1, 1}, {"osc_t", 2, 1, 0, 0},
            {"gbm_nd", 2, 2, 4, 2}};
        for (auto& m : dims)
            if (strcmp(m.n, d->builtin) == 0) {
                prog->info.n_state = m.N; prog->info.n_param = m.NP; prog->info.n_noise = m.M; prog->info.noise_kind = m.K;
            }
    }
    int v = 0;
    DRV(ctx, g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, f0)); prog->info.regs_fixed = v;
    DRV(ctx, g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f0)); prog->info.local_bytes_fixed = v;
    if (f1) {
        DRV(ctx, g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, f1)); prog->info.regs_adaptive = v;
        DRV(ctx, g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f1)); prog->info.local_bytes_adaptive = v;
    }
    DRV(ctx, g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&v, f1 ? f1 : f0, DEGK_BLOCK, 0));
    prog->info.max_blocks_per_sm = v;
    if (f2) {
        prog->w2 = slots;
        prog->rec_bytes2 = save_rec_bytes(d->dtype, prog->info.n_state);
        prog->qcap2 = save_queue_cap(prog->rec_bytes2, slots);
        DRV(ctx, g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, f2)); prog->info.regs_adaptive2 = v;
        DRV(ctx, g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f2)); prog->info.local_bytes_adaptive2 = v;
        prog->info.slots_per_thread2 = slots;
        const size_t smem = degk_smem2_bytes(prog, 1024), smem_max = degk_smem2_bytes(prog, DEGK_SAVEAT_STAGE_MAX);
        if (smem_max > 48 * 1024) DRV(ctx, g_drv.FuncSetAttribute(f2, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_max));
        DRV(ctx, g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&v, f2, DEGK_BLOCK2, smem));
        prog->info.max_blocks_per_sm2 = v;
    }
    prog->info.jit_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return DEGK_OK;
}

int degk_jit_launch(degk_program* prog, int which, unsigned grid, unsigned block, unsigned smem,
                    const degk::KArgs* args, cudaStream_t stream) {
    degk_ctx* ctx = prog->ctx;
    CUfunction f = (CUfunction)prog->jit_fn[which];
    if (!f) { degk_set_error(ctx, "program has no %s kernel", which ? "adaptive" : "fixed-dt"); return DEGK_ERR_UNSUPPORTED; }
    void* params[1] = {(void*)args};
    DRV(ctx, g_drv.LaunchKernel(f, grid, 1, 1, block, 1, 1, smem, (CUstream)stream, params, nullptr));
    return DEGK_OK;
}

void degk_jit_release(degk_program* prog) {
    if (prog->jit_module && g_drv.ok) g_drv.ModuleUnload((CUmodule)prog->jit_module);
    prog->jit_module = nullptr;
}
