// degk_aot.cu -- ahead-of-time instantiation of the stepper kernels for the built-in models.
//
// Compiled once per (fp mode, model group):
//   nvcc -gencode arch=compute_100a,code=sm_100a -DDEGK_STRICT=<0|1> --fmad=<true|false>
//        -DDEGK_AOT_GROUP=<g> -DDEGK_AOT_SUFFIX=<fast|strict>_<g>
// Each object exports one table of {model, alg, dtype, adaptive, kernel pointers}; degk_api.cu
// merges them.  User models take the NVRTC path instead (degk_jit.cpp) with the same headers.
//
// FPMODE is a template parameter of every __global__ wrapper so that the strict and the fast
// instantiation of the same kernel get different symbols (they live in different objects but
// would otherwise be merged by the linker as identical weak template instantiations).
#include "device/degk_common.cuh"
#include "device/degk_pack.cuh"
#include "device/degk_models.cuh"
#include "device/gen_erk_tsit5.cuh"
#include "device/gen_erk_vern7.cuh"
#include "device/gen_erk_vern9.cuh"
#include "device/degk_rosenbrock.cuh"
#include "device/degk_kvaerno.cuh"
#include "device/degk_ode_kernels.cuh"
#include "device/degk_ode_kernels4.cuh"
#include "device/degk_ode_lockstep.cuh"
#include "device/degk_sde_kernels.cuh"
#include "degk_internal.h"

namespace degk {

template <class T, class M> using Rodas4M = Rodas<T, M, false>;
template <class T, class M> using Rodas5PM = Rodas<T, M, true>;

// (forcing 4 blocks per SM -- 64 registers, 36 B spilled -- was measured slower on C1: 66 vs 75 G steps/s at N = 10^6)
template <int FPMODE, class T, class Model, template <class, class> class Method>
__global__ void __launch_bounds__(DEGK_BLOCK) k_ode_solve(const KArgs a) {
    extern __shared__ __align__(16) unsigned char degk_smem[];
    ode_solve_body<T, Model, Method<T, Model>>(a, degk_smem);
}
template <int FPMODE, class T, class Model, template <class, class> class Method>
__global__ void __launch_bounds__(DEGK_BLOCK) k_ode_asolve(const KArgs a) {
    ode_asolve_body<T, Model, Method<T, Model>>(a);
}
template <int FPMODE, class T, class Model, template <class, class> class Method, int W>
// register budget: see asolve4_minblocks (Float64: no cap -- Vern9 needs 234 registers, 3 blocks/SM spill ~500 B)
__global__ void __launch_bounds__(DEGK_BLOCK2, (asolve4_minblocks<T, Method<T, Model>>())) k_ode_asolve2(const KArgs a) {
    extern __shared__ __align__(16) unsigned char degk_smem[];
    ode_asolve4_body<T, Model, Method, W>(a, degk_smem);
}
template <int FPMODE, class T, class Model, template <class, class> class Method, int W>
__global__ void __launch_bounds__(DEGK_BLOCK2, (lockstep_minblocks<T>())) k_ode_lockstep(const KArgs a) {
    extern __shared__ __align__(16) unsigned char degk_smem[];
    ode_solve_lockstep_body<T, Model, Method, W>(a, degk_smem);
}
template <int FPMODE, class T, class Model, int ALG>
__global__ void __launch_bounds__(DEGK_BLOCK, DEGK_SDE_MINBLOCKS) k_sde_solve(const KArgs a) {
    sde_solve_body<T, Model, ALG>(a);
}

}  // namespace degk

using namespace degk;

// packed pairs only where un-fused arithmetic is not required (see degk_pack.cuh)
#if DEGK_STRICT
#define WF32 1
#else
#define WF32 2
#endif

#define DIMS(MD) MD::N, MD::NP, MD::M, MD::NOISE
#define V2(T, MD, METHOD, W)                                                                        \
    (const void*)&k_ode_asolve2<DEGK_STRICT, T, MD, METHOD, W>, W, asolve4_qcap<T, MD::N, W>(),     \
        (int)sizeof(SaveRec<T, MD::N>),                                                             \
        ((W) > 1 ? (const void*)&k_ode_asolve2<DEGK_STRICT, T, MD, METHOD, 1> : (const void*)nullptr), asolve4_qcap<T, MD::N, 1>()
#define NOV2 nullptr, 0, 0, 0, nullptr, 0
#define LS(T, MD, METHOD, W) (const void*)&k_ode_lockstep<DEGK_STRICT, T, MD, METHOD, W>, W
#if DEGK_STRICT
#define LS1(T, MD, METHOD) nullptr
#else
#define LS1(T, MD, METHOD) (const void*)&k_ode_lockstep<DEGK_STRICT, T, MD, METHOD, 1>
#endif
// explicit RK: packed pairs for Float32 in fast mode
#define ODE_ERK(NAME, MD, METHOD, ALG)                                                              \
    {NAME, ALG, 0, 0, DIMS(MD), (const void*)&k_ode_solve<DEGK_STRICT, float, MD, METHOD>, NOV2, LS(float, MD, METHOD, WF32), LS1(float, MD, METHOD)},    \
    {NAME, ALG, 0, 1, DIMS(MD), (const void*)&k_ode_asolve<DEGK_STRICT, float, MD, METHOD>, V2(float, MD, METHOD, WF32)}, \
    {NAME, ALG, 1, 0, DIMS(MD), (const void*)&k_ode_solve<DEGK_STRICT, double, MD, METHOD>, NOV2, LS(double, MD, METHOD, 1)},   \
    {NAME, ALG, 1, 1, DIMS(MD), (const void*)&k_ode_asolve<DEGK_STRICT, double, MD, METHOD>, V2(double, MD, METHOD, 1)},
// Rosenbrock: packed pairs too while the linear solve is the closed form (n <= 3: products, sums and one
// reciprocal of the determinant); the pivoting LU of larger systems compares values and stays scalar
#define ODE_ROS(NAME, MD, METHOD, ALG)                                                              \
    {NAME, ALG, 0, 0, DIMS(MD), (const void*)&k_ode_solve<DEGK_STRICT, float, MD, METHOD>, NOV2},    \
    {NAME, ALG, 0, 1, DIMS(MD), (const void*)&k_ode_asolve<DEGK_STRICT, float, MD, METHOD>, V2(float, MD, METHOD, (MD::N <= 3 ? WF32 : 1))}, \
    {NAME, ALG, 1, 0, DIMS(MD), (const void*)&k_ode_solve<DEGK_STRICT, double, MD, METHOD>, NOV2},   \
    {NAME, ALG, 1, 1, DIMS(MD), (const void*)&k_ode_asolve<DEGK_STRICT, double, MD, METHOD>, V2(double, MD, METHOD, 1)},
#define SDE(NAME, MD, ALGK, ALG)                                                                    \
    {NAME, ALG, 0, 0, DIMS(MD), (const void*)&k_sde_solve<DEGK_STRICT, float, MD, ALGK>, NOV2},      \
    {NAME, ALG, 1, 0, DIMS(MD), (const void*)&k_sde_solve<DEGK_STRICT, double, MD, ALGK>, NOV2},
#define ERK3(NAME, MD) ODE_ERK(NAME, MD, ErkTsit5, 0) ODE_ERK(NAME, MD, ErkVern7, 1) ODE_ERK(NAME, MD, ErkVern9, 2)
#define STIFF3(NAME, MD) ODE_ROS(NAME, MD, Rosenbrock23, 3) ODE_ROS(NAME, MD, Rodas4M, 4) ODE_ROS(NAME, MD, Rodas5PM, 5)
#define ODE_SCALAR(NAME, MD, METHOD, ALG)                                                           \
    {NAME, ALG, 0, 0, DIMS(MD), (const void*)&k_ode_solve<DEGK_STRICT, float, MD, METHOD>, NOV2},    \
    {NAME, ALG, 0, 1, DIMS(MD), (const void*)&k_ode_asolve<DEGK_STRICT, float, MD, METHOD>, V2(float, MD, METHOD, 1)}, \
    {NAME, ALG, 1, 0, DIMS(MD), (const void*)&k_ode_solve<DEGK_STRICT, double, MD, METHOD>, NOV2},   \
    {NAME, ALG, 1, 1, DIMS(MD), (const void*)&k_ode_asolve<DEGK_STRICT, double, MD, METHOD>, V2(double, MD, METHOD, 1)},
// Kvaerno: the Newton iteration count is data dependent per trajectory -> one trajectory per thread
#define KVAERNO2(NAME, MD) ODE_SCALAR(NAME, MD, Kvaerno3M, 8) ODE_SCALAR(NAME, MD, Kvaerno5M, 9)

static const degk_aot_entry g_table[] = {
#if DEGK_AOT_GROUP == 0
    ERK3("lorenz", Lorenz)
#elif DEGK_AOT_GROUP == 1
    STIFF3("lorenz", Lorenz)
    SDE("lorenz", Lorenz, ALG_EM, 6) SDE("lorenz", Lorenz, ALG_SIEA, 7)
#elif DEGK_AOT_GROUP == 2
    ERK3("henon_heiles", HenonHeiles)
#elif DEGK_AOT_GROUP == 3
    STIFF3("rober", Rober) ODE_ROS("rober", Rober, ErkTsit5, 0) KVAERNO2("rober", Rober)
    STIFF3("decay", Decay) ODE_ROS("decay", Decay, ErkTsit5, 0)
#elif DEGK_AOT_GROUP == 4
    SDE("gbm", Gbm, ALG_EM, 6) SDE("gbm", Gbm, ALG_SIEA, 7)
    SDE("scalar_sde", ScalarSde, ALG_EM, 6) SDE("scalar_sde", ScalarSde, ALG_SIEA, 7)
    SDE("gbm_nd", GbmNd, ALG_EM, 6)
    ODE_ROS("osc_t", OscT, ErkTsit5, 0) ODE_ROS("osc_t", OscT, ErkVern7, 1) ODE_ROS("osc_t", OscT, Rodas5PM, 5)
#else
#error "unknown DEGK_AOT_GROUP"
#endif
};

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
extern "C" const degk_aot_entry* CAT(degk_aot_table_, DEGK_AOT_SUFFIX)(int* n) {
    *n = (int)(sizeof(g_table) / sizeof(g_table[0]));
    return g_table;
}
