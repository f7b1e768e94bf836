// degk_aot.cu -- ahead-of-time instantiation of the stepper kernels for the built-in models.
//
// Compiled once per (fp mode, model group):
//   nvcc -gencode arch=compute_100a,code=sm_100a -DDEGK_STRICT=<0|1> --fmad=<true|false>
//        -DDEGK_AOT_GROUP=<g> -DDEGK_AOT_SUFFIX=<fast|strict>_<g>
// Each object exports one table of {model, alg, dtype, adaptive, kernel pointer}; degk_api.cu
// merges them.  User models take the NVRTC path instead (degk_jit.cpp) with the same headers.
#include "device/degk_common.cuh"
#include "device/degk_models.cuh"
#include "device/gen_erk_tsit5.cuh"
#include "device/gen_erk_vern7.cuh"
#include "device/gen_erk_vern9.cuh"
#include "device/degk_rosenbrock.cuh"
#include "device/degk_ode_kernels.cuh"
#include "device/degk_sde_kernels.cuh"
#include "degk_internal.h"

namespace degk {

template <class T, class M> using Rodas4M = Rodas<T, M, false>;
template <class T, class M> using Rodas5PM = Rodas<T, M, true>;

template <int FPMODE, class T, class Model, template <class, class> class Method>
__global__ void __launch_bounds__(DEGK_BLOCK) k_ode_solve(const KArgs a) {
    ode_solve_body<T, Model, Method<T, Model>>(a);
}
template <int FPMODE, class T, class Model, template <class, class> class Method>
__global__ void __launch_bounds__(DEGK_BLOCK) k_ode_asolve(const KArgs a) {
    ode_asolve_body<T, Model, Method<T, Model>>(a);
}
template <int FPMODE, class T, class Model, int ALG>
__global__ void __launch_bounds__(DEGK_BLOCK) k_sde_solve(const KArgs a) {
    sde_solve_body<T, Model, ALG>(a);
}

}  // namespace degk

using namespace degk;

#define DIMS(MD) MD::N, MD::NP, MD::M, MD::NOISE
#define ODE(NAME, MD, METHOD, ALG)                                                                   \
    {NAME, ALG, 0, 0, DIMS(MD), (const void*)&k_ode_solve<DEGK_STRICT, float, MD, METHOD>},                        \
    {NAME, ALG, 0, 1, DIMS(MD), (const void*)&k_ode_asolve<DEGK_STRICT, float, MD, METHOD>},                       \
    {NAME, ALG, 1, 0, DIMS(MD), (const void*)&k_ode_solve<DEGK_STRICT, double, MD, METHOD>},                       \
    {NAME, ALG, 1, 1, DIMS(MD), (const void*)&k_ode_asolve<DEGK_STRICT, double, MD, METHOD>},
#define SDE(NAME, MD, ALGK, ALG)                                                                     \
    {NAME, ALG, 0, 0, DIMS(MD), (const void*)&k_sde_solve<DEGK_STRICT, float, MD, ALGK>},                          \
    {NAME, ALG, 1, 0, DIMS(MD), (const void*)&k_sde_solve<DEGK_STRICT, double, MD, ALGK>},
#define ERK3(NAME, MD) ODE(NAME, MD, ErkTsit5, 0) ODE(NAME, MD, ErkVern7, 1) ODE(NAME, MD, ErkVern9, 2)
#define STIFF3(NAME, MD) ODE(NAME, MD, Rosenbrock23, 3) ODE(NAME, MD, Rodas4M, 4) ODE(NAME, MD, Rodas5PM, 5)

static const degk_aot_entry g_table[] = {
#if DEGK_AOT_GROUP == 0
    ERK3("lorenz", Lorenz)
#elif DEGK_AOT_GROUP == 1
    STIFF3("lorenz", Lorenz)
    SDE("lorenz", Lorenz, ALG_EM, 6) SDE("lorenz", Lorenz, ALG_SIEA, 7)
#elif DEGK_AOT_GROUP == 2
    ERK3("henon_heiles", HenonHeiles)
#elif DEGK_AOT_GROUP == 3
    STIFF3("rober", Rober) ODE("rober", Rober, ErkTsit5, 0)
    STIFF3("decay", Decay) ODE("decay", Decay, ErkTsit5, 0)
#elif DEGK_AOT_GROUP == 4
    SDE("gbm", Gbm, ALG_EM, 6) SDE("gbm", Gbm, ALG_SIEA, 7)
    SDE("scalar_sde", ScalarSde, ALG_EM, 6) SDE("scalar_sde", ScalarSde, ALG_SIEA, 7)
    SDE("gbm_nd", GbmNd, ALG_EM, 6)
    ODE("osc_t", OscT, ErkTsit5, 0) ODE("osc_t", OscT, ErkVern7, 1) ODE("osc_t", OscT, Rodas5PM, 5)
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
