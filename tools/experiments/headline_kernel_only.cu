// scratch: only the headline kernel (Lorenz, Tsit5, Float32, 2 slots, fast mode) for static SASS experiments
#include "device/degk_common.cuh"
#include "device/degk_pack.cuh"
#include "device/degk_models.cuh"
#include "device/gen_erk_tsit5.cuh"
#include "device/degk_ode_kernels.cuh"
#include "device/degk_ode_kernels2.cuh"
#include "device/degk_ode_kernels3.cuh"
#include "degk_internal.h"
namespace degk {
template <int FPMODE, class T, class Model, template <class, class> class Method, int W>
__global__ void __launch_bounds__(DEGK_BLOCK2, 4) k_one(const KArgs a) {
    extern __shared__ __align__(16) unsigned char degk_smem[];
    ode_asolve_gen_body<T, Model, Method, W>(a, degk_smem);
}
template __global__ void k_one<0, float, Lorenz, ErkTsit5, 2>(const KArgs);
}
