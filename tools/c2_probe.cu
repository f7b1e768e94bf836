// c2_probe.cu -- standalone timing harness for the headline kernel (development aid).
//
// Instantiates the adaptive kernel body selected with -DGEN=3|4 for Lorenz / GPUTsit5 / Float32
// on the C2 workload (adaptive tol 1e-6, saveat 0:1:10, p = U[0,1)^3 .* (10, 28, 8/3)), launches it
// exactly like degk_api.cu::launch (persistent grid, work queue) and prints attempted steps/s plus
// checksums of every output, so that two builds can be compared for speed and for identical results
// without Python, torch or a rebuild of libdegk.so:
//
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=true -DDEGK_STRICT=0 -DGEN=4 \
//        -I diffeqgpu.jl_b200/csrc -o tools/bin/c2_probe_v4 tools/c2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#include "device/degk_common.cuh"
#include "device/degk_pack.cuh"
#include "device/degk_models.cuh"
#include "device/gen_erk_tsit5.cuh"
#include "device/degk_ode_kernels.cuh"
#include "device/degk_ode_kernels4.cuh"
#include "degk_internal.h"

#undef GEN
#define GEN 4
#ifndef MINBLOCKS
#define MINBLOCKS 4
#endif
#ifndef WSLOTS
#define WSLOTS (DEGK_STRICT ? 1 : 2)
#endif
using namespace degk;

template <class T, class Model, template <class, class> class Method, int W>
__global__ void __launch_bounds__(DEGK_BLOCK2, MINBLOCKS) k_probe(const __grid_constant__ KArgs a) {
    extern __shared__ __align__(16) unsigned char degk_smem[];
    ode_asolve4_body<T, Model, Method, W>(a, degk_smem);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

static unsigned long long splitmix(unsigned long long& s) {
    unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : (1ll << 23);
    const int reps = argc > 2 ? atoi(argv[2]) : 3;
    const int with_ts = argc > 3 ? atoi(argv[3]) : 1;
    const int with_stats = argc > 4 ? atoi(argv[4]) : 0;      // retcode / naccept / nreject arrays
    const int sched = argc > 5 ? atoi(argv[5]) : 1;           // 0 static, 1 queue
    const int sorted = argc > 6 ? atoi(argv[6]) : 0;          // 1: start order sorted by rho (p[1])
    const int retire_batch = argc > 7 ? atoi(argv[7]) : 0;    // 0: the kernel's default
    constexpr int W = WSLOTS;
    typedef float T;
    const int nsv = 11;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));

    std::vector<float> hp((size_t)N * 3);
    unsigned long long s = 1234;
    const float p0[3] = {10.f, 28.f, 8.f / 3.f};
    for (size_t i = 0; i < hp.size(); ++i) hp[i] = (float)((splitmix(s) >> 40) * (1.0 / 16777216.0)) * p0[i % 3];
    float hu0[3] = {1.f, 0.f, 0.f}, htspan[2] = {0.f, 10.f}, hsv[11];
    for (int i = 0; i < nsv; ++i) hsv[i] = (float)i;
    float *dp, *du0, *dtspan, *dsv, *dus, *dts = nullptr; int *dns, *drc = nullptr, *dna = nullptr, *dnr = nullptr; unsigned long long *dtot, *dctr;
    CK(cudaMalloc(&dp, hp.size() * 4)); CK(cudaMemcpy(dp, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&du0, 12)); CK(cudaMemcpy(du0, hu0, 12, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dtspan, 8)); CK(cudaMemcpy(dtspan, htspan, 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dsv, 44)); CK(cudaMemcpy(dsv, hsv, 44, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dus, (size_t)N * nsv * 3 * 4));
    if (with_ts) CK(cudaMalloc(&dts, (size_t)N * nsv * 4));
    CK(cudaMalloc(&dns, (size_t)N * 4));
    if (with_stats) { CK(cudaMalloc(&drc, (size_t)N * 4)); CK(cudaMalloc(&dna, (size_t)N * 4)); CK(cudaMalloc(&dnr, (size_t)N * 4)); }
    CK(cudaMalloc(&dtot, 32)); CK(cudaMalloc(&dctr, 8));

    int* dorder = nullptr;
    if (sorted) {
        std::vector<int> ord((size_t)N);
        for (long long i = 0; i < N; ++i) ord[i] = (int)i;
        std::sort(ord.begin(), ord.end(), [&](int x, int y) { return hp[(size_t)x * 3 + 1] < hp[(size_t)y * 3 + 1]; });
        CK(cudaMalloc(&dorder, (size_t)N * 4)); CK(cudaMemcpy(dorder, ord.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
    }
    KArgs k; memset(&k, 0, sizeof k);
    k.n_traj = N; k.u0 = du0; k.u0_stride = 0; k.p = dp; k.p_stride = 3; k.tspan = dtspan; k.tspan_stride = 0;
    k.saveat = dsv; k.n_saveat = nsv; k.n_rows = nsv; k.us = dus; k.ts = dts; k.out_layout = LAYOUT_REF; k.schedule = sched ? SCHED_QUEUE : SCHED_STATIC;
    k.order = dorder;
    k.retcode = drc; k.naccept = dna; k.nreject = dnr; k.nsaved = dns;
    k.dt = 0.1f; k.abstol = 1e-6f; k.reltol = 1e-6f; k.totals = dtot; k.work_counter = dctr; k.max_iters = 10000000; k.retire_batch = retire_batch;

    auto kern = k_probe<T, Lorenz, ErkTsit5, W>;
    const size_t smem = asolve4_smem_bytes<T, Lorenz::N, Lorenz::NP, W>(DEGK_BLOCK2 / 32, nsv);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DEGK_BLOCK2, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    const int per_thread = W;
    long long blocks = (N + DEGK_BLOCK2 * per_thread - 1) / (DEGK_BLOCK2 * per_thread);
    const long long resident = (long long)prop.multiProcessorCount * occ;
    if (sched && blocks > resident) blocks = resident;

    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    unsigned long long tot[4] = {0, 0, 0, 0};
    for (int r = 0; r < reps + 1; ++r) {
        CK(cudaMemsetAsync(dtot, 0, 32)); CK(cudaMemsetAsync(dctr, 0, 8));
        CK(cudaMemsetAsync(dus, 0xff, (size_t)N * nsv * 3 * 4));
        if (dts) CK(cudaMemsetAsync(dts, 0xff, (size_t)N * nsv * 4));
        CK(cudaEventRecord(e0));
        void* params[1] = {(void*)&k};
        CK(cudaLaunchKernel((const void*)kern, dim3((unsigned)blocks), dim3(DEGK_BLOCK2), params, smem, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
        CK(cudaMemcpy(tot, dtot, 32, cudaMemcpyDeviceToHost));
    }
    // checksums over a sample of the outputs (first 1M trajectories): exact, order independent
    const long long M = N < (1 << 20) ? N : (1 << 20);
    std::vector<unsigned> hus((size_t)M * nsv * 3), hts((size_t)M * nsv), hns((size_t)M), hrc;
    CK(cudaMemcpy(hus.data(), dus, hus.size() * 4, cudaMemcpyDeviceToHost));
    if (dts) CK(cudaMemcpy(hts.data(), dts, hts.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hns.data(), dns, hns.size() * 4, cudaMemcpyDeviceToHost));
    unsigned long long cu = 0, ct = 0, cn = 0, cs = 0;
    for (size_t i = 0; i < hus.size(); ++i) cu += (unsigned long long)hus[i] * (i % 1000003 + 1);
    if (dts) for (size_t i = 0; i < hts.size(); ++i) ct += (unsigned long long)hts[i] * (i % 1000003 + 1);
    for (size_t i = 0; i < hns.size(); ++i) cn += (unsigned long long)hns[i] * (i % 1000003 + 1);
    if (with_stats) {
        hrc.resize((size_t)M * 3);
        CK(cudaMemcpy(hrc.data(), drc, (size_t)M * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hrc.data() + M, dna, (size_t)M * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hrc.data() + 2 * M, dnr, (size_t)M * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < hrc.size(); ++i) cs += (unsigned long long)hrc[i] * (i % 1000003 + 1);
    }
    const double steps = (double)(tot[0] + tot[1]);
    printf("{\"sched\": %d, \"sorted\": %d, \"gen\": %d, \"strict\": %d, \"W\": %d, \"N\": %lld, \"regs\": %d, \"smem\": %zu, \"blocks_per_sm\": %d, \"ms\": %.3f, "
           "\"gsteps_per_s\": %.2f, \"frac_of_74.45\": %.4f, \"acc\": %llu, \"rej\": %llu, \"fail\": %llu, "
           "\"sum_us\": \"%016llx\", \"sum_ts\": \"%016llx\", \"sum_nsaved\": \"%016llx\", \"sum_stats\": \"%016llx\"}\n",
           sched, sorted, GEN, (int)DEGK_STRICT, W, N, fa.numRegs, smem, occ, best, steps / best / 1e6, 263.0 * steps / (best * 1e-3) / 74.45e12,
           tot[0], tot[1], tot[2], cu, ct, cn, cs);
    return 0;
}
