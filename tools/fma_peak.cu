// fma_peak.cu -- measures the FP32/FP64 FMA issue peak of the GPU it runs on, so that bench.py
// can state "% of measured FMA peak" (MEASURED_PEAKS.json only holds HBM and tensor numbers).
// Variants: 3-register FFMA, immediate-operand FFMA, packed FFMA2 (fma.rn.f32x2), DFMA.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__global__ void k_ffma_reg(float* out, float a, float b) {
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-6f + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma_imm(float* out) {
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-6f + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], 0.999f, 0.25f);
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long x[ILP], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    for (int i = 0; i < ILP; ++i) { float v = threadIdx.x * 1e-6f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    unsigned long long s = 0;
    for (int i = 0; i < ILP; ++i) s ^= x[i];
    if (s == 12345ull) out[0] = 1.f;
}
__global__ void k_dfma(float* out, double a, double b) {
    double x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-6 + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456) out[0] = (float)s;
}

template <class F> double run(F f, int blocks, double flops_per_thread) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(blocks); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(blocks); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return flops_per_thread * blocks * 256.0 / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d; cudaMalloc(&d, 4);
    int blocks = p.multiProcessorCount * 8 * 4;
    double fl = 2.0 * ITERS * ILP;
    double t_reg = run([&](int b) { k_ffma_reg<<<b, 256>>>(d, 0.999f, 0.25f); }, blocks, fl);
    double t_imm = run([&](int b) { k_ffma_imm<<<b, 256>>>(d); }, blocks, fl);
    double t_f2 = run([&](int b) { k_ffma2<<<b, 256>>>(d, 0.999f, 0.25f); }, blocks, 2 * fl);
    double t_d = run([&](int b) { k_dfma<<<b, 256>>>(d, 0.999, 0.25); }, blocks, fl);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"ffma_reg_tflops\": %.2f, \"ffma_imm_tflops\": %.2f, "
           "\"ffma2_tflops\": %.2f, \"dfma_tflops\": %.2f}\n",
           p.name, p.multiProcessorCount, t_reg, t_imm, t_f2, t_d);
    return 0;
}
