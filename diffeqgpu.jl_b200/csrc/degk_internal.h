// degk_internal.h -- definitions shared by the host-side translation units of libdegk.
#pragma once

#define DEGK_BLOCK 256    // threads per block of the first-generation kernels (__launch_bounds__)
#ifndef DEGK_LOCKSTEP_W1_FILL
#define DEGK_LOCKSTEP_W1_FILL 6   // lock-step launches below this many 2-trajectory blocks per SM run one trajectory per thread (measured)
#endif
#ifndef DEGK_ADAPTIVE_W1_FILL
#define DEGK_ADAPTIVE_W1_FILL 5   // adaptive launches below this many resident two-trajectory grids (Lorenz Tsit5: 7.6e5 trajectories) run one trajectory per thread (measured, profiles/r2g_c2_sizes.md)
#endif
#ifndef DEGK_LOCKSTEP_SMEM_MAX
// staging area of the lock-step kernel per block: up to half an SM's shared memory (two blocks per SM).  Measured at 10^6
// trajectories: Henon-Heiles Float32 (67.6 KB) 4.2 ms unstaged -> 1.06 ms staged, Lorenz Float64 (103 KB) 4.5 -> 1.56 ms
#define DEGK_LOCKSTEP_SMEM_MAX (112 * 1024)
#endif
#ifndef DEGK_BLOCK2
#define DEGK_BLOCK2 128   // threads per block of the adaptive kernel (degk_ode_kernels4.cuh)
#endif
#define DEGK_SAVEAT_STAGE_MAX 4096   // saveat grids up to this length are staged in shared memory by the adaptive kernel

struct degk_aot_entry {
    const char* model;
    int alg;        // degk_alg
    int dtype;      // degk_dtype
    int adaptive;   // 0 fixed-dt (and SDE), 1 adaptive
    int n_state, n_param, n_noise, noise_kind;
    const void* fn; // __global__ kernel taking (const degk::KArgs)
    // second-generation adaptive kernel (deferred saves, optionally packed pairs); null if none
    const void* fn2;
    int w2;         // trajectories per thread of fn2 (1 or 2)
    int qcap2;      // save-queue capacity per warp (records)
    int rec_bytes2; // sizeof(SaveRec<T, N>)
    // the same kernel with one trajectory per thread where fn2 carries two (fast Float32 build), for launches that do
    // not fill the GPU (latency-bound: a dependent chain of packed FMAs runs at half the rate of a scalar one); else null
    const void* fn2b;
    int qcap2b;     // its save-queue capacity per warp
    // lock-step fixed-dt kernel (degk_ode_lockstep.cuh): uniform (t0, tf, dt), every-step saves; null if none
    const void* fn3;
    int w3;         // trajectories per thread of fn3 (1 or 2)
    // the same kernel with one trajectory per thread where fn3 carries two (fast Float32 build): launches too small
    // to fill the GPU are latency-bound, and a dependent FFMA2 chain runs at half the rate of a scalar one; else null
    const void* fn4;
};
