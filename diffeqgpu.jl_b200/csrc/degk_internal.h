// degk_internal.h -- definitions shared by the host-side translation units of libdegk.
#pragma once

#define DEGK_BLOCK 256    // threads per block of the first-generation kernels (__launch_bounds__)
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
    // lock-step fixed-dt kernel (degk_ode_lockstep.cuh): uniform (t0, tf, dt), every-step saves; null if none
    const void* fn3;
    int w3;         // trajectories per thread of fn3 (1 or 2)
};
