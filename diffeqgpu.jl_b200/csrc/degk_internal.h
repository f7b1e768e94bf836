// degk_internal.h -- definitions shared by the host-side translation units of libdegk.
#pragma once

#define DEGK_BLOCK 256   // threads per block of every stepper kernel (__launch_bounds__)

struct degk_aot_entry {
    const char* model;
    int alg;        // degk_alg
    int dtype;      // degk_dtype
    int adaptive;   // 0 fixed-dt (and SDE), 1 adaptive
    int n_state, n_param, n_noise, noise_kind;
    const void* fn; // __global__ kernel taking (const degk::KArgs)
};
