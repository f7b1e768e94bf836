// degk_host.h -- host-side object layouts shared by degk_api.cu and degk_jit.cpp.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/degk.h"

#define DEGK_NCOUNTERS 256   // ring of work-queue counters (one per in-flight adaptive launch)
#define DEGK_NSTREAMS 3      // streams used by degk_solve_host to overlap H2D / solve / D2H
#define DEGK_NWSBUF 13

namespace degk { struct KArgs; }

struct degk_wsbuf { void* ptr = nullptr; size_t cap = 0; };
struct degk_workspace { degk_wsbuf bufs[DEGK_NWSBUF]; };

struct degk_ctx {
    int device = 0, sm_count = 0, cc_major = 0, cc_minor = 0;
    std::mutex mu;            // guards err
    std::mutex host_mu;       // serialises degk_solve_host (per-ctx workspaces)
    std::string err, err_copy;
    unsigned long long* d_counters = nullptr;
    std::atomic<unsigned> next_counter{0};
    cudaStream_t streams[DEGK_NSTREAMS] = {nullptr, nullptr, nullptr};
    degk_workspace work[DEGK_NSTREAMS];
    void* d_saveat = nullptr; size_t saveat_cap = 0;
    // degk_solve_host, compact ts: pinned per-stream row-count buffers and "chunk downloaded" events
    int32_t* h_nsaved = nullptr; size_t h_nsaved_cap = 0;
    std::vector<cudaEvent_t> chunk_done;
    std::atomic<int> chunks_enqueued{0};
};

struct degk_program {
    degk_ctx* ctx = nullptr;
    degk_program_info info;
    bool is_sde = false;
    bool has_events = false;                  // built with the tstops / callback kernel pair (degk_ode_events.cuh)
    // AOT kernels: [0] fixed-dt / SDE, [1] adaptive v1, [2] adaptive v2, [3] lock-step fixed-dt
    const void* fn[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // 4 / 5: lock-step / adaptive kernel, one trajectory per thread (fast Float32 build)
    int w2 = 0, qcap2 = 0, rec_bytes2 = 0;             // geometry of the v2 kernel (see degk_internal.h)
    int qcap2b = 0, max_blocks_per_sm2b = 0;           // ... of its one-trajectory-per-thread twin fn[5]
    int w3 = 0;                               // trajectories per thread of the lock-step kernel
    void* jit_module = nullptr;               // CUmodule
    void* jit_fn[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};     // CUfunction, same indexing as fn
};

void degk_set_error(degk_ctx* ctx, const char* fmt, ...);
size_t degk_smem2_bytes(const degk_program* prog, int n_saveat_staged, int qcap = 0);
size_t degk_lockstep_smem_bytes(const degk_program* prog, int w);
size_t degk_lockstep_smem_max();

// NVRTC path (degk_jit.cpp)
int degk_jit_build(degk_ctx* ctx, const degk_model_desc* d, degk_program* prog);
int degk_jit_launch(degk_program* prog, int which, unsigned grid, unsigned block, unsigned smem,
                    const degk::KArgs* args, cudaStream_t stream);
void degk_jit_release(degk_program* prog);
