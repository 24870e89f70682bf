// internal.h -- shared declarations of the B200 primitive library (not part of the C ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/x265b200.h"

namespace b200 {

constexpr int FENC_STRIDE = 64;             // common.h:71
constexpr size_t LANE_BYTES = 1u << 20;     // initial staging per host-call lane (pinned + device); lanes grow on demand
constexpr size_t LANE_MAX_BYTES = (size_t)2 << 30;   // a single host call may stage at most this much (whole 8K planes fit)

// One host-call lane: a stream plus pinned/device staging.  A thread that enters a host (per-call)
// entry borrows a lane for the duration of the call, so slots are re-entrant from any number of
// encoder worker threads without a global lock on the fast path (SURVEY.md 8b "Threading").
struct Lane
{
    cudaStream_t stream = nullptr;
    uint8_t* h = nullptr;       // pinned host staging
    uint8_t* d = nullptr;       // device staging
    size_t cap = 0;             // bytes of each staging buffer
};

} // namespace b200

struct x265b200_ctx
{
    int device = 0;
    int depth = 8;
    int pixbytes = 1;
    int sm_count = 0;
    int dct_path = 0;           // 0 = tensor-core IMMA for N >= 16 (default), 1 = CUDA-core butterfly everywhere
    std::atomic<int> status{0};
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> h2d_bytes{0}, d2h_bytes{0};   // planes + frame jobs (x265b200_transfer_stats)
    std::mutex mu;
    char err[256] = {0};        // first error, written once under `mu` before `status` is published (readers need no lock)
    std::vector<b200::Lane*> free_lanes;
    std::vector<b200::Lane*> all_lanes;
};

namespace b200 {

int fail(x265b200_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess);

#define B200_CUDA(ctx, call)                                                         \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) return b200::fail((ctx), X265B200_ERR_CUDA, #call, e__); \
    } while (0)

// checks the launch itself (configuration errors); asynchronous faults surface at the next sync
#define B200_LAUNCH_CHECK(ctx)                                                       \
    do {                                                                             \
        (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                     \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) return b200::fail((ctx), X265B200_ERR_CUDA, "kernel launch", e__); \
    } while (0)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// tuning lab only: X265B200_LAB="k0,k1,..." selects kernel variants under test; unset = library defaults.
//   k0: 0 = per-CU SATD of 32 / 64 wide CUs on the packed-integer kernel instead of the f16 tensor-core one      csrc/pixel.cu launch_cu_satd
//   k1: 1 = star search as ONE pattern kernel (raster pass inside the warp) instead of the three-launch split   csrc/mesearch.cu launch_me_pattern
inline int lab_knob(int i, int dflt)
{
    static int vals[16];
    static int count = [] {
        int c = 0;
        if (const char* e = getenv("X265B200_LAB"))
            while (*e && c < 16) { vals[c++] = atoi(e); while (*e && *e != ',') e++; if (*e == ',') e++; }
        return c;
    }();
    return i < count ? vals[i] : dflt;
}

} // namespace b200
