// Host-side context shared by the C-ABI entry points (not part of the public ABI).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <exception>
#include <new>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/mcp.h"

struct mcp_scratch {
    void* p = nullptr;
    size_t cap = 0;
};

struct mcp_context {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t own_stream = nullptr;   // created by mcp_create
    cudaStream_t stream = nullptr;       // current stream (own_stream or the caller's)
    cudaStream_t side_stream[2] = {nullptr, nullptr};   // HOST-space chunk pipeline
    cudaEvent_t ev[8] = {};
    std::string err;
    uint64_t launches = 0;
    double last_ms = 0.0;
    bool allreduce_stream_ordered = false;   // mcp_set_allreduce_stream_ordered
    uint64_t const_epoch = 0;                // bumped by every writer of dev[6] (kernel constants): a cached table is valid only
                                             // while the epoch it was written under is still current
    // device scratch (grow-only): [0] candidates/records, [1..2] pipeline slot inputs,
    // [3..4] pipeline slot outputs, [5] quantile histograms, [6] kernel constants, [7] replay,
    // [8] envelope bins, [9..10] envelope risk/return scratch per slot, [11] recheck rows,
    // [12..13] 1/sum(e) per portfolio of the tcgen05 sweep (per stream), [14] host-buffer collectives (mcp_comm_allgather /
    // allreduce), [15] merge block of a sharded call (records, counts, bins: send + gathered), [16] terminal values of
    // mcp_paths_stats when the caller does not want them, [17] first radix histogram filled by the path kernel,
    // [18] mu / Sigma estimation (returns matrix, results), [19] historical recheck (candidate lists, FP64 rows),
    // [20] constants of the tiled SIMT sweep when it runs next to the tcgen05 sweep (bounded route), [21..22] deferred-row
    // list of the bounded tcgen05 sweep (per pipeline stream), [23] per-block partial terminal sums of the tcgen05 path kernel
    // for 128 < N <= 256
    mcp_scratch dev[24];
    // pinned host scratch: [0..1] HOST-space staging of pageable outputs (per pipeline slot), [2] inputs, [3] collectives,
    // [4] small results (records / stats) read back with one copy
    mcp_scratch pinned[6];
    // multi-GPU (mcp_comm.cu): NCCL communicator of this handle, null until mcp_comm_init
    struct mcp_comm_state* comm = nullptr;
    int comm_rank = 0, comm_size = 0;
    // host thread of this handle in the one-call multi-GPU entry points (mcp_multi.cu), null until first used
    struct mcp_worker* worker = nullptr;
};

// reduction kinds of mcp_comm_allreduce(_dev) -- values are part of the C ABI (include/mcp.h MCP_REDUCE_*)
int mcp_comm_allgather_dev(mcp_context* h, const void* send_dev, void* recv_dev, size_t bytes, cudaStream_t st);
int mcp_comm_allreduce_dev(mcp_context* h, void* buf_dev, size_t count, int kind, cudaStream_t st);
int mcp_comm_check(mcp_context* h);
void mcp_comm_release(mcp_context* h);
void mcp_worker_release(mcp_context* h);

// select + tail passes on device-resident values (mcp_quantile.cu); hist0_dev: optional pre-filled histogram of the first radix
// digit (2048 uint64 counts of this shard, FP32 keys only)
constexpr int MCP_SEL_BITS = 11;
int mcp_quantiles_device(mcp_context* h, const void* v_dev, int dtype, uint64_t n, uint64_t n_total, const double* alphas, int n_alphas,
                         double* var_out, double* cvar_out, mcp_allreduce_fn allreduce, void* user,
                         const unsigned long long* hist0_dev, double* ms_out);

int mcp_fail(mcp_context* h, int code, const char* fmt, ...);
int mcp_dev_reserve(mcp_context* h, int slot, size_t bytes, void** out);
int mcp_pinned_reserve(mcp_context* h, int slot, size_t bytes, void** out);

#define MCP_CUDA(h, call)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess)                                                             \
            return mcp_fail((h), MCP_ERR_CUDA, "%s failed: %s (%s:%d)", #call,             \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                   \
    } while (0)

#define MCP_CHECK(expr)                                                                    \
    do {                                                                                   \
        int _rc = (expr);                                                                  \
        if (_rc != MCP_OK) return _rc;                                                     \
    } while (0)

#define MCP_REQUIRE(h, cond, ...)                                                          \
    do {                                                                                   \
        if (!(cond)) return mcp_fail((h), MCP_ERR_INVALID, __VA_ARGS__);                   \
    } while (0)

// No C++ exception crosses the C ABI ("return int, never throw"): host-side containers can raise std::bad_alloc /
// std::length_error; every entry point that allocates runs its body through this.
template <typename F>
static inline int mcp_guarded(mcp_context* h, const char* what, F&& body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        try { return mcp_fail(h, MCP_ERR_NOMEM, "%s: host allocation failed", what); } catch (...) { return MCP_ERR_NOMEM; }
    } catch (const std::exception& e) {
        try { return mcp_fail(h, MCP_ERR_INVALID, "%s: %s", what, e.what()); } catch (...) { return MCP_ERR_INVALID; }
    } catch (...) {
        try { return mcp_fail(h, MCP_ERR_INVALID, "%s: unknown C++ exception", what); } catch (...) { return MCP_ERR_INVALID; }
    }
}

struct mcp_device_guard {
    int prev = -1;
    explicit mcp_device_guard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~mcp_device_guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
