// Internal (non-ABI) declarations shared by the path-simulator translation units.
#pragma once
#include <vector>

#include "mcp_context.h"

namespace mcp {

// One launch of a path kernel over the contiguous global index range [first, first + M).  Device pointers.
struct PathJob {
    int n = 0, n_steps = 0, dtype = MCP_F32, rounds = 10;
    uint64_t seed = 0, first = 0, M = 0;
    double dt = 0;
    const double* mu = nullptr;          // host, annualised
    const double* w = nullptr;           // host, portfolio weights
    const std::vector<double>* L = nullptr;   // host, lower Cholesky factor of Sigma (row-major n x n)
    const void* z_in = nullptr;          // [M, S, n] supplied normals or null (Philox)
    void* terminal = nullptr;            // [M]
    unsigned long long* hist0 = nullptr; // optional: [1 << MCP_SEL_BITS] counts of the terminal values' first radix digit (FP32 keys), zeroed by the caller
    cudaStream_t stream = nullptr;
    bool hist0_filled = false;           // out: the kernel did fill hist0
};

// tcgen05 path kernel (mcp_paths_tc.cu): FP32, Philox normals, N <= 32
bool path_tc_eligible(const PathJob& job);
int path_launch_tc(mcp_context* h, PathJob& job);
// two-stage tcgen05 kernel with 16-bit split operands for 32 < N <= 256 (mcp_paths_tc16.cu); MCP_PATHS_TC_WIDE16=0 disables it
bool path_tc16_enabled();
int path_launch_tc16(mcp_context* h, PathJob& job);

// warp-per-path kernel for wide universes (mcp_paths.cu): 32 < N <= PATH_WIDE_MAX_N, FP32 / FP64, Philox or supplied normals
constexpr int PATH_WIDE_MAX_N = 1024;

}  // namespace mcp
