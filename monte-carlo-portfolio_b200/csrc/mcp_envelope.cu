// Frontier envelope: per risk bin the maximum return and the
// first (lowest) global index that attains it.  Replaces the reference's per-point scatter
// (app.py:726-736), which is meaningless at 1e9 points.
//
// Runs as a post-pass over the (risk, return) arrays a sweep chunk has just written (8 or 16
// bytes per portfolio, L2 / HBM-bound, < 1 % of the sweep's time at N = 256):
//   env_max  : bin = floor((risk - lo) * K / (hi - lo)); per-CTA shared-memory atomicMax on an
//              order-preserving key of the return, flushed with global atomicMax;
//   env_idx  : atomicMin of the global index over the elements whose key equals the bin maximum;
//   env_merge: folds a chunk's bins into the running bins (larger key, then lower index).
#include "mcp_device.cuh"
#include "mcp_portfolio.h"

namespace mcp {

constexpr int ENV_BLOCK = 256;

template <typename T> struct EnvKey;
template <> struct EnvKey<float> {
    static __device__ __forceinline__ unsigned long long key(float v) { return (unsigned long long)f32_to_key(__float_as_uint(v)); }
};
template <> struct EnvKey<double> {
    static __device__ __forceinline__ unsigned long long key(double v) { return f64_to_key((uint64_t)__double_as_longlong(v)); }
};

template <typename T>
__device__ __forceinline__ int env_bin(T risk, T lo, T hi, T scale, int K) {
    if (!(risk >= lo && risk <= hi)) return -1;            // also drops NaN (skipped portfolios)
    int b = (int)floor((risk - lo) * scale);
    if (risk == hi || b >= K) b = K - 1;
    return b;
}

template <typename T>
__global__ void __launch_bounds__(ENV_BLOCK) env_max(const T* __restrict__ risk, const T* __restrict__ ret, uint64_t n, T lo, T hi,
                                                     T scale, int K, unsigned long long* __restrict__ binmax) {
    extern __shared__ unsigned long long sbin[];
    for (int i = threadIdx.x; i < K; i += ENV_BLOCK) sbin[i] = 0ull;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * ENV_BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * ENV_BLOCK) {
        const int b = env_bin<T>(risk[i], lo, hi, scale, K);
        if (b >= 0) atomicMax(&sbin[b], EnvKey<T>::key(ret[i]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += ENV_BLOCK)
        if (sbin[i]) atomicMax(&binmax[i], sbin[i]);
}

template <typename T>
__global__ void __launch_bounds__(ENV_BLOCK) env_idx(const T* __restrict__ risk, const T* __restrict__ ret, uint64_t n, T lo, T hi,
                                                     T scale, int K, const unsigned long long* __restrict__ binmax,
                                                     unsigned long long* __restrict__ binidx, uint64_t base) {
    for (uint64_t i = (uint64_t)blockIdx.x * ENV_BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * ENV_BLOCK) {
        const int b = env_bin<T>(risk[i], lo, hi, scale, K);
        if (b >= 0 && EnvKey<T>::key(ret[i]) == binmax[b]) atomicMin(&binidx[b], (unsigned long long)(base + i));
    }
}

__global__ void env_merge(int K, const unsigned long long* cmax, const unsigned long long* cidx, unsigned long long* fmax,
                          unsigned long long* fidx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    if (cmax[i] > fmax[i] || (cmax[i] == fmax[i] && cidx[i] < fidx[i])) { fmax[i] = cmax[i]; fidx[i] = cidx[i]; }
}

__global__ void env_clear(int K, unsigned long long* mx, unsigned long long* ix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) { mx[i] = 0ull; ix[i] = ~0ull; }
}

int env_reset(mcp_context* h, int K, unsigned long long* mx, unsigned long long* ix, cudaStream_t st) {
    env_clear<<<(K + 255) / 256, 256, 0, st>>>(K, mx, ix);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

int env_fold(mcp_context* h, int K, const unsigned long long* cmax, const unsigned long long* cidx, unsigned long long* fmax,
             unsigned long long* fidx, cudaStream_t st) {
    env_merge<<<(K + 255) / 256, 256, 0, st>>>(K, cmax, cidx, fmax, fidx);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <typename T>
static int env_chunk_t(mcp_context* h, const void* risk, const void* ret, uint64_t n, uint64_t base, double lo, double hi, int K,
                       unsigned long long* cmax, unsigned long long* cidx, cudaStream_t st) {
    const T scale = (T)((double)K / (hi - lo));
    uint64_t g = (n + ENV_BLOCK - 1) / ENV_BLOCK;
    const uint64_t cap = (uint64_t)h->prop.multiProcessorCount * 8;
    const unsigned grid = (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
    env_max<T><<<grid, ENV_BLOCK, (size_t)K * 8, st>>>((const T*)risk, (const T*)ret, n, (T)lo, (T)hi, scale, K, cmax);
    env_idx<T><<<grid, ENV_BLOCK, 0, st>>>((const T*)risk, (const T*)ret, n, (T)lo, (T)hi, scale, K, cmax, cidx, base);
    MCP_CUDA(h, cudaGetLastError());
    h->launches += 2;
    return MCP_OK;
}

// bins one chunk into (cmax, cidx) -- which the caller has reset -- with global indices base + i
int env_chunk(mcp_context* h, int dtype, const void* risk, const void* ret, uint64_t n, uint64_t base, double lo, double hi, int K,
              unsigned long long* cmax, unsigned long long* cidx, cudaStream_t st) {
    return dtype == MCP_F64 ? env_chunk_t<double>(h, risk, ret, n, base, lo, hi, K, cmax, cidx, st)
                            : env_chunk_t<float>(h, risk, ret, n, base, lo, hi, K, cmax, cidx, st);
}

}  // namespace mcp
