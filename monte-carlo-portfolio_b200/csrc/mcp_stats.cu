// mcp_asset_stats: per-asset statistics of a returns matrix.
//
// Reference: sharpe_ratio / sortino_ratio / annual_volatility / annual_return / max_drawdown
// (app.py:231-256), var / cvar (258-263) as combined by calc_asset_stats (286-335).  One warp per
// asset, FP64 throughout (a few thousand values per asset: latency-, not throughput-bound).
// Each lane owns a contiguous chunk of the series, so the one genuinely sequential statistic --
// max drawdown over np.cumprod(1 + r) with its running peak (app.py:252-256) -- becomes two
// warp scans (prefix product, prefix max) plus a local replay.  VaR / CVaR reuse the MSB-first
// warp radix select of mcp_historical.cu, re-reading the column from L1/L2 instead of registers.
#include <cmath>
#include <vector>

#include "mcp_context.h"
#include "mcp_device.cuh"

namespace mcp {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// out[asset][MCP_STATS_FIELDS]: see include/mcp.h
__global__ void __launch_bounds__(32) asset_stats_kernel(const double* __restrict__ R, int T, int N, double rf, double A, int k_lo,
                                                         int k_hi, double gamma, double* __restrict__ out) {
    const int asset = blockIdx.x, lane = threadIdx.x;
    const double* col = R + asset;                        // element t at col[t * N]
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    // ---- moments: mean, std (ddof=1), min, max; downside moments of the excess returns ----
    double s = 0, mn = INFINITY, mx = -INFINITY;
    for (int t = lane; t < T; t += 32) { const double x = col[(size_t)t * N]; s += x; mn = fmin(mn, x); mx = fmax(mx, x); }
    s = wsum(s); mn = wmin(mn); mx = wmax(mx);
    const double mean = s / T;
    const double rf_p = rf / A;
    double ss = 0, ns = 0, ncnt = 0;
    for (int t = lane; t < T; t += 32) {
        const double x = col[(size_t)t * N], d = x - mean, e = x - rf_p;
        ss += d * d;
        if (e < 0) { ns += e; ncnt += 1; }
    }
    ss = wsum(ss); ns = wsum(ns); ncnt = wsum(ncnt);
    const double std1 = T > 1 ? sqrt(ss / (T - 1)) : nan;
    const double nmean = ncnt > 0 ? ns / ncnt : 0.0;
    double nss = 0;
    for (int t = lane; t < T; t += 32) {
        const double e = col[(size_t)t * N] - rf_p;
        if (e < 0) nss += (e - nmean) * (e - nmean);
    }
    nss = wsum(nss);
    const double down = ncnt > 0 ? (ncnt > 1 ? sqrt(nss / (ncnt - 1)) : nan) : 0.0001;      // app.py:243
    const double ex_mean = mean - rf_p;
    const double sharpe = std1 == 0 ? 0.0 : ex_mean / std1 * sqrt(A);                       // app.py:235-236
    const double sortino = ex_mean / down * sqrt(A);
    // ---- compounding: total product and max drawdown (contiguous chunk per lane) ----
    const int chunk = (T + 31) / 32, t0 = lane * chunk, t1 = min(T, t0 + chunk);
    double prod = 1.0, lpeak = 0.0;                        // chunk product, max of the local cumprod
    for (int t = t0; t < t1; ++t) { prod *= 1.0 + col[(size_t)t * N]; lpeak = fmax(lpeak, prod); }
    double pre = prod;                                     // inclusive prefix product
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const double o = __shfl_up_sync(0xffffffffu, pre, d); if (lane >= d) pre *= o; }
    double start = __shfl_up_sync(0xffffffffu, pre, 1);    // cumprod just before this chunk
    if (lane == 0) start = 1.0;
    double gpeak = t1 > t0 ? start * lpeak : 0.0;          // best cumulative value inside this chunk
    double ppeak = gpeak;                                  // inclusive prefix max
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const double o = __shfl_up_sync(0xffffffffu, ppeak, d); if (lane >= d) ppeak = fmax(ppeak, o); }
    double peak = __shfl_up_sync(0xffffffffu, ppeak, 1);
    if (lane == 0) peak = 0.0;
    double cum = start, dd = INFINITY;
    for (int t = t0; t < t1; ++t) {
        cum *= 1.0 + col[(size_t)t * N];
        peak = fmax(peak, cum);
        dd = fmin(dd, (cum - peak) / peak);
    }
    dd = wmin(dd);
    const double total = __shfl_sync(0xffffffffu, pre, 31);
    const double ann_ret = pow(total, A / T) - 1.0;                                         // app.py:249
    // ---- VaR / CVaR: exact k_lo-th order statistic by MSB-first radix select over the column ----
    uint64_t prefix = 0;
    int rank = k_lo;
    for (int b = 63; b >= 0; --b) {
        const uint64_t himask = b == 63 ? 0ull : (~0ull << (b + 1));
        int c = 0;
        for (int t = lane; t < T; t += 32) {
            const uint64_t k = f64_to_key((uint64_t)__double_as_longlong(col[(size_t)t * N]));
            c += ((k & himask) == prefix && !((k >> b) & 1)) ? 1 : 0;
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (rank >= c) { rank -= c; prefix |= 1ull << b; }
    }
    const double v_lo = __longlong_as_double((long long)key_to_f64(prefix));
    int le = 0;
    double above = INFINITY;
    for (int t = lane; t < T; t += 32) {
        const double x = col[(size_t)t * N];
        const uint64_t k = f64_to_key((uint64_t)__double_as_longlong(x));
        le += k <= prefix ? 1 : 0;
        if (k > prefix) above = fmin(above, x);
    }
    le = __reduce_add_sync(0xffffffffu, le);
    above = wmin(above);
    const double v_hi = (k_hi == k_lo || le >= k_hi + 1) ? v_lo : above;
    const double diff = v_hi - v_lo;
    double var = v_lo + diff * gamma;
    if (gamma >= 0.5) var = v_hi - diff * (1 - gamma);
    double ts = 0;
    int tc = 0;
    for (int t = lane; t < T; t += 32) { const double x = col[(size_t)t * N]; if (x <= var) { ts += x; ++tc; } }
    ts = wsum(ts);
    tc = __reduce_add_sync(0xffffffffu, tc);
    const double cvar = tc > 0 ? ts / tc : var;
    if (lane == 0) {
        double* o = out + (size_t)asset * MCP_STATS_FIELDS;
        o[0] = sharpe; o[1] = sortino; o[2] = std1 * sqrt(A); o[3] = ann_ret; o[4] = mean * A; o[5] = mean;
        o[6] = std1; o[7] = mn; o[8] = mx; o[9] = dd; o[10] = var; o[11] = cvar;
    }
}

}  // namespace mcp

using namespace mcp;

static int asset_stats_impl(mcp_handle h, const double* returns_host, int n_periods, int n_assets, double risk_free,
                            double annual_factor, double alpha, double* stats_out) {
    MCP_REQUIRE(h, returns_host && stats_out, "mcp_asset_stats: NULL argument");
    MCP_REQUIRE(h, n_periods >= 1 && n_assets >= 1, "mcp_asset_stats: empty returns matrix");
    MCP_REQUIRE(h, annual_factor > 0 && alpha >= 0 && alpha <= 1, "mcp_asset_stats: bad annual_factor / alpha");
    mcp_device_guard guard(h->device);
    cudaStream_t st = h->stream;
    const size_t in_b = sizeof(double) * (size_t)n_periods * n_assets, out_b = sizeof(double) * (size_t)n_assets * MCP_STATS_FIELDS;
    unsigned char* d = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 1, in_b + out_b + 256, (void**)&d));
    double* d_in = (double*)d;
    double* d_out = (double*)(d + (in_b + 255) / 256 * 256);
    MCP_CUDA(h, cudaMemcpyAsync(d_in, returns_host, in_b, cudaMemcpyHostToDevice, st));
    const double percent = (1 - alpha) * 100, q = percent / 100.0, hidx = (double)(n_periods - 1) * q;
    int k_lo, k_hi;
    if (hidx >= (double)(n_periods - 1)) k_lo = k_hi = n_periods - 1;
    else if (hidx < 0) k_lo = k_hi = 0;
    else { k_lo = (int)std::floor(hidx); k_hi = k_lo + 1; }
    asset_stats_kernel<<<n_assets, 32, 0, st>>>(d_in, n_periods, n_assets, risk_free, annual_factor, k_lo, k_hi,
                                               hidx - std::floor(hidx), d_out);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    MCP_CUDA(h, cudaMemcpyAsync(stats_out, d_out, out_b, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));
    return MCP_OK;
}

extern "C" int mcp_asset_stats(mcp_handle h, const double* returns_host, int n_periods, int n_assets, double risk_free,
                               double annual_factor, double alpha, double* stats_out) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_asset_stats", [&] { return asset_stats_impl(h, returns_host, n_periods, n_assets, risk_free, annual_factor, alpha, stats_out); });
}

// ---------------------------------------------------------------------------------------------
// mu / Sigma estimation (app.py:679-680): mean_returns = returns_df.mean() * A, cov_matrix = returns_df.cov() * A
// (pandas: ddof = 1; the leading fillna(0) row of the returns frame counts like any other row).  FP64 throughout, two
// passes so that the covariance sums products of CENTRED values (what np.cov / pandas' nancorr do): one CTA per column
// for the means, one CTA per (i, j >= i) pair for the co-moments.
// ---------------------------------------------------------------------------------------------
namespace mcp {

constexpr int MOM_BLOCK = 128;

__device__ __forceinline__ double mom_block_sum(double v, double* sh) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < MOM_BLOCK / 32; ++w) s += sh[w];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(MOM_BLOCK) moments_mean_kernel(const double* __restrict__ R, int T, int N, double* __restrict__ mean) {
    __shared__ double sh[MOM_BLOCK / 32];
    const int i = blockIdx.x;
    double s = 0;
    for (int t = threadIdx.x; t < T; t += MOM_BLOCK) s += R[(size_t)t * N + i];
    s = mom_block_sum(s, sh);
    if (threadIdx.x == 0) mean[i] = s / (double)T;
}

__global__ void __launch_bounds__(MOM_BLOCK) moments_cov_kernel(const double* __restrict__ R, int T, int N, const double* __restrict__ mean,
                                                                double A, double* __restrict__ mu, double* __restrict__ sigma) {
    __shared__ double sh[MOM_BLOCK / 32];
    // pair index -> (i, j), j >= i, row-major over the upper triangle
    int p = blockIdx.x, i = 0;
    while (p >= N - i) { p -= N - i; ++i; }
    const int j = i + p;
    const double mi = mean[i], mj = mean[j];
    double s = 0;
    for (int t = threadIdx.x; t < T; t += MOM_BLOCK) s = fma(R[(size_t)t * N + i] - mi, R[(size_t)t * N + j] - mj, s);
    s = mom_block_sum(s, sh);
    if (threadIdx.x == 0) {
        const double c = s / (double)(T - 1) * A;             // T = 1: 0 / 0 = NaN, as pandas
        sigma[(size_t)i * N + j] = c;
        sigma[(size_t)j * N + i] = c;
        if (i == j) mu[i] = mi * A;
    }
}

}  // namespace mcp

static int moments_impl(mcp_handle h, const double* returns_host, int T, int N, double A, double* mu_out, double* sigma_out) {
    MCP_REQUIRE(h, returns_host && mu_out && sigma_out, "mcp_moments: NULL argument");
    MCP_REQUIRE(h, T >= 1 && N >= 1 && N <= 4096, "mcp_moments: bad shape T=%d N=%d", T, N);
    MCP_REQUIRE(h, std::isfinite(A), "mcp_moments: annual_factor is not finite");
    mcp_device_guard guard(h->device);
    cudaStream_t st = h->stream;
    const size_t in_b = sizeof(double) * (size_t)T * N, out_d = (size_t)N * N + 2 * (size_t)N;
    unsigned char* d = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 18, in_b + out_d * sizeof(double) + 256, (void**)&d));
    double* d_in = (double*)d;
    double* d_sigma = (double*)(d + (in_b + 255) / 256 * 256);
    double* d_mu = d_sigma + (size_t)N * N;
    double* d_mean = d_mu + N;
    MCP_CUDA(h, cudaMemcpyAsync(d_in, returns_host, in_b, cudaMemcpyHostToDevice, st));
    moments_mean_kernel<<<N, MOM_BLOCK, 0, st>>>(d_in, T, N, d_mean);
    MCP_CUDA(h, cudaGetLastError());
    moments_cov_kernel<<<(unsigned)((size_t)N * (N + 1) / 2), MOM_BLOCK, 0, st>>>(d_in, T, N, d_mean, A, d_mu, d_sigma);
    MCP_CUDA(h, cudaGetLastError());
    h->launches += 2;
    MCP_CUDA(h, cudaMemcpyAsync(sigma_out, d_sigma, sizeof(double) * (size_t)N * N, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaMemcpyAsync(mu_out, d_mu, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));
    return MCP_OK;
}

extern "C" int mcp_moments(mcp_handle h, const double* returns_host, int n_periods, int n_assets, double annual_factor,
                           double* mu_out, double* sigma_out) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_moments", [&] { return moments_impl(h, returns_host, n_periods, n_assets, annual_factor, mu_out, sigma_out); });
}
