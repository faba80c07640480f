// mcp_paths: correlated-return path simulator (north-star row a10; not in the reference).
//
//   L = chol(Sigma_annual);  per step  r = mu dt + sqrt(dt) L z,  z ~ N(0, I_N)
//   V_i *= (1 + r_i)  (per-asset cumulative product, the compounding convention of
//   np.cumprod(1 + returns), app.py:253);  terminal[m] = w . V_T - 1
//
// One thread per path, the cumulative product stays in registers for all S steps.  L * sqrt(dt)
// (packed lower triangle), mu * dt and w are kernel parameters -> constant-bank FFMA operands,
// so RNG mode touches HBM only for the 4- or 8-byte terminal value per path.  Normals come from
// Philox4x32-10 (counter = global path index, step, 4-normal block) through Box-Muller on the
// MUFU pipe (lg2, sqrt, sin, cos).  Supplied-normals mode reads Z[m, s, :] with 128-bit loads;
// a thread walks its own contiguous S*N row, so every fetched sector is fully consumed
// (L1 holds the line between the two halves).
#include <cmath>
#include <vector>

#include "mcp_device.cuh"
#include "mcp_paths.h"

namespace mcp {

constexpr int PATH_BLOCK = 128;

template <typename T, int NP>
struct PathArgs {
    T lp[NP * (NP + 1) / 2];   // L * sqrt(dt), row i holds j = 0..i
    T drift[NP];               // mu * dt
    T w[NP];                   // portfolio weights (0 for padded assets)
    int n, n_steps;
    uint32_t k0, k1;
    uint64_t first, M;
    const T* z_in;             // [M, S, n] or null
    T* terminal;               // [M]
};

template <typename T> struct PathConst;
template <> struct PathConst<float> {
    static __device__ __forceinline__ float neg2ln2() { return -1.3862943611198906f; }
    static __device__ __forceinline__ float pi() { return 3.14159265358979323846f; }
    // 2f - 1 in [-1, 1) from 23 random bits: float in [2, 4) minus 3
    static __device__ __forceinline__ float centred(uint32_t x) { return __uint_as_float((x & 0x007fffffu) | 0x40000000u) - 3.0f; }
};
template <> struct PathConst<double> {
    static __device__ __forceinline__ double neg2ln2() { return -1.3862943611198906; }
    static __device__ __forceinline__ double pi() { return 3.14159265358979323846; }
    static __device__ __forceinline__ double centred(uint32_t x) { return 2.0 * ((double)x * 0x1p-32) - 1.0; }
};

template <typename T, int NP, int ROUNDS>
__device__ __forceinline__ void draw_normals(const PathArgs<T, NP>& a, uint32_t c0, uint32_t c1, uint32_t step, T (&z)[NP]) {
    if constexpr (sizeof(T) == 4) {           // FP32: 24-bit fields (mcp_device.cuh), pair (2k, 2k+1) -> normals (2k, 2k+1)
        uint32_t f[NP];
        philox_fields<NP, ROUNDS>(c0, c1, step, STREAM_NORMALS, a.k0, a.k1, f);
#pragma unroll
        for (int k = 0; k < NP / 2; ++k) {
            const T u1 = Math<T>::unit_open0(f[2 * k]);
            const T r = Math<T>::sqrt(Math<T>::lg2(u1) * PathConst<T>::neg2ln2());
            const T th = PathConst<T>::centred(f[2 * k + 1]) * PathConst<T>::pi();
            z[2 * k] = r * Math<T>::cosf_(th);
            z[2 * k + 1] = r * Math<T>::sinf_(th);
        }
    } else {
#pragma unroll
        for (int b = 0; b < NP / 4; ++b) {
            uint32_t x[4];
            philox4x32_r<ROUNDS>(c0, c1, step, STREAM_NORMALS | (uint32_t)b, a.k0, a.k1, x);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const T u1 = Math<T>::unit_open0(x[2 * k]);
                const T r = Math<T>::sqrt(Math<T>::lg2(u1) * PathConst<T>::neg2ln2());
                // FP64: sincospi on the angle in units of pi (no Payne-Hanek reduction, one call for both values)
                double sn, cs;
                ::sincospi((double)PathConst<T>::centred(x[2 * k + 1]), &sn, &cs);
                z[4 * b + 2 * k] = r * (T)cs;
                z[4 * b + 2 * k + 1] = r * (T)sn;
            }
        }
    }
}

template <typename T, int NP, int SRC, int ROUNDS>
__global__ void __launch_bounds__(PATH_BLOCK) path_kernel(const __grid_constant__ PathArgs<T, NP> a) {
    const uint64_t stride = (uint64_t)gridDim.x * PATH_BLOCK;
    for (uint64_t m = (uint64_t)blockIdx.x * PATH_BLOCK + threadIdx.x; m < a.M; m += stride) {
        const uint64_t gidx = a.first + m;
        const uint32_t c0 = (uint32_t)gidx, c1 = (uint32_t)(gidx >> 32);
        T V[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) V[i] = (T)1;
        const T* zrow = SRC == 1 ? a.z_in + m * (uint64_t)a.n_steps * a.n : nullptr;
        for (int s = 0; s < a.n_steps; ++s) {
            T z[NP];
            if (SRC == 1) {
                const T* zp = zrow + (size_t)s * a.n;
                if ((a.n * sizeof(T)) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.z_in) & 15) == 0) {
                    constexpr int V4 = 16 / sizeof(T);
#pragma unroll
                    for (int i = 0; i < NP; i += V4) {
                        if (i < a.n) {
                            const int4 q = __ldg(reinterpret_cast<const int4*>(zp + i));
                            const T* qv = reinterpret_cast<const T*>(&q);
#pragma unroll
                            for (int k = 0; k < V4; ++k) z[i + k] = qv[k];
                        } else {
#pragma unroll
                            for (int k = 0; k < V4; ++k) z[i + k] = (T)0;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NP; ++i) z[i] = i < a.n ? __ldg(zp + i) : (T)0;
                }
            } else {
                draw_normals<T, NP, ROUNDS>(a, c0, c1, (uint32_t)s, z);
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                T r = a.drift[i];
#pragma unroll
                for (int j = 0; j <= i; ++j) r = Math<T>::fma(a.lp[i * (i + 1) / 2 + j], z[j], r);
                V[i] = Math<T>::fma(V[i], r, V[i]);          // V *= (1 + r)
            }
        }
        T x = (T)-1;
#pragma unroll
        for (int i = 0; i < NP; ++i) x = Math<T>::fma(a.w[i], V[i], x);
        a.terminal[m] = x;
    }
}

// Packed FP32x2 variant (Philox mode): a thread carries TWO paths in float2 lanes, so the
// L z mat-vec, the compounding and the Box-Muller scaling issue one FFMA2 per two FMAs.  The
// kernel is issue-bound (FMA / ALU / XU pipes all < 55 % busy), so this is the lever that matters.
template <int NP, int ROUNDS>
__global__ void __launch_bounds__(PATH_BLOCK) path_kernel_packed(const __grid_constant__ PathArgs<float, NP> a) {
    const uint64_t n_sub = (a.M + PATH_BLOCK - 1) / PATH_BLOCK;
    const uint64_t n_tiles = (n_sub + 1) / 2;
    const float kPi = 3.14159265358979323846f;
    const uint32_t one_bits = opaque_u32(0x3f800000u), two_bits = opaque_u32(0x40000000u);
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t mA = tile * (2 * PATH_BLOCK) + threadIdx.x, mB = mA + PATH_BLOCK;
        const uint64_t gA = a.first + mA, gB = a.first + mB;
        float2 V[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) V[i] = make_float2(1.f, 1.f);
        for (int s = 0; s < a.n_steps; ++s) {
            float2 z[NP];
            uint32_t fa[NP], fb[NP];
            philox_fields<NP, ROUNDS>((uint32_t)gA, (uint32_t)(gA >> 32), (uint32_t)s, STREAM_NORMALS, a.k0, a.k1, fa);
            philox_fields<NP, ROUNDS>((uint32_t)gB, (uint32_t)(gB >> 32), (uint32_t)s, STREAM_NORMALS, a.k0, a.k1, fb);
#pragma unroll
            for (int k = 0; k < NP / 2; ++k) {
                const float2 f1 = make_float2(__uint_as_float(mant_or(fa[2 * k], one_bits)), __uint_as_float(mant_or(fb[2 * k], one_bits)));
                const float2 u1 = fma2(f1, bcast2(-1.0f), bcast2(2.0f));                      // (0, 1]
                // r / sqrt(2 ln 2) = sqrt(-lg2 U1): the constant rides in lp (host), the negation in the MUFU operand
                const float2 r = make_float2(Math<float>::sqrt(-Math<float>::lg2(u1.x)), Math<float>::sqrt(-Math<float>::lg2(u1.y)));
                const float2 f2 = make_float2(__uint_as_float(mant_or(fa[2 * k + 1], two_bits)), __uint_as_float(mant_or(fb[2 * k + 1], two_bits)));
                const float2 th = fma2(f2, bcast2(kPi), bcast2(-3.0f * kPi));                 // pi (2f - 1) in [-pi, pi), one rounding
                const float2 cs = make_float2(Math<float>::cosf_(th.x), Math<float>::cosf_(th.y));
                const float2 sn = make_float2(Math<float>::sinf_(th.x), Math<float>::sinf_(th.y));
                z[2 * k] = fma2(r, cs, bcast2(0.0f));
                z[2 * k + 1] = fma2(r, sn, bcast2(0.0f));
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                float2 r = bcast2(a.drift[i]);
#pragma unroll
                for (int j = 0; j <= i; ++j) r = fma2(bcast2(a.lp[i * (i + 1) / 2 + j]), z[j], r);
                V[i] = fma2(V[i], r, V[i]);                                                      // V *= (1 + r)
            }
        }
        float2 x = bcast2(-1.0f);
#pragma unroll
        for (int i = 0; i < NP; ++i) x = fma2(bcast2(a.w[i]), V[i], x);
        if (mA < a.M) a.terminal[mA] = x.x;
        if (mB < a.M) a.terminal[mB] = x.y;
    }
}

// ---------------------------------------------------------------------------------------------
// Wide universes (32 < N <= PATH_WIDE_MAX_N): one WARP per path, lane l owns assets l, l + 32, ...  V lives in the lanes'
// registers (VPL = ceil(N / 32) values each), the step's normals in the warp's shared-memory row, L sqrt(dt) TRANSPOSED in
// global memory (Lt[j][i], i >= j: consecutive lanes read consecutive addresses; N = 256 FP32 is 256 KB, L2 / L1 resident).
// r_i = drift_i + sum_{j <= i} Lt[j][i] z_j.  This is the envelope-widening kernel (a C5-sized universe can be simulated, with
// supplied-normals parity); the throughput path is the tcgen05 kernel for N <= 32.
// ---------------------------------------------------------------------------------------------
constexpr int PW_WARPS = 8;

template <typename T>
struct WideArgs {
    const T* lt;            // [n][np] transposed L sqrt(dt), zero above the diagonal of L (i < j)
    const T* drift;         // [np]
    const T* w;             // [np]
    const T* z_in;          // [M, S, n] or null
    T* terminal;
    int n, np, n_steps;
    uint32_t k0, k1;
    uint64_t first, M;
};

template <typename T, int VPL, int ROUNDS>
__global__ void __launch_bounds__(PW_WARPS * 32) path_kernel_wide(const WideArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sZ = reinterpret_cast<T*>(smem_raw) + (size_t)(threadIdx.x >> 5) * a.np;
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t)gridDim.x * PW_WARPS;
    T drift[VPL], wv[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int i = lane + 32 * v;
        drift[v] = i < a.np ? a.drift[i] : (T)0;
        wv[v] = i < a.np ? a.w[i] : (T)0;
    }
    for (uint64_t m = (uint64_t)blockIdx.x * PW_WARPS + (threadIdx.x >> 5); m < a.M; m += warps_total) {
        const uint64_t gidx = a.first + m;
        const uint32_t c0 = (uint32_t)gidx, c1 = (uint32_t)(gidx >> 32);
        T V[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) V[v] = (T)1;
        for (int s = 0; s < a.n_steps; ++s) {
            __syncwarp();
            if (a.z_in != nullptr) {
                const T* zp = a.z_in + (m * (uint64_t)a.n_steps + (uint64_t)s) * (uint64_t)a.n;
                for (int i = lane; i < a.np; i += 32) sZ[i] = i < a.n ? __ldg(zp + i) : (T)0;
            } else {
                // normals 4q .. 4q+3 of this step: the same fields / words every other path kernel (and oracle/philox_np.py) uses
                for (int q = lane; q < a.np / 4; q += 32) {
                    T z4[4];
                    if constexpr (sizeof(T) == 4) {
                        // 24-bit fields 4q .. 4q+3 = words 3q .. 3q+2 of the stream: blocks (3q)/4 .. (3q+2)/4
                        uint32_t wds[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int wi = 3 * q + k;
                            uint32_t x[4];
                            philox4x32_r<ROUNDS>(c0, c1, (uint32_t)s, STREAM_NORMALS + (uint32_t)(wi >> 2), a.k0, a.k1, x);
                            wds[k] = x[wi & 3];
                        }
                        uint32_t f[4];
                        fields_from_triple(wds[0], wds[1], wds[2], f);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const T u1 = Math<T>::unit_open0(f[2 * k]);
                            const T r = Math<T>::sqrt(Math<T>::lg2(u1) * PathConst<T>::neg2ln2());
                            const T th = PathConst<T>::centred(f[2 * k + 1]) * PathConst<T>::pi();
                            z4[2 * k] = r * Math<T>::cosf_(th);
                            z4[2 * k + 1] = r * Math<T>::sinf_(th);
                        }
                    } else {
                        uint32_t x[4];
                        philox4x32_r<ROUNDS>(c0, c1, (uint32_t)s, STREAM_NORMALS | (uint32_t)q, a.k0, a.k1, x);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const T u1 = Math<T>::unit_open0(x[2 * k]);
                            const T r = Math<T>::sqrt(Math<T>::lg2(u1) * PathConst<T>::neg2ln2());
                            double sn, cs;
                            ::sincospi((double)PathConst<T>::centred(x[2 * k + 1]), &sn, &cs);
                            z4[2 * k] = r * (T)cs;
                            z4[2 * k + 1] = r * (T)sn;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) sZ[4 * q + k] = 4 * q + k < a.n ? z4[k] : (T)0;
                }
            }
            __syncwarp();
            T r[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) r[v] = drift[v];
            for (int j = 0; j < a.n; ++j) {
                const T zj = sZ[j];
                const T* col = a.lt + (size_t)j * a.np + lane;
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    if (32 * v + 31 >= j) r[v] = Math<T>::fma(col[32 * v], zj, r[v]);      // Lt[j][i] = 0 for i < j (stored), rows fully below j skipped
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) V[v] = Math<T>::fma(V[v], r[v], V[v]);
        }
        T x = (T)0;
#pragma unroll
        for (int v = 0; v < VPL; ++v) x = Math<T>::fma(wv[v], V[v], x);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) x += shfl_xor<T>(x, d);
        if (lane == 0) a.terminal[m] = x - (T)1;
    }
}

// host: lower Cholesky factor (FP64), row-major n x n; returns false if not positive definite
static bool cholesky_lower(const double* sigma, int n, std::vector<double>& L) {
    L.assign((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        double d = sigma[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
        if (!(d > 0.0) || !std::isfinite(d)) return false;
        const double djj = std::sqrt(d);
        L[(size_t)j * n + j] = djj;
        for (int i = j + 1; i < n; ++i) {
            double v = sigma[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) v -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
            L[(size_t)i * n + j] = v / djj;
        }
    }
    return true;
}

template <typename T, int NP>
static int path_launch_t(mcp_context* h, const PathJob& job) {
    PathArgs<T, NP> a;
    const int n = job.n;
    const std::vector<double>& L = *job.L;
    const double sdt = std::sqrt(job.dt);
    for (int i = 0; i < NP; ++i) {
        for (int j = 0; j <= i; ++j) a.lp[i * (i + 1) / 2 + j] = (i < n && j < n) ? (T)(L[(size_t)i * n + j] * sdt) : (T)0;
        a.drift[i] = i < n ? (T)(job.mu[i] * job.dt) : (T)0;
        a.w[i] = i < n ? (T)job.w[i] : (T)0;
    }
    a.n = n;
    a.n_steps = job.n_steps;
    a.k0 = (uint32_t)job.seed;
    a.k1 = (uint32_t)(job.seed >> 32);
    a.first = job.first;
    a.M = job.M;
    a.z_in = (const T*)job.z_in;
    a.terminal = (T*)job.terminal;
    const bool r7 = job.rounds == 7;
    void (*kern)(PathArgs<T, NP>) = job.z_in ? path_kernel<T, NP, 1, 10> : (r7 ? path_kernel<T, NP, 0, 7> : path_kernel<T, NP, 0, 10>);
    int per_thread = 1;
    if constexpr (sizeof(T) == 4 && NP <= 16) {
        if (!job.z_in) {
            // the packed kernel's normals come out divided by sqrt(2 ln 2) (its Box-Muller radius is sqrt(-lg2 U1))
            kern = r7 ? path_kernel_packed<NP, 7> : path_kernel_packed<NP, 10>;
            per_thread = 2;
            const double c = std::sqrt(2.0 * std::log(2.0));
            for (int i = 0; i < NP; ++i)
                for (int j = 0; j <= i; ++j) a.lp[i * (i + 1) / 2 + j] = (i < n && j < n) ? (T)(L[(size_t)i * n + j] * sdt * c) : (T)0;
        }
    }
    int per_sm = 0;
    MCP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PATH_BLOCK, 0));
    if (per_sm < 1) return mcp_fail(h, MCP_ERR_CUDA, "path_kernel<N=%d>: zero occupancy", NP);
    uint64_t grid = (uint64_t)h->prop.multiProcessorCount * per_sm;
    const uint64_t need = (job.M + (uint64_t)PATH_BLOCK * per_thread - 1) / ((uint64_t)PATH_BLOCK * per_thread);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, PATH_BLOCK, 0, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <typename T, int VPL>
static int path_launch_wide_v(mcp_context* h, const PathJob& job, const WideArgs<T>& a) {
    auto kern = job.rounds == 7 ? path_kernel_wide<T, VPL, 7> : path_kernel_wide<T, VPL, 10>;
    const size_t smem = (size_t)PW_WARPS * a.np * sizeof(T);
    if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MCP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PW_WARPS * 32, smem));
    if (per_sm < 1) return mcp_fail(h, MCP_ERR_CUDA, "path_kernel_wide: zero occupancy (smem %zu B)", smem);
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount * per_sm, (job.M + PW_WARPS - 1) / PW_WARPS);
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, PW_WARPS * 32, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <typename T>
static int path_launch_wide(mcp_context* h, const PathJob& job) {
    const int n = job.n, np = (n + 31) / 32 * 32;
    const std::vector<double>& L = *job.L;
    const double sdt = std::sqrt(job.dt);
    std::vector<T> host((size_t)n * np + 2 * np, (T)0);
    for (int j = 0; j < n; ++j)
        for (int i = j; i < n; ++i) host[(size_t)j * np + i] = (T)(L[(size_t)i * n + j] * sdt);
    T* hd = host.data() + (size_t)n * np;
    T* hw = hd + np;
    for (int i = 0; i < n; ++i) { hd[i] = (T)(job.mu[i] * job.dt); hw[i] = (T)job.w[i]; }
    T* dev = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 6, host.size() * sizeof(T), (void**)&dev));
    ++h->const_epoch;
    MCP_CUDA(h, cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, job.stream));
    MCP_CUDA(h, cudaStreamSynchronize(job.stream));        // `host` is pageable and dies at scope exit
    WideArgs<T> a;
    a.lt = dev; a.drift = dev + (size_t)n * np; a.w = a.drift + np;
    a.z_in = (const T*)job.z_in; a.terminal = (T*)job.terminal;
    a.n = n; a.np = np; a.n_steps = job.n_steps;
    a.k0 = (uint32_t)job.seed; a.k1 = (uint32_t)(job.seed >> 32);
    a.first = job.first; a.M = job.M;
    const int vpl = np / 32;
#define MCP_PW(V) if (vpl <= V) return path_launch_wide_v<T, V>(h, job, a);
    MCP_PW(2) MCP_PW(4) MCP_PW(8) MCP_PW(16) MCP_PW(32)
#undef MCP_PW
    return mcp_fail(h, MCP_ERR_INVALID, "mcp_paths: n_assets=%d exceeds the supported maximum of %d", n, PATH_WIDE_MAX_N);
}

template <typename T>
static int path_dispatch(mcp_context* h, const PathJob& job) {
    const int n = job.n;
    if (n <= 4) return path_launch_t<T, 4>(h, job);
    if (n <= 8) return path_launch_t<T, 8>(h, job);
    if (n <= 16) return path_launch_t<T, 16>(h, job);
    if (n <= 24) return path_launch_t<T, 24>(h, job);
    if (n <= 32) return path_launch_t<T, 32>(h, job);
    return path_launch_wide<T>(h, job);
}

static int path_run(mcp_context* h, PathJob& job) {
    job.hist0_filled = false;
    if (path_tc_eligible(job)) return path_launch_tc(h, job);
    return job.dtype == MCP_F64 ? path_dispatch<double>(h, job) : path_dispatch<float>(h, job);
}

}  // namespace mcp

using namespace mcp;

static int paths_check(mcp_handle h, const mcp_path_params* p, const double* mu, const double* sigma, const double* weights) {
    MCP_REQUIRE(h, p && mu && sigma && weights, "mcp_paths: NULL argument");
    MCP_REQUIRE(h, p->n_assets >= 1 && p->n_assets <= PATH_WIDE_MAX_N, "mcp_paths: n_assets=%d out of range [1, %d]", p->n_assets, PATH_WIDE_MAX_N);
    MCP_REQUIRE(h, p->dtype == MCP_F32 || p->dtype == MCP_F64, "mcp_paths: bad dtype %d", p->dtype);
    MCP_REQUIRE(h, p->space == MCP_HOST || p->space == MCP_DEVICE, "mcp_paths: bad space %d", p->space);
    MCP_REQUIRE(h, p->n_steps >= 1, "mcp_paths: n_steps must be >= 1");
    MCP_REQUIRE(h, p->dt > 0 && std::isfinite(p->dt), "mcp_paths: dt must be positive");
    MCP_REQUIRE(h, p->philox_rounds == 0 || p->philox_rounds == 10 || p->philox_rounds == 7,
                "mcp_paths: philox_rounds=%d (supported: 10 = default, 7)", p->philox_rounds);
    return MCP_OK;
}

// Runs the path kernel of one shard; terminal values end up at *term_dev_out (device).  Leaves ev[2] / ev[3] around the kernel.
static int paths_launch(mcp_handle h, const mcp_path_params* p, const double* mu, const double* sigma, const double* weights,
                        void* terminal_out, bool terminal_optional, unsigned long long* hist0, void** term_dev_out, bool* hist0_filled) {
    const int n = p->n_assets;
    std::vector<double> L;
    if (!cholesky_lower(sigma, n, L))
        return mcp_fail(h, MCP_ERR_NUMERIC, "mcp_paths: cov_matrix is not positive definite (Cholesky failed)");
    const size_t es = p->dtype == MCP_F64 ? 8 : 4;
    cudaStream_t st = h->stream;
    const void* z_dev = p->normals_in;
    void* term_dev = terminal_out;
    if (p->space == MCP_HOST) {
        if (p->normals_in) {
            const size_t zb = (size_t)p->n_paths * p->n_steps * n * es;
            void* d = nullptr;
            MCP_CHECK(mcp_dev_reserve(h, 1, zb, &d));
            MCP_CUDA(h, cudaMemcpyAsync(d, p->normals_in, zb, cudaMemcpyHostToDevice, st));
            z_dev = d;
        }
        MCP_CHECK(mcp_dev_reserve(h, 3, (size_t)p->n_paths * es, &term_dev));
    } else if (!terminal_out) {
        if (!terminal_optional) return mcp_fail(h, MCP_ERR_INVALID, "mcp_paths: terminal_out is NULL");
        MCP_CHECK(mcp_dev_reserve(h, 16, (size_t)p->n_paths * es, &term_dev));
    }
    PathJob job;
    job.n = n; job.n_steps = p->n_steps; job.dtype = p->dtype; job.rounds = p->philox_rounds == 7 ? 7 : 10;
    job.seed = p->seed; job.first = p->first_index; job.M = p->n_paths; job.dt = p->dt;
    job.mu = mu; job.w = weights; job.L = &L;
    job.z_in = z_dev; job.terminal = term_dev; job.hist0 = hist0; job.stream = st;
    MCP_CUDA(h, cudaEventRecord(h->ev[2], st));
    MCP_CHECK(path_run(h, job));
    MCP_CUDA(h, cudaEventRecord(h->ev[3], st));
    if (p->space == MCP_HOST && terminal_out)
        MCP_CUDA(h, cudaMemcpyAsync(terminal_out, term_dev, (size_t)p->n_paths * es, cudaMemcpyDeviceToHost, st));
    *term_dev_out = term_dev;
    if (hist0_filled) *hist0_filled = job.hist0_filled;
    return MCP_OK;
}

static int paths_impl(mcp_handle h, const mcp_path_params* p, const double* mu, const double* sigma,
                      const double* weights, void* terminal_out, double* kernel_ms) {
    MCP_CHECK(paths_check(h, p, mu, sigma, weights));
    MCP_REQUIRE(h, terminal_out, "mcp_paths: terminal_out is NULL");
    mcp_device_guard guard(h->device);
    if (kernel_ms) *kernel_ms = 0;
    if (p->n_paths == 0) return MCP_OK;
    void* term_dev = nullptr;
    MCP_CHECK(paths_launch(h, p, mu, sigma, weights, terminal_out, false, nullptr, &term_dev, nullptr));
    MCP_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0;
    MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
    h->last_ms = ms;
    if (kernel_ms) *kernel_ms = ms;
    return MCP_OK;
}

// paths + VaR / CVaR in one call: the kernel's epilogue fills the first radix histogram, everything else is stream-ordered
static int paths_stats_impl(mcp_handle h, const mcp_path_params* p, const double* mu, const double* sigma,
                            const double* weights, void* terminal_out, mcp_path_stats* stats) {
    MCP_CHECK(paths_check(h, p, mu, sigma, weights));
    MCP_REQUIRE(h, stats != nullptr, "mcp_paths_stats: stats is NULL");
    MCP_REQUIRE(h, stats->n_alphas >= 1 && stats->n_alphas <= MCP_MAX_ALPHAS, "mcp_paths_stats: n_alphas=%d out of range [1, %d]", stats->n_alphas, MCP_MAX_ALPHAS);
    MCP_REQUIRE(h, !stats->comm_merge || h->comm != nullptr, "mcp_paths_stats: comm_merge needs a communicator on this handle (mcp_comm_init)");
    mcp_device_guard guard(h->device);
    stats->kernel_ms = stats->quantile_ms = 0;
    for (int a = 0; a < MCP_MAX_ALPHAS; ++a) stats->var[a] = stats->cvar[a] = NAN;
    const uint64_t n_total = stats->comm_merge ? stats->n_total : p->n_paths;
    MCP_REQUIRE(h, n_total >= 1 && n_total >= p->n_paths, "mcp_paths_stats: n_total=%llu is smaller than this shard (or zero)", (unsigned long long)n_total);
    cudaStream_t st = h->stream;
    unsigned long long* hist0 = nullptr;
    bool filled = false;
    void* term_dev = nullptr;
    if (p->n_paths > 0) {
        if (p->dtype == MCP_F32) {
            MCP_CHECK(mcp_dev_reserve(h, 17, sizeof(unsigned long long) << MCP_SEL_BITS, (void**)&hist0));
            MCP_CUDA(h, cudaMemsetAsync(hist0, 0, sizeof(unsigned long long) << MCP_SEL_BITS, st));
        }
        MCP_CHECK(paths_launch(h, p, mu, sigma, weights, terminal_out, true, hist0, &term_dev, &filled));
    }
    double qms = 0;
    MCP_CHECK(mcp_quantiles_device(h, term_dev, p->dtype, p->n_paths, n_total, stats->alphas, stats->n_alphas, stats->var, stats->cvar,
                                   stats->comm_merge ? MCP_ALLREDUCE_COMM : nullptr, nullptr, filled ? hist0 : nullptr, &qms));
    if (p->n_paths > 0) {
        float ms = 0;
        MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
        stats->kernel_ms = ms;
    }
    stats->quantile_ms = qms;
    h->last_ms = stats->kernel_ms;
    return MCP_OK;
}

extern "C" int mcp_paths(mcp_handle h, const mcp_path_params* p, const double* mu, const double* sigma,
                         const double* weights, void* terminal_out, double* kernel_ms) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_paths", [&] { return paths_impl(h, p, mu, sigma, weights, terminal_out, kernel_ms); });
}

extern "C" int mcp_paths_stats(mcp_handle h, const mcp_path_params* p, const double* mu, const double* sigma,
                               const double* weights, void* terminal_out, mcp_path_stats* stats) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_paths_stats", [&] { return paths_stats_impl(h, p, mu, sigma, weights, terminal_out, stats); });
}
