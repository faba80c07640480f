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

#include "mcp_context.h"
#include "mcp_device.cuh"

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

template <typename T, int NP>
__device__ __forceinline__ void draw_normals(const PathArgs<T, NP>& a, uint32_t c0, uint32_t c1, uint32_t step, T (&z)[NP]) {
    if constexpr (sizeof(T) == 4) {           // FP32: 24-bit fields (mcp_device.cuh), pair (2k, 2k+1) -> normals (2k, 2k+1)
        uint32_t f[NP];
        philox_fields<NP>(c0, c1, step, STREAM_NORMALS, a.k0, a.k1, f);
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
            philox4x32_10(c0, c1, step, STREAM_NORMALS | (uint32_t)b, a.k0, a.k1, x);
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

template <typename T, int NP, int SRC>
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
                draw_normals<T, NP>(a, c0, c1, (uint32_t)s, z);
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
template <int NP>
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
            philox_fields<NP>((uint32_t)gA, (uint32_t)(gA >> 32), (uint32_t)s, STREAM_NORMALS, a.k0, a.k1, fa);
            philox_fields<NP>((uint32_t)gB, (uint32_t)(gB >> 32), (uint32_t)s, STREAM_NORMALS, a.k0, a.k1, fb);
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
static int path_launch_t(mcp_context* h, const mcp_path_params* p, const double* mu, const std::vector<double>& L,
                         const double* w, const void* z_dev, void* term_dev, cudaStream_t st) {
    PathArgs<T, NP> a;
    const int n = p->n_assets;
    const double sdt = std::sqrt(p->dt);
    for (int i = 0; i < NP; ++i) {
        for (int j = 0; j <= i; ++j) a.lp[i * (i + 1) / 2 + j] = (i < n && j < n) ? (T)(L[(size_t)i * n + j] * sdt) : (T)0;
        a.drift[i] = i < n ? (T)(mu[i] * p->dt) : (T)0;
        a.w[i] = i < n ? (T)w[i] : (T)0;
    }
    a.n = n;
    a.n_steps = p->n_steps;
    a.k0 = (uint32_t)p->seed;
    a.k1 = (uint32_t)(p->seed >> 32);
    a.first = p->first_index;
    a.M = p->n_paths;
    a.z_in = (const T*)z_dev;
    a.terminal = (T*)term_dev;
    void (*kern)(PathArgs<T, NP>) = z_dev ? path_kernel<T, NP, 1> : path_kernel<T, NP, 0>;
    int per_thread = 1;
    if constexpr (sizeof(T) == 4 && NP <= 16) {
        if (!z_dev) {
            // the packed kernel's normals come out divided by sqrt(2 ln 2) (its Box-Muller radius is sqrt(-lg2 U1))
            kern = path_kernel_packed<NP>;
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
    const uint64_t need = (p->n_paths + (uint64_t)PATH_BLOCK * per_thread - 1) / ((uint64_t)PATH_BLOCK * per_thread);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, PATH_BLOCK, 0, st>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <typename T>
static int path_dispatch(mcp_context* h, const mcp_path_params* p, const double* mu, const std::vector<double>& L,
                         const double* w, const void* z_dev, void* term_dev, cudaStream_t st) {
    const int n = p->n_assets;
    if (n <= 4) return path_launch_t<T, 4>(h, p, mu, L, w, z_dev, term_dev, st);
    if (n <= 8) return path_launch_t<T, 8>(h, p, mu, L, w, z_dev, term_dev, st);
    if (n <= 16) return path_launch_t<T, 16>(h, p, mu, L, w, z_dev, term_dev, st);
    if (n <= 24) return path_launch_t<T, 24>(h, p, mu, L, w, z_dev, term_dev, st);
    return path_launch_t<T, 32>(h, p, mu, L, w, z_dev, term_dev, st);
}

}  // namespace mcp

using namespace mcp;

static int paths_impl(mcp_handle h, const mcp_path_params* p, const double* mu, const double* sigma,
                      const double* weights, void* terminal_out, double* kernel_ms) {
    MCP_REQUIRE(h, p && mu && sigma && weights && terminal_out, "mcp_paths: NULL argument");
    MCP_REQUIRE(h, p->n_assets >= 1 && p->n_assets <= 32, "mcp_paths: n_assets=%d out of range [1, 32]", p->n_assets);
    MCP_REQUIRE(h, p->dtype == MCP_F32 || p->dtype == MCP_F64, "mcp_paths: bad dtype %d", p->dtype);
    MCP_REQUIRE(h, p->space == MCP_HOST || p->space == MCP_DEVICE, "mcp_paths: bad space %d", p->space);
    MCP_REQUIRE(h, p->n_steps >= 1, "mcp_paths: n_steps must be >= 1");
    MCP_REQUIRE(h, p->dt > 0 && std::isfinite(p->dt), "mcp_paths: dt must be positive");
    mcp_device_guard guard(h->device);
    if (kernel_ms) *kernel_ms = 0;
    if (p->n_paths == 0) return MCP_OK;
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
    }
    MCP_CUDA(h, cudaEventRecord(h->ev[0], st));
    if (p->dtype == MCP_F64) MCP_CHECK(path_dispatch<double>(h, p, mu, L, weights, z_dev, term_dev, st));
    else MCP_CHECK(path_dispatch<float>(h, p, mu, L, weights, z_dev, term_dev, st));
    MCP_CUDA(h, cudaEventRecord(h->ev[1], st));
    if (p->space == MCP_HOST)
        MCP_CUDA(h, cudaMemcpyAsync(terminal_out, term_dev, (size_t)p->n_paths * es, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));
    float ms = 0;
    MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->last_ms = ms;
    if (kernel_ms) *kernel_ms = ms;
    return MCP_OK;
}

extern "C" int mcp_paths(mcp_handle h, const mcp_path_params* p, const double* mu, const double* sigma,
                         const double* weights, void* terminal_out, double* kernel_ms) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_paths", [&] { return paths_impl(h, p, mu, sigma, weights, terminal_out, kernel_ms); });
}
