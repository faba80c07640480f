// Device-side building blocks shared by the sm_100a kernels: Philox4x32-10, the uniform ->
// exponential / normal transforms, MUFU-level math wrappers and (key, index) reductions.
// The generator spec (counter layout, bit -> float construction) is restated for the tests in
// oracle/philox_np.py; keep the two in sync.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mcp_lg2_table.cuh"

namespace mcp {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t STREAM_WEIGHTS = 1u << 24;
constexpr uint32_t STREAM_NORMALS = 2u << 24;

// Philox4x32-10 (Salmon et al., SC'11).  The key schedule k + r*W is uniform across the
// grid (seed is a kernel parameter), so it lives in uniform registers; each round costs two
// IMAD.WIDE.U32 and two LOP3 per thread.
// The round count is a template parameter: 10 is the default everywhere (Random123's and cuRAND's choice, and the generator
// oracle/philox_np.py restates); 7 is Random123's documented minimum that still passes BigCrush ("Crush-resistant"), offered as an
// option (mcp_portfolio_params.philox_rounds) -- it is a different stream, not a cheaper way to the same numbers.
template <int ROUNDS = 10>
__device__ __forceinline__ void philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                             uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
    philox4x32_r<10>(c0, c1, c2, c3, k0, k1, out);
}

// The same generator with the ten round keys precomputed on the host (kernel parameter block -> constant-bank LOP3 operands).
struct PhiloxKeys {
    uint32_t k[20];                      // k[2r] = k0 + r W0, k[2r+1] = k1 + r W1
};
inline void philox_keys_fill(PhiloxKeys& pk, uint64_t seed) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { pk.k[2 * r] = k0; pk.k[2 * r + 1] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
}
template <int ROUNDS = 10>
__device__ __forceinline__ void philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& pk, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ pk.k[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ pk.k[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& pk, uint32_t (&out)[4]) {
    philox4x32_r<10>(c0, c1, c2, c3, pk, out);
}

// FP32 streams spend 24 bits per uniform instead of a whole 32-bit word: uniform field i of a row is
// bits [24 i, 24 i + 23) of the concatenation of its stream's Philox blocks 0, 1, 2, ... (word 0 of block 0
// lowest).  Four fields share three words, so 16 uniforms cost 3 Philox calls instead of 4 (+11 % on the
// N = 16 sweep); the price is two funnel shifts and a shift per four fields.  The FP64 kernels keep one
// 32-bit word per uniform.  `f` holds the fields UNMASKED: unit_open0 / unit_frac / centred mask the low 23 bits.
__device__ __forceinline__ void fields_from_triple(uint32_t a, uint32_t b, uint32_t c, uint32_t* f) {
    f[0] = a;
    f[1] = __funnelshift_r(a, b, 24);
    f[2] = __funnelshift_r(b, c, 16);
    f[3] = c >> 8;
}
__host__ __device__ constexpr int philox_blocks_for_fields(int nf) { return (3 * ((nf + 3) / 4) + 3) / 4; }

// NF fields (multiple of 4) that start at a block boundary: field 0 = bit 0 of block `c3 & 0xffffff`.
template <int NF, int ROUNDS = 10>
__device__ __forceinline__ void philox_fields(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&f)[NF]) {
    static_assert(NF % 4 == 0, "fields come in groups of four");
    constexpr int NB = philox_blocks_for_fields(NF);
    uint32_t w[4 * NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint32_t x[4];
        philox4x32_r<ROUNDS>(c0, c1, c2, c3 + (uint32_t)b, k0, k1, x);
#pragma unroll
        for (int k = 0; k < 4; ++k) w[4 * b + k] = x[k];
    }
#pragma unroll
    for (int t = 0; t < NF / 4; ++t) fields_from_triple(w[3 * t], w[3 * t + 1], w[3 * t + 2], &f[4 * t]);
}
template <int NF, int ROUNDS = 10>
__device__ __forceinline__ void philox_fields(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& pk, uint32_t (&f)[NF]) {
    static_assert(NF % 4 == 0, "fields come in groups of four");
    constexpr int NB = philox_blocks_for_fields(NF);
    uint32_t w[4 * NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint32_t x[4];
        philox4x32_r<ROUNDS>(c0, c1, c2, c3 + (uint32_t)b, pk, x);
#pragma unroll
        for (int k = 0; k < 4; ++k) w[4 * b + k] = x[k];
    }
#pragma unroll
    for (int t = 0; t < NF / 4; ++t) fields_from_triple(w[3 * t], w[3 * t + 1], w[3 * t + 2], &f[4 * t]);
}

// ---- math wrappers ---------------------------------------------------------------------
// FP32 uses the MUFU approximations directly (lg2 / rcp / rsqrt / sin / cos: <= 2 ulp-class
// error, far inside the 1e-4 parity tolerance); FP64 uses the IEEE library routines.
template <typename T> struct Math;

template <> struct Math<float> {
    static __device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float sinf_(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float cosf_(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    // U in (0, 1] with 23 random bits: float in [1, 2) built with one LOP3, then 2 - f.
    static __device__ __forceinline__ float unit_open0(uint32_t x) { return 2.0f - __uint_as_float((x & 0x007fffffu) | 0x3f800000u); }
    // fraction in [0, 1) with 23 random bits
    static __device__ __forceinline__ float unit_frac(uint32_t x) { return __uint_as_float((x & 0x007fffffu) | 0x3f800000u) - 1.0f; }
};

template <> struct Math<double> {
    // log2 on the generator's domain u in [2^-32, 1] (NOT a general log2): u = 2^k m, table entry by the top 10 mantissa
    // bits of m (mcp_lg2_table.cuh, 16 KB, L1-resident), r = fma(m, inv, -1) exact with |r| < 2^-10, degree-5 Taylor polynomial
    // of log2(1 + r) (truncation < 3e-19).
    // Absolute error < 4e-15 (half an ulp of the largest results), exactly 0 at u = 1; 6 DFMA instead of libdevice's ~30.
    static __device__ __forceinline__ double lg2(double u) {
        const long long b = __double_as_longlong(u);
        const int k = (int)(b >> 52) - 1023;
        const double m = __longlong_as_double((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
        const double2 t = MCP_LG2_TAB[(int)(b >> 42) & 1023];
        const double r = ::fma(m, t.x, -1.0);
        const double c[5] = {1.4426950408889634, -0.7213475204444817, 0.48089834696298783, -0.36067376022224085, 0.28853900817779266};
        double p = c[4];
#pragma unroll
        for (int j = 3; j >= 0; --j) p = ::fma(p, r, c[j]);
        return ::fma(p, r, (double)k + t.y);
    }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double sinf_(double x) { return ::sin(x); }
    static __device__ __forceinline__ double cosf_(double x) { return ::cos(x); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double unit_open0(uint32_t x) { return 1.0 - (double)x * 0x1p-32; }
    static __device__ __forceinline__ double unit_frac(uint32_t x) { return (double)x * 0x1p-32; }
};

// (x & 0x007fffff) | expo as ONE LOP3: written as `(x & m) | c` with two literals, ptxas often emits an AND and an OR (one
// 32-bit immediate per instruction).  `expo` comes from opaque_u32(), a register the optimiser cannot fold back into a literal.
// The sweeps are issue-bound: this is 4-5 % of their instruction stream.
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
    uint32_t r;
    asm("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t mant_or(uint32_t x, uint32_t expo) {
    uint32_t d;
    asm("lop3.b32 %0, %1, 0x007fffff, %2, 0xEA;" : "=r"(d) : "r"(x), "r"(expo));
    return d;
}

// ---- packed FP32x2 (Blackwell FFMA2): one instruction = two FMAs, operands may be register
// pairs, uniform-register pairs (constant bank) or a broadcast scalar.  The fused sweep is
// issue-bound, so halving the FFMA count is worth more than any pipe-level trick.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 bcast2(float x) { return make_float2(x, x); }

// ---- (key, index) argmax with first-occurrence tie-break ---------------------------------
// `better(a, ia, b, ib)`: does candidate a beat b?  Larger key wins; equal keys -> lower index.
template <typename T>
struct Cand {
    T key;
    uint64_t idx;
};

template <typename T>
__device__ __forceinline__ bool cand_better(T ka, uint64_t ia, T kb, uint64_t ib) {
    return (ka > kb) || (ka == kb && ia < ib);
}

template <typename T> __device__ __forceinline__ T shfl_xor(T v, int m);
template <> __device__ __forceinline__ float shfl_xor<float>(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ double shfl_xor<double>(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ uint64_t shfl_xor<uint64_t>(uint64_t v, int m) { return __shfl_xor_sync(0xffffffffu, (unsigned long long)v, m); }

template <typename T>
__device__ __forceinline__ void warp_argmax(T& key, uint64_t& idx) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const T ok = shfl_xor<T>(key, m);
        const uint64_t oi = shfl_xor<uint64_t>(idx, m);
        if (cand_better<T>(ok, oi, key, idx)) { key = ok; idx = oi; }
    }
}

template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { const T o = shfl_xor<T>(v, m); v = o < v ? o : v; }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { const T o = shfl_xor<T>(v, m); v = o > v ? o : v; }
    return v;
}

// metrics from un-normalised weights (s = their sum; s = 1 for supplied weights)
template <typename T>
__device__ __forceinline__ void metrics_from(T q, T r, T s, T rf, bool normalised, T& ret, T& risk, T& sharpe) {
    if (normalised) {
        ret = r;
        risk = Math<T>::sqrt(q);
        sharpe = risk > (T)0 ? (r - rf) * Math<T>::rcp(risk) : (T)0;
    } else {
        const T inv = Math<T>::rcp(s);
        const T rs = Math<T>::rsqrt(q);
        ret = r * inv;
        risk = q * rs * inv;                       // sqrt(q) / s
        sharpe = q > (T)0 ? (r - rf * s) * rs : (T)0;   // (ret - rf) / risk
        if (!(q > (T)0)) risk = (T)0;
    }
}

// order-preserving float -> unsigned key (ascending), and back
__device__ __host__ __forceinline__ uint32_t f32_to_key(uint32_t b) { return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __host__ __forceinline__ uint32_t key_to_f32(uint32_t k) { return k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu); }
__device__ __host__ __forceinline__ uint64_t f64_to_key(uint64_t b) { return b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull); }
__device__ __host__ __forceinline__ uint64_t key_to_f64(uint64_t k) { return k ^ ((k >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull); }

}  // namespace mcp
