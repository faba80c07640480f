// Fused portfolio sweep for N <= 32 assets: one thread per portfolio, everything in registers.
//
// Replaces app.py:699-717 (sampling, bounds rejection, return / risk / Sharpe), 719-722
// (optional write-back) and the argmax at 672/747 plus the nearest-risk pick, in ONE kernel:
//   * weights are generated in-register from Philox4x32-10 (counter = global portfolio index,
//     attempt, 4-asset block) as normalised base-2 exponentials (flat Dirichlet, app.py:702);
//   * Sigma (packed lower triangle, off-diagonals pre-doubled) and mu are kernel parameters,
//     i.e. they sit in the constant bank and feed the FFMAs as immediate-like operands -- no
//     shared-memory or global loads in the inner loop, zero HBM traffic without write-back;
//   * the quadratic form runs on the un-normalised exponentials (q_e = e' S e, r_e = e.mu) and
//     is scaled once: ret = r_e / s, risk = sqrt(q_e) / s, sharpe = (r_e - rf s) rsqrt(q_e);
//   * per-thread running (max Sharpe, min |risk - target|) -> warp shuffles -> one candidate
//     per CTA; a second tiny kernel merges the CTA candidates (first-occurrence tie-break).
// Supplied-weights mode stages each warp's 32 rows through padded shared memory so global
// loads / stores are fully coalesced while each thread still owns one row.
#pragma once
#include <cstdlib>

#include "mcp_device.cuh"
#include "mcp_portfolio.h"

namespace mcp {

template <typename T, int NP>
struct SmallArgs {
    T sig[NP * (NP + 1) / 2];   // row i holds j = 0..i; off-diagonals doubled
    T mu[NP];
    T mask[NP];                 // 1 for real assets, 0 for padded ones
    T nmu[NP], nmask[NP];       // -mu, -mask: the packed path works on l = lg2(U) = -e
    T lo[NP], hi[NP];
    T rf, target;
    int n;                      // real asset count (<= NP); padded assets carry weight 0
    int max_tries, keep_last;
    uint32_t k0, k1;            // Philox key = seed
    PhiloxKeys rk;              // its ten round keys (packed kernel: constant-bank LOP3 operands instead of 18 UIADD3 per tile)
    uint64_t first, P;
    const T* w_in;
    T* w_out;
    T* ret_out;
    T* risk_out;
    T* sharpe_out;
    uint8_t* acc_out;
    PfCand* cands;
    unsigned long long* n_accepted;
};

// One flat-Dirichlet draw: e[i] = -lg2(U_i), s = sum over the real assets (padded assets keep a
// draw but are masked out of s, and their Sigma / mu entries are 0, so they never contribute).
template <typename T, int NP, int ROUNDS>
__device__ __forceinline__ void draw_exponentials(const SmallArgs<T, NP>& a, uint32_t c0, uint32_t c1,
                                                  uint32_t attempt, T (&e)[NP], T& s) {
    s = (T)0;
    if constexpr (sizeof(T) == 4) {           // FP32: 24-bit fields, 3 Philox calls per 16 uniforms
        uint32_t f[NP];
        philox_fields<NP, ROUNDS>(c0, c1, attempt, STREAM_WEIGHTS, a.k0, a.k1, f);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            e[i] = -Math<T>::lg2(Math<T>::unit_open0(f[i]));
            s = Math<T>::fma(e[i], a.mask[i], s);
        }
    } else {
#pragma unroll
        for (int b = 0; b < NP / 4; ++b) {
            uint32_t x[4];
            philox4x32_r<ROUNDS>(c0, c1, attempt, STREAM_WEIGHTS | (uint32_t)b, a.k0, a.k1, x);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * b + k;
                e[i] = -Math<T>::lg2(Math<T>::unit_open0(x[k]));
                s = Math<T>::fma(e[i], a.mask[i], s);
            }
        }
    }
}

// q = e' Sigma e over the packed, pre-doubled lower triangle; r = e . mu
template <typename T, int NP>
__device__ __forceinline__ void quad_and_dot(const SmallArgs<T, NP>& a, const T (&e)[NP], T& q, T& r) {
    q = (T)0;
    r = (T)0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        T t = (T)0;
#pragma unroll
        for (int j = 0; j <= i; ++j) t = Math<T>::fma(a.sig[i * (i + 1) / 2 + j], e[j], t);
        q = Math<T>::fma(e[i], t, q);
        r = Math<T>::fma(a.mu[i], e[i], r);
    }
}

template <typename T, int NP>
__device__ __forceinline__ bool in_bounds(const SmallArgs<T, NP>& a, const T (&e)[NP], T inv) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const T w = e[i] * inv;
        ok = ok && ((i >= a.n) || (w >= a.lo[i] && w <= a.hi[i]));
    }
    return ok;
}

// Draw with bounds rejection (app.py:700-707): returns accepted?; e / s hold the last draw.
template <typename T, int NP, bool BOUNDS, int ROUNDS>
__device__ __forceinline__ bool draw_accepted(const SmallArgs<T, NP>& a, uint64_t gidx, T (&e)[NP], T& s) {
    const uint32_t c0 = (uint32_t)gidx, c1 = (uint32_t)(gidx >> 32);
    if (!BOUNDS) {
        draw_exponentials<T, NP, ROUNDS>(a, c0, c1, 0u, e, s);
        return true;
    }
    bool ok = false;
    for (int t = 0; t < a.max_tries && !ok; ++t) {
        draw_exponentials<T, NP, ROUNDS>(a, c0, c1, (uint32_t)t, e, s);
        ok = in_bounds<T, NP>(a, e, Math<T>::rcp(s));
    }
    return ok || (a.keep_last != 0);
}

// ---- warp-tile staging with 128-bit accesses (float, n % 4 == 0, 16-byte aligned base) ----
// A warp owns 32 consecutive rows = 32*n contiguous floats in global memory.  Shared-memory row
// stride n + 4 floats keeps rows 16-byte aligned and makes the per-thread row accesses
// (LDS.128 / STS.128, one quarter-warp per phase) bank-conflict-free.
// compile-time row length (the common n == NP case): 32 / (N/4) rows per step, fixed column,
// immediate offsets -- one LDS.128 + one STG.128 per step and nothing else
template <int N>
__device__ __forceinline__ void warp_tile_store_vec_ct(const float* stage, float* dst, int rows, int lane) {
    constexpr int UPR = N / 4, VS = N + 4, RPS = 32 / UPR;       // N in {4, 8, 16, 32}: UPR divides 32
    const int r0 = lane / UPR, c = lane % UPR;
    const float* src = stage + r0 * VS + 4 * c;
    float4* out = reinterpret_cast<float4*>(dst) + lane;
#pragma unroll
    for (int k = 0; k < UPR; ++k)
        if (r0 + k * RPS < rows) __stcs(out + 32 * k, *reinterpret_cast<const float4*>(src + k * RPS * VS));
}
template <int N>
__device__ __forceinline__ void warp_tile_load_vec_ct(float* stage, const float* src, int rows, int lane) {
    constexpr int UPR = N / 4, VS = N + 4, RPS = 32 / UPR;
    const int r0 = lane / UPR, c = lane % UPR;
    float* dstp = stage + r0 * VS + 4 * c;
    const float4* in = reinterpret_cast<const float4*>(src) + lane;
#pragma unroll
    for (int k = 0; k < UPR; ++k)
        if (r0 + k * RPS < rows) *reinterpret_cast<float4*>(dstp + k * RPS * VS) = __ldcs(in + 32 * k);
}

__device__ __forceinline__ void warp_tile_store_vec(const float* stage, int vstride, float* dst, int rows, int n, int lane) {
    const int upr = n >> 2;                    // float4 units per row
    const int total = rows * upr;
    int r = lane / upr, c = lane - r * upr;
    const int dr = 32 / upr, dc = 32 - dr * upr;
    for (int u = lane; u < total; u += 32) {
        const float4 v = *reinterpret_cast<const float4*>(stage + r * vstride + 4 * c);
        __stcs(reinterpret_cast<float4*>(dst) + u, v);          // streaming store: written once, never re-read
        r += dr; c += dc;
        if (c >= upr) { c -= upr; ++r; }
    }
}
__device__ __forceinline__ void warp_tile_load_vec(float* stage, int vstride, const float* src, int rows, int n, int lane) {
    const int upr = n >> 2;
    const int total = rows * upr;
    int r = lane / upr, c = lane - r * upr;
    const int dr = 32 / upr, dc = 32 - dr * upr;
    for (int u = lane; u < total; u += 32) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + u);
        *reinterpret_cast<float4*>(stage + r * vstride + 4 * c) = v;
        r += dr; c += dc;
        if (c >= upr) { c -= upr; ++r; }
    }
}

// K portfolios per thread and iteration: every Sigma / mu operand fetched from the constant
// bank (LDCU -> uniform register) feeds K FFMAs, the K Philox / lg2 chains are independent
// (ILP), and the loop / index overhead is paid once per K portfolios.
template <typename T, int NP> struct SweepK { static constexpr int value = (sizeof(T) == 4) ? (NP <= 16 ? 4 : 2) : (NP <= 16 ? 2 : 1); };

template <typename T, int NP, int K>
__device__ __forceinline__ void quad_and_dot_k(const SmallArgs<T, NP>& a, const T (&e)[K][NP], T (&q)[K], T (&r)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) { q[k] = (T)0; r[k] = (T)0; }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        T t[K];
#pragma unroll
        for (int k = 0; k < K; ++k) t[k] = (T)0;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const T sij = a.sig[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < K; ++k) t[k] = Math<T>::fma(sij, e[k][j], t[k]);
        }
        const T mui = a.mu[i];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            q[k] = Math<T>::fma(e[k][i], t[k], q[k]);
            r[k] = Math<T>::fma(mui, e[k][i], r[k]);
        }
    }
}

template <typename T, int NP, int K, int SRC /*0 Philox, 1 supplied*/, bool BOUNDS, int ROUNDS>
__global__ void __launch_bounds__(PF_BLOCK) small_sweep(const __grid_constant__ SmallArgs<T, NP> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // 128-bit staging when rows are float4-tileable (see warp_tile_*_vec), else odd scalar stride
    const bool vec = sizeof(T) == 4 && (a.n & 3) == 0 && (reinterpret_cast<uintptr_t>(a.w_in) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(a.w_out) & 15) == 0;
    const int stride = vec ? a.n + 4 : (a.n | 1);
    T* stage = reinterpret_cast<T*>(smem_raw) + (size_t)warp * 32 * stride;
    const int q32 = 32 / a.n, m32 = 32 % a.n;

    // a "sub-tile" is PF_BLOCK consecutive portfolios (one per thread); a tile is K sub-tiles
    const uint64_t n_sub = (a.P + PF_BLOCK - 1) / PF_BLOCK;
    const uint64_t n_tiles = (n_sub + K - 1) / K;
    constexpr uint32_t NONE = 0xffffffffu;
    T best_s = -Math<T>::inf(), best_d = -Math<T>::inf();
    uint32_t sub_s = NONE, sub_d = NONE;
    T rmin = Math<T>::inf(), rmax = -Math<T>::inf();
    uint32_t n_acc = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        T e[K][NP];
        T s[K];
        bool accepted[K];
        const uint64_t local0 = tile * (uint64_t)(K * PF_BLOCK) + threadIdx.x;      // sub-tile k: + k * PF_BLOCK
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint64_t local = local0 + (uint64_t)k * PF_BLOCK;
            const bool active = local < a.P;
            s[k] = (T)1;
            accepted[k] = active;
            if (SRC == 1) {
                // coalesced warp load of rows*n contiguous values into padded smem, then row -> registers
                const uint64_t warp_row0 = local - lane;
                const int rows = warp_row0 >= a.P ? 0 : (int)((a.P - warp_row0) < 32 ? (a.P - warp_row0) : 32);
                const T* src = a.w_in + warp_row0 * (uint64_t)a.n;
                if constexpr (sizeof(T) == 4) {
                    if (vec) {
                        if constexpr (NP == 4 || NP == 8 || NP == 16 || NP == 32) {
                            if (a.n == NP) warp_tile_load_vec_ct<NP>(reinterpret_cast<float*>(stage), reinterpret_cast<const float*>(src), rows, lane);
                            else warp_tile_load_vec(reinterpret_cast<float*>(stage), stride, reinterpret_cast<const float*>(src), rows, a.n, lane);
                        } else {
                            warp_tile_load_vec(reinterpret_cast<float*>(stage), stride, reinterpret_cast<const float*>(src), rows, a.n, lane);
                        }
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < NP; i += 4) {
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (active && i < a.n) v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(stage) + lane * stride + i);
                            e[k][i] = v.x; e[k][i + 1] = v.y; e[k][i + 2] = v.z; e[k][i + 3] = v.w;
                        }
                        __syncwarp();
                    }
                }
                if (!vec) {
                    const int total = rows * a.n;
                    int rr = lane / a.n, cc = lane % a.n;
                    for (int f = lane; f < total; f += 32) {
                        stage[rr * stride + cc] = src[f];
                        rr += q32; cc += m32;
                        if (cc >= a.n) { cc -= a.n; ++rr; }
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < NP; ++i) e[k][i] = (active && i < a.n) ? stage[lane * stride + i] : (T)0;
                    __syncwarp();
                }
                if (BOUNDS) accepted[k] = active && (in_bounds<T, NP>(a, e[k], (T)1) || a.keep_last != 0);
            } else if (BOUNDS) {
                if (active) accepted[k] = draw_accepted<T, NP, true, ROUNDS>(a, a.first + local, e[k], s[k]);
                else {
#pragma unroll
                    for (int i = 0; i < NP; ++i) e[k][i] = (T)0;
                }
            } else {
                // no rejection loop: draw unconditionally (tail threads draw too, results unused)
                draw_accepted<T, NP, false, ROUNDS>(a, a.first + local, e[k], s[k]);
            }
        }

        T q[K], r[K];
        quad_and_dot_k<T, NP, K>(a, e, q, r);

#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint64_t local = local0 + (uint64_t)k * PF_BLOCK;
            const bool active = local < a.P;
            T ret, risk, sharpe;
            metrics_from<T>(q[k], r[k], s[k], a.rf, SRC == 1, ret, risk, sharpe);
            if (accepted[k]) {
                const uint32_t sub = (uint32_t)(tile * K + k);
                ++n_acc;
                if (sharpe > best_s) { best_s = sharpe; sub_s = sub; }
                const T d = -Math<T>::abs(risk - a.target);
                if (d > best_d) { best_d = d; sub_d = sub; }
                rmin = risk < rmin ? risk : rmin;
                rmax = risk > rmax ? risk : rmax;
            }
            // ---- optional write-back (app.py:719-722) ----
            if (active) {
                const T nanv = Math<T>::nan();
                if (a.ret_out != nullptr) a.ret_out[local] = accepted[k] ? ret : nanv;
                if (a.risk_out != nullptr) a.risk_out[local] = accepted[k] ? risk : nanv;
                if (a.sharpe_out != nullptr) a.sharpe_out[local] = accepted[k] ? sharpe : nanv;
                if (a.acc_out != nullptr) a.acc_out[local] = accepted[k] ? 1 : 0;
            }
            if (a.w_out != nullptr) {
                const T inv = SRC == 1 ? (T)1 : Math<T>::rcp(s[k]);
                const uint64_t warp_row0 = local - lane;
                const int rows = warp_row0 >= a.P ? 0 : (int)((a.P - warp_row0) < 32 ? (a.P - warp_row0) : 32);
                T* dst = a.w_out + warp_row0 * (uint64_t)a.n;
                bool done = false;
                if constexpr (sizeof(T) == 4) {
                    if (vec) {
#pragma unroll
                        for (int i = 0; i < NP; i += 4)
                            if (i < a.n)
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(stage) + lane * stride + i) =
                                    make_float4(e[k][i] * inv, e[k][i + 1] * inv, e[k][i + 2] * inv, e[k][i + 3] * inv);
                        __syncwarp();
                        if constexpr (NP == 4 || NP == 8 || NP == 16 || NP == 32) {
                            if (a.n == NP) warp_tile_store_vec_ct<NP>(reinterpret_cast<const float*>(stage), reinterpret_cast<float*>(dst), rows, lane);
                            else warp_tile_store_vec(reinterpret_cast<const float*>(stage), stride, reinterpret_cast<float*>(dst), rows, a.n, lane);
                        } else {
                            warp_tile_store_vec(reinterpret_cast<const float*>(stage), stride, reinterpret_cast<float*>(dst), rows, a.n, lane);
                        }
                        done = true;
                    }
                }
                if (!done) {
#pragma unroll
                    for (int i = 0; i < NP; ++i)
                        if (i < a.n) stage[lane * stride + i] = e[k][i] * inv;
                    __syncwarp();
                    const int total = rows * a.n;
                    int rr = lane / a.n, cc = lane % a.n;
                    for (int f = lane; f < total; f += 32) {
                        dst[f] = stage[rr * stride + cc];
                        rr += q32; cc += m32;
                        if (cc >= a.n) { cc -= a.n; ++rr; }
                    }
                }
                __syncwarp();
            }
        }
    }

    // ---- CTA reduction: (key, global index) argmax with first-occurrence tie-break ----
    uint64_t idx_s = sub_s == NONE ? MCP_NO_INDEX : a.first + (uint64_t)sub_s * PF_BLOCK + threadIdx.x;
    uint64_t idx_d = sub_d == NONE ? MCP_NO_INDEX : a.first + (uint64_t)sub_d * PF_BLOCK + threadIdx.x;
    warp_argmax<T>(best_s, idx_s);
    warp_argmax<T>(best_d, idx_d);
    rmin = warp_min<T>(rmin);
    rmax = warp_max<T>(rmax);
    n_acc = __reduce_add_sync(0xffffffffu, n_acc);

    __shared__ PfCand wc[PF_BLOCK / 32];
    __shared__ unsigned int wacc[PF_BLOCK / 32];
    if (lane == 0) {
        wc[warp] = PfCand{(double)best_s, idx_s, (double)best_d, idx_d, (double)rmin, (double)rmax};
        wacc[warp] = n_acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        PfCand b = wc[0];
        unsigned long long acc = wacc[0];
        for (int w = 1; w < PF_BLOCK / 32; ++w) {
            const PfCand o = wc[w];
            if (cand_better<double>(o.key_s, o.idx_s, b.key_s, b.idx_s)) { b.key_s = o.key_s; b.idx_s = o.idx_s; }
            if (cand_better<double>(o.key_d, o.idx_d, b.key_d, b.idx_d)) { b.key_d = o.key_d; b.idx_d = o.idx_d; }
            b.rmin = o.rmin < b.rmin ? o.rmin : b.rmin;
            b.rmax = o.rmax > b.rmax ? o.rmax : b.rmax;
            acc += wacc[w];
        }
        a.cands[blockIdx.x] = b;
        if (acc) atomicAdd(a.n_accepted, acc);
    }
}

// ---- packed FP32x2 variant of the Philox / no-bounds / no-weight-write sweep (the C3 kernel) ----
// Two portfolios share every instruction of the arithmetic: l2[i] = (lg2 U_i of portfolio A,
// of portfolio B).  It works on l = lg2(U) = -e so that no negation is ever materialised:
// q = l' S l (= e' S e), s = sum(-mask_i l_i), r = sum(-mu_i l_i).  Every value is bit-identical
// to the scalar path (IEEE fma per lane, same operation order), so small_replay reproduces it.
// OUT = false is the instance the 10^10-portfolio sweep runs: nothing is written per portfolio, so the
// per-portfolio pointer tests, staging addresses and stores are not even compiled in (the kernel is
// issue-bound: they were about 1/8 of its instruction stream).
template <int NP, int K, bool OUT, int ROUNDS>
__global__ void __launch_bounds__(PF_BLOCK) small_sweep_packed(const __grid_constant__ SmallArgs<float, NP> a) {
    static_assert(K % 2 == 0, "packed sweep pairs portfolios");
    constexpr int KP = K / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec = OUT && (a.n & 3) == 0 && (reinterpret_cast<uintptr_t>(a.w_out) & 15) == 0;
    const int stride = vec ? a.n + 4 : (a.n | 1);
    float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)warp * 32 * stride;
    const int q32 = OUT ? 32 / a.n : 0, m32 = OUT ? 32 % a.n : 0;
    const uint64_t n_sub = (a.P + PF_BLOCK - 1) / PF_BLOCK;
    const uint64_t n_tiles = (n_sub + K - 1) / K;
    constexpr uint32_t NONE = 0xffffffffu;
    float best_s = -Math<float>::inf(), best_d = -Math<float>::inf();
    uint32_t sub_s = NONE, sub_d = NONE;
    float rmin = Math<float>::inf(), rmax = -Math<float>::inf();
    uint32_t n_acc = 0;
    const uint32_t one_bits = opaque_u32(0x3f800000u);

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t local0 = tile * (uint64_t)(K * PF_BLOCK) + threadIdx.x;
        float2 l2[KP][NP];
        float2 s2[KP];
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
            const uint64_t ga = a.first + local0 + (uint64_t)(2 * kp) * PF_BLOCK, gb = ga + PF_BLOCK;
            s2[kp] = make_float2(0.f, 0.f);
            uint32_t fa[NP], fb[NP];
            philox_fields<NP, ROUNDS>((uint32_t)ga, (uint32_t)(ga >> 32), 0u, STREAM_WEIGHTS, a.rk, fa);
            philox_fields<NP, ROUNDS>((uint32_t)gb, (uint32_t)(gb >> 32), 0u, STREAM_WEIGHTS, a.rk, fb);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const float2 f = make_float2(__uint_as_float(mant_or(fa[i], one_bits)), __uint_as_float(mant_or(fb[i], one_bits)));
                const float2 u = fma2(f, bcast2(-1.0f), bcast2(2.0f));            // U = 2 - f in (0, 1]
                l2[kp][i] = make_float2(Math<float>::lg2(u.x), Math<float>::lg2(u.y));
                s2[kp] = fma2(l2[kp][i], bcast2(a.nmask[i]), s2[kp]);
            }
        }
        float2 q2[KP], r2[KP];
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) { q2[kp] = make_float2(0.f, 0.f); r2[kp] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            float2 t[KP];
#pragma unroll
            for (int kp = 0; kp < KP; ++kp) t[kp] = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const float2 sij = bcast2(a.sig[i * (i + 1) / 2 + j]);
#pragma unroll
                for (int kp = 0; kp < KP; ++kp) t[kp] = fma2(sij, l2[kp][j], t[kp]);
            }
            const float2 nmu = bcast2(a.nmu[i]);
#pragma unroll
            for (int kp = 0; kp < KP; ++kp) {
                q2[kp] = fma2(l2[kp][i], t[kp], q2[kp]);
                r2[kp] = fma2(nmu, l2[kp][i], r2[kp]);
            }
        }
        if constexpr (!OUT) {
            // selection only; whole tiles (all but the last of the range) skip the per-portfolio range test
            auto pick = [&](int k, bool active) {
                const float q = (k & 1) ? q2[k / 2].y : q2[k / 2].x;
                const float r = (k & 1) ? r2[k / 2].y : r2[k / 2].x;
                const float s = (k & 1) ? s2[k / 2].y : s2[k / 2].x;
                float ret, risk, sharpe;
                metrics_from<float>(q, r, s, a.rf, false, ret, risk, sharpe);
                if (active) {
                    const uint32_t sub = (uint32_t)(tile * K + k);
                    ++n_acc;
                    if (sharpe > best_s) { best_s = sharpe; sub_s = sub; }
                    const float d = -fabsf(risk - a.target);
                    if (d > best_d) { best_d = d; sub_d = sub; }
                    rmin = fminf(rmin, risk);
                    rmax = fmaxf(rmax, risk);
                }
            };
            if ((tile + 1) * (uint64_t)(K * PF_BLOCK) <= a.P) {
#pragma unroll
                for (int k = 0; k < K; ++k) pick(k, true);
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) pick(k, local0 + (uint64_t)k * PF_BLOCK < a.P);
            }
            continue;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint64_t local = local0 + (uint64_t)k * PF_BLOCK;
            const bool active = local < a.P;
            const float q = (k & 1) ? q2[k / 2].y : q2[k / 2].x;
            const float r = (k & 1) ? r2[k / 2].y : r2[k / 2].x;
            const float s = (k & 1) ? s2[k / 2].y : s2[k / 2].x;
            float ret, risk, sharpe;
            metrics_from<float>(q, r, s, a.rf, false, ret, risk, sharpe);
            if (active) {
                const uint32_t sub = (uint32_t)(tile * K + k);
                ++n_acc;
                if (sharpe > best_s) { best_s = sharpe; sub_s = sub; }
                const float d = -fabsf(risk - a.target);
                if (d > best_d) { best_d = d; sub_d = sub; }
                rmin = fminf(rmin, risk);
                rmax = fmaxf(rmax, risk);
                if constexpr (OUT) {
                    if (a.ret_out != nullptr) a.ret_out[local] = ret;
                    if (a.risk_out != nullptr) a.risk_out[local] = risk;
                    if (a.sharpe_out != nullptr) a.sharpe_out[local] = sharpe;
                    if (a.acc_out != nullptr) a.acc_out[local] = 1;
                }
            }
            if (OUT && a.w_out != nullptr) {
                // w_i = e_i / s = l_i * (-1 / s): stage this thread's row, then the warp streams the tile out
                const float ninv = -Math<float>::rcp(s);
                const uint64_t warp_row0 = local - lane;
                const int rows = warp_row0 >= a.P ? 0 : (int)((a.P - warp_row0) < 32 ? (a.P - warp_row0) : 32);
                float* dst = a.w_out + warp_row0 * (uint64_t)a.n;
                if (vec) {
#pragma unroll
                    for (int i = 0; i < NP; i += 4) {
                        if (i < a.n) {
                            float4 w4;
                            w4.x = ((k & 1) ? l2[k / 2][i + 0].y : l2[k / 2][i + 0].x) * ninv;
                            w4.y = ((k & 1) ? l2[k / 2][i + 1].y : l2[k / 2][i + 1].x) * ninv;
                            w4.z = ((k & 1) ? l2[k / 2][i + 2].y : l2[k / 2][i + 2].x) * ninv;
                            w4.w = ((k & 1) ? l2[k / 2][i + 3].y : l2[k / 2][i + 3].x) * ninv;
                            *reinterpret_cast<float4*>(stage + lane * stride + i) = w4;
                        }
                    }
                    __syncwarp();
                    if constexpr (NP == 4 || NP == 8 || NP == 16 || NP == 32) {
                        if (a.n == NP) warp_tile_store_vec_ct<NP>(stage, dst, rows, lane);
                        else warp_tile_store_vec(stage, stride, dst, rows, a.n, lane);
                    } else {
                        warp_tile_store_vec(stage, stride, dst, rows, a.n, lane);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NP; ++i)
                        if (i < a.n) stage[lane * stride + i] = ((k & 1) ? l2[k / 2][i].y : l2[k / 2][i].x) * ninv;
                    __syncwarp();
                    const int total = rows * a.n;
                    int rr = lane / a.n, cc = lane % a.n;
                    for (int f = lane; f < total; f += 32) {
                        dst[f] = stage[rr * stride + cc];
                        rr += q32; cc += m32;
                        if (cc >= a.n) { cc -= a.n; ++rr; }
                    }
                }
                __syncwarp();
            }
        }
    }

    uint64_t idx_s = sub_s == NONE ? MCP_NO_INDEX : a.first + (uint64_t)sub_s * PF_BLOCK + threadIdx.x;
    uint64_t idx_d = sub_d == NONE ? MCP_NO_INDEX : a.first + (uint64_t)sub_d * PF_BLOCK + threadIdx.x;
    warp_argmax<float>(best_s, idx_s);
    warp_argmax<float>(best_d, idx_d);
    rmin = warp_min<float>(rmin);
    rmax = warp_max<float>(rmax);
    n_acc = __reduce_add_sync(0xffffffffu, n_acc);
    __shared__ PfCand wc[PF_BLOCK / 32];
    __shared__ unsigned int wacc[PF_BLOCK / 32];
    if (lane == 0) {
        wc[warp] = PfCand{(double)best_s, idx_s, (double)best_d, idx_d, (double)rmin, (double)rmax};
        wacc[warp] = n_acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        PfCand b = wc[0];
        unsigned long long acc = wacc[0];
        for (int w = 1; w < PF_BLOCK / 32; ++w) {
            const PfCand o = wc[w];
            if (cand_better<double>(o.key_s, o.idx_s, b.key_s, b.idx_s)) { b.key_s = o.key_s; b.idx_s = o.idx_s; }
            if (cand_better<double>(o.key_d, o.idx_d, b.key_d, b.idx_d)) { b.key_d = o.key_d; b.idx_d = o.idx_d; }
            b.rmin = o.rmin < b.rmin ? o.rmin : b.rmin;
            b.rmax = o.rmax > b.rmax ? o.rmax : b.rmax;
            acc += wacc[w];
        }
        a.cands[blockIdx.x] = b;
        if (acc) atomicAdd(a.n_accepted, acc);
    }
}

// Re-evaluates the selected portfolios with exactly the sweep's arithmetic and emits
// (index, key, ret, risk, sharpe, weights[N]) records in FP64.
template <typename T, int NP, int ROUNDS>
__global__ void small_replay(const __grid_constant__ SmallArgs<T, NP> a, int n_sel, uint64_t idx0, uint64_t idx1,
                             const T* rows, int bounds, double* rec) {
    const int k = threadIdx.x;
    if (k >= n_sel) return;
    const uint64_t gidx = k == 0 ? idx0 : idx1;
    double* out = rec + (size_t)k * (PF_REC_HEADER + a.n);
    out[0] = __longlong_as_double((long long)gidx);
    if (gidx == MCP_NO_INDEX) {
        for (int i = 1; i < PF_REC_HEADER + a.n; ++i) out[i] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    T e[NP];
    T s = (T)1;
    const bool supplied = rows != nullptr;
    if (supplied) {
#pragma unroll
        for (int i = 0; i < NP; ++i) e[i] = i < a.n ? rows[(size_t)k * a.n + i] : (T)0;
    } else if (bounds) {
        draw_accepted<T, NP, true, ROUNDS>(a, gidx, e, s);
    } else {
        draw_accepted<T, NP, false, ROUNDS>(a, gidx, e, s);
    }
    T q, r, ret, risk, sharpe;
    quad_and_dot<T, NP>(a, e, q, r);
    metrics_from<T>(q, r, s, a.rf, supplied, ret, risk, sharpe);
    out[1] = k == 0 ? (double)sharpe : (double)Math<T>::abs(risk - a.target);
    out[2] = (double)ret;
    out[3] = (double)risk;
    out[4] = (double)sharpe;
    const T inv = supplied ? (T)1 : Math<T>::rcp(s);
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (i < a.n) out[PF_REC_HEADER + i] = (double)(e[i] * inv);
}

template <typename T, int NP>
static void fill_small_args(const PfJob& job, SmallArgs<T, NP>& a) {
    const int n = job.n;
    for (int i = 0; i < NP; ++i) {
        for (int j = 0; j <= i; ++j) {
            double v = 0.0;
            if (i < n && j < n) v = (i == j) ? job.sigma[(size_t)i * n + i] : job.sigma[(size_t)i * n + j] + job.sigma[(size_t)j * n + i];
            a.sig[i * (i + 1) / 2 + j] = (T)v;
        }
        a.mu[i] = i < n ? (T)job.mu[i] : (T)0;
        a.mask[i] = i < n ? (T)1 : (T)0;
        a.nmu[i] = -a.mu[i];
        a.nmask[i] = -a.mask[i];
        a.lo[i] = (i < n && job.lo) ? (T)job.lo[i] : (T)-1e30;
        a.hi[i] = (i < n && job.hi) ? (T)job.hi[i] : (T)1e30;
    }
    a.rf = (T)job.rf;
    a.target = (T)job.target;
    a.n = n;
    a.max_tries = job.max_tries;
    a.keep_last = job.keep_last;
    a.k0 = (uint32_t)job.seed;
    a.k1 = (uint32_t)(job.seed >> 32);
    philox_keys_fill(a.rk, job.seed);
    a.first = job.first;
    a.P = job.P;
    a.w_in = (const T*)job.w_in;
    a.w_out = (T*)job.w_out;
    a.ret_out = (T*)job.ret_out;
    a.risk_out = (T*)job.risk_out;
    a.sharpe_out = (T*)job.sharpe_out;
    a.acc_out = job.acc_out;
    a.cands = job.cands;
    a.n_accepted = job.n_accepted;
}

template <typename T, int NP, int K>
static int small_launch_k(mcp_context* h, PfJob& job, const SmallArgs<T, NP>& a) {
    void (*kern)(SmallArgs<T, NP>) = nullptr;
    const bool r7 = job.rounds == 7;                 // supplied weights draw nothing: one instance
    if (job.w_in) kern = job.bounds ? small_sweep<T, NP, K, 1, true, 10> : small_sweep<T, NP, K, 1, false, 10>;
    else if (job.bounds) kern = r7 ? small_sweep<T, NP, K, 0, true, 7> : small_sweep<T, NP, K, 0, true, 10>;
    else kern = r7 ? small_sweep<T, NP, K, 0, false, 7> : small_sweep<T, NP, K, 0, false, 10>;
    const bool staging = job.w_in != nullptr || job.w_out != nullptr;
    const size_t smem = staging ? (size_t)(PF_BLOCK / 32) * 32 * (job.n + 4) * sizeof(T) : 0;
    if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MCP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PF_BLOCK, smem));
    if (per_sm < 1) return mcp_fail(h, MCP_ERR_CUDA, "small_sweep<N=%d>: zero occupancy (smem %zu B)", NP, smem);
    const uint64_t n_tiles = ((job.P + PF_BLOCK - 1) / PF_BLOCK + K - 1) / K;
    uint64_t grid = (uint64_t)h->prop.multiProcessorCount * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    if (grid > (uint64_t)job.max_blocks) grid = job.max_blocks;
    if (grid < 1) grid = 1;
    job.blocks_used = (int)grid;
    kern<<<(unsigned)grid, PF_BLOCK, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <int NP, int K>
static int small_launch_packed(mcp_context* h, PfJob& job, const SmallArgs<float, NP>& a) {
    const bool any_out = job.w_out || job.ret_out || job.risk_out || job.sharpe_out || job.acc_out;
    auto kern = any_out ? small_sweep_packed<NP, K, true, 10> : small_sweep_packed<NP, K, false, 10>;
    if (job.rounds == 7) kern = any_out ? small_sweep_packed<NP, K, true, 7> : small_sweep_packed<NP, K, false, 7>;
    const size_t smem = job.w_out ? (size_t)(PF_BLOCK / 32) * 32 * (job.n + 4) * sizeof(float) : 0;
    if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MCP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PF_BLOCK, smem));
    if (per_sm < 1) return mcp_fail(h, MCP_ERR_CUDA, "small_sweep_packed<N=%d>: zero occupancy", NP);
    const uint64_t n_tiles = ((job.P + PF_BLOCK - 1) / PF_BLOCK + K - 1) / K;
    uint64_t grid = (uint64_t)h->prop.multiProcessorCount * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    if (grid > (uint64_t)job.max_blocks) grid = job.max_blocks;
    if (grid < 1) grid = 1;
    job.blocks_used = (int)grid;
    kern<<<(unsigned)grid, PF_BLOCK, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <typename T, int NP>
int pf_small_launch_t(mcp_context* h, PfJob& job) {
    SmallArgs<T, NP> a;
    fill_small_args<T, NP>(job, a);
    // K portfolios per thread only when nothing streams through HBM: with supplied weights or
    // weight write-back the kernel is memory-side bound and the extra registers only cost occupancy
    constexpr int K = SweepK<T, NP>::value;
    if constexpr (sizeof(T) == 4 && K % 2 == 0) {
        if (job.w_in == nullptr && !job.bounds) {
            if constexpr (K == 4) {
                static const int k_override = getenv("MCP_SWEEP_K") ? atoi(getenv("MCP_SWEEP_K")) : 0;   // tuning knob
                if (k_override == 2) return small_launch_packed<NP, 2>(h, job, a);
            }
            return small_launch_packed<NP, K>(h, job, a);
        }
    }
    if (K > 1 && job.w_in == nullptr && job.w_out == nullptr) return small_launch_k<T, NP, K>(h, job, a);
    return small_launch_k<T, NP, 1>(h, job, a);
}

template <typename T, int NP>
int pf_small_replay_t(mcp_context* h, const PfJob& job, const PfReplay& rp) {
    SmallArgs<T, NP> a;
    fill_small_args<T, NP>(job, a);
    auto kern = job.rounds == 7 ? small_replay<T, NP, 7> : small_replay<T, NP, 10>;
    kern<<<1, 32, 0, job.stream>>>(a, rp.n_sel, rp.idx[0], rp.idx[1], (const T*)rp.rows, job.bounds ? 1 : 0, rp.rec);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

}  // namespace mcp
