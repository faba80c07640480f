// Portfolio sweep for N > 32 assets (C5: N = 256, 1e9 portfolios).
//
// FP32, 32 < N <= 256 -- `large_sweep`: the per-portfolio state no longer fits a thread, so a CTA
// works on a tile of 64 portfolios:
//
//   1. GENERATE  the tile's un-normalised exponentials e[p][i] (Philox4x32-10, same counter
//                layout and 24-bit fields as the small kernel: global index / attempt / block) straight
//                into shared memory, asset-major (Wt[i][p], row stride 66 -> conflict-free
//                writes by the 4 threads that share a portfolio and conflict-free LDS.64 reads);
//   2. QUADRATIC FORM as a register-tiled SIMT "SGEMM with row-dot":  Y' = W S' restricted to the
//                lower triangle (S' = Sigma with doubled off-diagonals, so q = sum_j w_j Y'_j).
//                S' lives in shared memory k-major and padded-triangular (139 KB at N = 256).
//                A warp owns one pair of 16-column groups (g, G-1-g): its K-extent is uniform
//                across the warp (no divergence) and all warps do the same 16(G+1) k-steps.
//                Per k-step a lane does one LDS.64 (its 2 portfolios) + 8 broadcast LDS.128
//                (S' row slices) for 32 FFMA2 (packed FP32x2: the float2 of the lane's two
//                portfolios times a broadcast S' element) = 64 FMAs on 2 x 32 accumulators;
//   3. EPILOGUE  multiplies the accumulators with W again (row-dot), and reduces q, sum(e) and
//                e.mu over the 8 warps through shared memory in a fixed order (deterministic,
//                independent of the portfolio's slot in the tile -> replays are bit-identical);
//   4. FINALISE  ret / risk / Sharpe per portfolio, bounds check (4 threads per portfolio rescan
//                its row; rejected rows are redrawn with attempt+1, app.py:700-707), selection
//                tracking, optional coalesced write-back.
//
// Executed FMAs per portfolio: 16*16*G(G+1)/2 = 34 816 at N = 256 vs 32 896 algorithmic (+6 %).
//
// FP64 (any N > 32) and FP32 with N > 256 -- `generic_sweep`: one warp per portfolio, Sigma read
// from global memory (L2-resident); a parity / coverage path, not a throughput path.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "mcp_device.cuh"
#include "mcp_portfolio.h"

namespace mcp {

constexpr int LG_TP = 64;            // portfolios per tile
constexpr int LG_THREADS = 256;
constexpr int LG_WARPS = 8;
constexpr int LG_CG = 16;            // columns per group
constexpr int LG_WSTRIDE = 66;       // row stride of Wt (floats): even and == 2 (mod 8)
constexpr int LG_MAX_N = 256;
constexpr int GEN_MAX_N = 1024;

struct LargeArgs {
    const float* st;                 // global, padded-triangular k-major S'
    const float* mu;                 // global [np]
    const float* lo;                 // global [np]
    const float* hi;
    const float* w_in;               // supplied weights [P, n] or null
    float* w_out;
    float* ret_out;
    float* risk_out;
    float* sharpe_out;
    uint8_t* acc_out;
    PfCand* cands;
    unsigned long long* n_accepted;
    uint64_t first, P;
    // list mode (the tensor-core sweep's deferred rows, mcp_portfolio_large_tc.cu): row p of this launch is the portfolio with
    // global index list[p], *list_count rows in all (a device-side count: no host round trip between the two kernels)
    const uint64_t* list;
    const unsigned long long* list_count;
    const uint16_t* list_att;        // optional: attempts row p has already used up (its first draw here is attempt list_att[p])
    int n, np, st_floats;
    int max_tries, keep_last, bounds;
    uint32_t k0, k1;
    float rf, target;
};

// offset of row k in the padded-triangular layout: sum_{k' < k} (np - 16 floor(k'/16))
__host__ __device__ inline int lg_row_offset(int k, int np) {
    const int b = k / LG_CG, r = k % LG_CG;
    return b * LG_CG * np - LG_CG * LG_CG * (b * (b - 1) / 2) + r * (np - LG_CG * b);
}

// slot states
constexpr int ST_INACTIVE = 0, ST_PENDING = 1, ST_ACCEPTED = 2, ST_SKIPPED = 3;

template <int ROUNDS>
__global__ void __launch_bounds__(LG_THREADS, 1) large_sweep(const LargeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sSt = reinterpret_cast<float*>(smem_raw);                 // [st_floats]
    float* sW = sSt + a.st_floats;                                   // [np][LG_WSTRIDE]
    float* sMu = sW + (size_t)a.np * LG_WSTRIDE;                     // [np]
    float* sRedQ = sMu + a.np;                                       // [LG_WARPS][LG_TP]
    float* sRedS = sRedQ + LG_WARPS * LG_TP;
    float* sRedR = sRedS + LG_WARPS * LG_TP;
    float* sSum = sRedR + LG_WARPS * LG_TP;                          // [LG_TP] sum(e)
    float* sInv = sSum + LG_TP;                                      // [LG_TP] 1 / sum(e)
    int* sState = reinterpret_cast<int*>(sInv + LG_TP);              // [LG_TP]
    int* sFlag = sState + LG_TP;                                     // [1] "some row must be redrawn" (4 ints reserved)
    int* sStart = sFlag + 4;                                         // [LG_TP] attempt number of the row's first draw (list mode)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < a.st_floats; i += LG_THREADS) sSt[i] = a.st[i];
    for (int i = tid; i < a.np; i += LG_THREADS) sMu[i] = a.mu[i];
    __syncthreads();

    const int G = a.np / LG_CG;                 // column groups (even)
    const int gA = warp, gB = G - 1 - warp;     // this warp's pair (valid if warp < G/2)
    const bool has_pair = warp < G / 2;
    const int gp = tid >> 2, gc = tid & 3;      // generation role: portfolio slot, block phase
    const bool supplied = a.w_in != nullptr;

    const bool listed = a.list != nullptr;
    const uint64_t P = listed ? (uint64_t)*a.list_count : a.P;
    const uint64_t n_tiles = (P + LG_TP - 1) / LG_TP;
    float best_s = -Math<float>::inf(), best_d = -Math<float>::inf();
    uint64_t idx_s = MCP_NO_INDEX, idx_d = MCP_NO_INDEX;
    float rmin = Math<float>::inf(), rmax = -Math<float>::inf();
    unsigned int n_acc = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t p0 = tile * LG_TP;
        const int rows = (P - p0) < (uint64_t)LG_TP ? (int)(P - p0) : LG_TP;
        if (tid < LG_TP) {
            sState[tid] = tid < rows ? ST_PENDING : ST_INACTIVE;
            sStart[tid] = (listed && a.list_att != nullptr && tid < rows) ? (int)a.list_att[p0 + tid] : 0;
        }
        __syncthreads();
        const int tries = (a.bounds && !supplied) ? a.max_tries : 1;
        for (int attempt = 0; attempt < tries; ++attempt) {
            // ---- 1. generate / load the rows that are pending ----
            if (supplied) {
                const float* src = a.w_in + p0 * (uint64_t)a.n;
                for (int f = tid; f < LG_TP * a.np; f += LG_THREADS) {
                    const int p = f / a.np, i = f - p * a.np;
                    sW[i * LG_WSTRIDE + p] = (p < rows && i < a.n) ? src[(size_t)p * a.n + i] : 0.f;
                }
            } else if (sState[gp] == ST_PENDING || (attempt == 0 && sState[gp] == ST_INACTIVE)) {
                const uint64_t gidx = listed ? (gp < rows ? a.list[p0 + gp] : 0ull) : a.first + p0 + gp;
                const uint32_t c0 = (uint32_t)gidx, c1 = (uint32_t)(gidx >> 32);
                const bool live = sState[gp] == ST_PENDING;
                for (int G = gc; G < a.np / 16; G += 4) {          // 16-asset group = 16 fields = Philox blocks 3G .. 3G+2
                    uint32_t f[16];
                    philox_fields<16, ROUNDS>(c0, c1, (uint32_t)(attempt + sStart[gp]), STREAM_WEIGHTS | (uint32_t)(3 * G), a.k0, a.k1, f);
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int i = 16 * G + k;
                        const float e = -Math<float>::lg2(Math<float>::unit_open0(f[k]));
                        sW[i * LG_WSTRIDE + gp] = (live && i < a.n) ? e : 0.f;
                    }
                }
            }
            if (tid == 0) sFlag[0] = 0;
            __syncthreads();

            // ---- 2. Y' = W S' over the lower triangle: 2 portfolios x (16 + 16) columns per lane ----
            // accumulators are float2 = this lane's two portfolios: every FFMA2 does both (packed FP32x2)
            float2 qp = make_float2(0.f, 0.f), sp = make_float2(0.f, 0.f), rp = make_float2(0.f, 0.f);
            if (has_pair) {
                float2 accA[LG_CG], accB[LG_CG];
#pragma unroll
                for (int c = 0; c < LG_CG; ++c) { accA[c] = make_float2(0.f, 0.f); accB[c] = make_float2(0.f, 0.f); }
                const float* wrow = sW + 2 * lane;
                for (int kblk = 0; kblk <= gB; ++kblk) {          // 16 rows sharing one row length
                    const int rowlen = a.np - LG_CG * kblk;
                    const float* srow = sSt + lg_row_offset(LG_CG * kblk, a.np);
                    const int offA = LG_CG * (gA - kblk), offB = LG_CG * (gB - kblk);
                    const float* wk = wrow + (size_t)(LG_CG * kblk) * LG_WSTRIDE;
                    if (kblk <= gA) {
#pragma unroll 4
                        for (int r = 0; r < LG_CG; ++r) {
                            const float2 w2 = *reinterpret_cast<const float2*>(wk + (size_t)r * LG_WSTRIDE);
                            const float4* sb = reinterpret_cast<const float4*>(srow + (size_t)r * rowlen + offB);
                            const float4* sa = reinterpret_cast<const float4*>(srow + (size_t)r * rowlen + offA);
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const float4 s4 = sb[v];
                                accB[4 * v + 0] = fma2(w2, bcast2(s4.x), accB[4 * v + 0]);
                                accB[4 * v + 1] = fma2(w2, bcast2(s4.y), accB[4 * v + 1]);
                                accB[4 * v + 2] = fma2(w2, bcast2(s4.z), accB[4 * v + 2]);
                                accB[4 * v + 3] = fma2(w2, bcast2(s4.w), accB[4 * v + 3]);
                            }
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const float4 s4 = sa[v];
                                accA[4 * v + 0] = fma2(w2, bcast2(s4.x), accA[4 * v + 0]);
                                accA[4 * v + 1] = fma2(w2, bcast2(s4.y), accA[4 * v + 1]);
                                accA[4 * v + 2] = fma2(w2, bcast2(s4.z), accA[4 * v + 2]);
                                accA[4 * v + 3] = fma2(w2, bcast2(s4.w), accA[4 * v + 3]);
                            }
                        }
                    } else {
#pragma unroll 4
                        for (int r = 0; r < LG_CG; ++r) {
                            const float2 w2 = *reinterpret_cast<const float2*>(wk + (size_t)r * LG_WSTRIDE);
                            const float4* sb = reinterpret_cast<const float4*>(srow + (size_t)r * rowlen + offB);
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const float4 s4 = sb[v];
                                accB[4 * v + 0] = fma2(w2, bcast2(s4.x), accB[4 * v + 0]);
                                accB[4 * v + 1] = fma2(w2, bcast2(s4.y), accB[4 * v + 1]);
                                accB[4 * v + 2] = fma2(w2, bcast2(s4.z), accB[4 * v + 2]);
                                accB[4 * v + 3] = fma2(w2, bcast2(s4.w), accB[4 * v + 3]);
                            }
                        }
                    }
                }
                // ---- 3. row-dot with W, plus sum(e) and e.mu over this warp's 32 columns ----
#pragma unroll
                for (int c = 0; c < LG_CG; ++c) {
                    const int ja = LG_CG * gA + c, jb = LG_CG * gB + c;
                    const float2 wa = *reinterpret_cast<const float2*>(wrow + (size_t)ja * LG_WSTRIDE);
                    const float2 wb = *reinterpret_cast<const float2*>(wrow + (size_t)jb * LG_WSTRIDE);
                    qp = fma2(accA[c], wa, qp);
                    qp = fma2(accB[c], wb, qp);
                    sp = fma2(wa, bcast2(1.0f), sp);
                    sp = fma2(wb, bcast2(1.0f), sp);
                    rp = fma2(bcast2(sMu[ja]), wa, rp);
                    rp = fma2(bcast2(sMu[jb]), wb, rp);
                }
            }
            *reinterpret_cast<float2*>(sRedQ + warp * LG_TP + 2 * lane) = qp;
            *reinterpret_cast<float2*>(sRedS + warp * LG_TP + 2 * lane) = sp;
            *reinterpret_cast<float2*>(sRedR + warp * LG_TP + 2 * lane) = rp;
            __syncthreads();

            // ---- 4a. sum(e) per portfolio (fixed order) ----
            if (tid < LG_TP) {
                float s = 0.f;
                for (int w = 0; w < LG_WARPS; ++w) s += sRedS[w * LG_TP + tid];
                sSum[tid] = s;
                sInv[tid] = supplied ? 1.f : Math<float>::rcp(s);
            }
            __syncthreads();
            // ---- 4b. bounds (app.py:703-705): the 4 threads of a portfolio rescan its row ----
            if (a.bounds) {
                int ok = 1;
                if (sState[gp] == ST_PENDING) {
                    const float inv = sInv[gp];
                    for (int i = gc; i < a.n; i += 4) {
                        const float w = sW[i * LG_WSTRIDE + gp] * inv;
                        ok &= (w >= a.lo[i] && w <= a.hi[i]) ? 1 : 0;
                    }
                }
                ok &= __shfl_xor_sync(0xffffffffu, ok, 1);
                ok &= __shfl_xor_sync(0xffffffffu, ok, 2);
                if (gc == 0 && sState[gp] == ST_PENDING) {
                    if (ok) sState[gp] = ST_ACCEPTED;
                    else if (attempt + sStart[gp] + 1 < tries) sFlag[0] = 1;      // benign race: all writers store 1
                    else sState[gp] = a.keep_last ? ST_ACCEPTED : ST_SKIPPED;
                }
            } else if (tid < LG_TP && sState[tid] == ST_PENDING) {
                sState[tid] = ST_ACCEPTED;
            }
            __syncthreads();
            if (sFlag[0] == 0) break;
            __syncthreads();                    // everyone has read the flag before thread 0 clears it
        }

        // ---- 4c. metrics, selection, write-back ----
        if (tid < LG_TP) {
            const bool active = tid < rows;
            const uint64_t local = listed ? (active ? a.list[p0 + tid] - a.first : 0ull) : p0 + tid;    // row in the output arrays
            const bool accepted = sState[tid] == ST_ACCEPTED;
            float q = 0.f, r = 0.f;
            for (int w = 0; w < LG_WARPS; ++w) { q += sRedQ[w * LG_TP + tid]; r += sRedR[w * LG_TP + tid]; }
            float ret, risk, sharpe;
            metrics_from<float>(q, r, sSum[tid], a.rf, supplied, ret, risk, sharpe);
            if (accepted) {
                ++n_acc;
                const uint64_t g = a.first + local;
                // tiles ascend: the first occurrence is kept (a list is in arrival order, so ties compare indices there)
                if (sharpe > best_s || (listed && sharpe == best_s && g < idx_s)) { best_s = sharpe; idx_s = g; }
                const float d = -fabsf(risk - a.target);
                if (d > best_d || (listed && d == best_d && g < idx_d)) { best_d = d; idx_d = g; }
                rmin = fminf(rmin, risk);
                rmax = fmaxf(rmax, risk);
            }
            if (active) {
                const float nanv = Math<float>::nan();
                if (a.ret_out) a.ret_out[local] = accepted ? ret : nanv;
                if (a.risk_out) a.risk_out[local] = accepted ? risk : nanv;
                if (a.sharpe_out) a.sharpe_out[local] = accepted ? sharpe : nanv;
                if (a.acc_out) a.acc_out[local] = accepted ? 1 : 0;
            }
        }
        if (a.w_out != nullptr) {
            float* dst = a.w_out + p0 * (uint64_t)a.n;
            const int total = rows * a.n;
            for (int f = tid; f < total; f += LG_THREADS) {
                const int p = f / a.n, i = f - p * a.n;
                if (listed) a.w_out[(a.list[p0 + p] - a.first) * (uint64_t)a.n + i] = sW[i * LG_WSTRIDE + p] * sInv[p];
                else dst[f] = sW[i * LG_WSTRIDE + p] * sInv[p];
            }
        }
        __syncthreads();
    }

    // ---- CTA reduction of the selection candidates ----
    warp_argmax<float>(best_s, idx_s);
    warp_argmax<float>(best_d, idx_d);
    rmin = warp_min<float>(rmin);
    rmax = warp_max<float>(rmax);
    n_acc = __reduce_add_sync(0xffffffffu, n_acc);
    __shared__ PfCand wc[LG_WARPS];
    __shared__ unsigned int wacc[LG_WARPS];
    if (lane == 0) {
        wc[warp] = PfCand{(double)best_s, idx_s, (double)best_d, idx_d, (double)rmin, (double)rmax};
        wacc[warp] = n_acc;
    }
    __syncthreads();
    if (tid == 0) {
        PfCand b = wc[0];
        unsigned long long acc = wacc[0];
        for (int w = 1; w < LG_WARPS; ++w) {
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

// ---------------------------------------------------------------------------------------------
// generic warp-per-portfolio sweep (FP64 any N > 32; FP32 N > 256): Sigma (full, row-major) in
// global memory, the portfolio's exponentials in shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct GenArgs {
    const T* sigma;      // [n, n]
    const T* mu;
    const T* lo;
    const T* hi;
    const T* w_in;
    T* w_out;
    T* ret_out;
    T* risk_out;
    T* sharpe_out;
    uint8_t* acc_out;
    PfCand* cands;
    unsigned long long* n_accepted;
    uint64_t first, P;
    int n, np;
    int max_tries, keep_last, bounds;
    uint32_t k0, k1;
    T rf, target;
};

constexpr int GEN_THREADS = 256;
constexpr int GEN_WARPS = GEN_THREADS / 32;

template <typename T> __device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor<T>(v, m);
    return v;
}

template <typename T, int ROUNDS>
__global__ void __launch_bounds__(GEN_THREADS) generic_sweep(const GenArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* e = reinterpret_cast<T*>(smem_raw) + (size_t)warp * a.np;
    const bool supplied = a.w_in != nullptr;

    T best_s = -Math<T>::inf(), best_d = -Math<T>::inf();
    uint64_t idx_s = MCP_NO_INDEX, idx_d = MCP_NO_INDEX;
    T rmin = Math<T>::inf(), rmax = -Math<T>::inf();
    unsigned int n_acc = 0;
    const uint64_t warps_total = (uint64_t)gridDim.x * GEN_WARPS;
    for (uint64_t p = (uint64_t)blockIdx.x * GEN_WARPS + warp; p < a.P; p += warps_total) {
        const uint64_t gidx = a.first + p;
        const uint32_t c0 = (uint32_t)gidx, c1 = (uint32_t)(gidx >> 32);
        bool accepted = true;
        T s = (T)1;
        const int tries = (a.bounds && !supplied) ? a.max_tries : 1;
        for (int attempt = 0; attempt < tries; ++attempt) {
            __syncwarp();
            if (supplied) {
                for (int i = lane; i < a.np; i += 32) e[i] = i < a.n ? a.w_in[p * (uint64_t)a.n + i] : (T)0;
                s = (T)1;
            } else {
                T part = (T)0;
                for (int b = lane; b < a.np / 4; b += 32) {          // b = group of 4 assets
                    uint32_t x[4];
                    if constexpr (sizeof(T) == 4) {
                        // 24-bit fields 4b .. 4b+3 = words 3b .. 3b+2 of the stream (one or two Philox blocks)
                        const int w0 = 3 * b, b0 = w0 >> 2, b1 = (w0 + 2) >> 2;
                        uint32_t y0[4], y1[4];
                        philox4x32_r<ROUNDS>(c0, c1, (uint32_t)attempt, STREAM_WEIGHTS | (uint32_t)b0, a.k0, a.k1, y0);
                        if (b1 != b0) philox4x32_r<ROUNDS>(c0, c1, (uint32_t)attempt, STREAM_WEIGHTS | (uint32_t)b1, a.k0, a.k1, y1);
                        uint32_t w3[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const int wi = w0 + j, k4 = wi & 3;
                            const uint32_t* src = (wi >> 2) == b0 ? y0 : y1;
                            w3[j] = k4 == 0 ? src[0] : k4 == 1 ? src[1] : k4 == 2 ? src[2] : src[3];
                        }
                        fields_from_triple(w3[0], w3[1], w3[2], x);
                    } else {
                        philox4x32_r<ROUNDS>(c0, c1, (uint32_t)attempt, STREAM_WEIGHTS | (uint32_t)b, a.k0, a.k1, x);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = 4 * b + k;
                        const T v = i < a.n ? -Math<T>::lg2(Math<T>::unit_open0(x[k])) : (T)0;
                        e[i] = v;
                        part += v;
                    }
                }
                s = warp_sum_t<T>(part);
            }
            __syncwarp();
            if (!a.bounds) break;
            const T inv = supplied ? (T)1 : Math<T>::rcp(s);
            int ok = 1;
            for (int i = lane; i < a.n; i += 32) {
                const T w = e[i] * inv;
                ok &= (w >= a.lo[i] && w <= a.hi[i]) ? 1 : 0;
            }
            ok = __all_sync(0xffffffffu, ok);
            accepted = ok || (attempt + 1 == tries && a.keep_last);
            if (ok) break;
        }
        // q = e' Sigma e, r = e . mu  (Sigma symmetric: column access is coalesced)
        T q = (T)0, r = (T)0;
        for (int i0 = 0; i0 < a.n; i0 += 32) {
            const int i = i0 + lane;
            T t = (T)0;
            if (i < a.n) {
                for (int j = 0; j < a.n; ++j) t = Math<T>::fma(a.sigma[(size_t)j * a.n + i], e[j], t);
                q = Math<T>::fma(e[i], t, q);
                r = Math<T>::fma(a.mu[i], e[i], r);
            }
        }
        q = warp_sum_t<T>(q);
        r = warp_sum_t<T>(r);
        T ret, risk, sharpe;
        metrics_from<T>(q, r, s, a.rf, supplied, ret, risk, sharpe);
        if (accepted) {
            ++n_acc;
            if (sharpe > best_s) { best_s = sharpe; idx_s = gidx; }
            const T d = -Math<T>::abs(risk - a.target);
            if (d > best_d) { best_d = d; idx_d = gidx; }
            rmin = risk < rmin ? risk : rmin;
            rmax = risk > rmax ? risk : rmax;
        }
        if (lane == 0) {
            const T nanv = Math<T>::nan();
            if (a.ret_out) a.ret_out[p] = accepted ? ret : nanv;
            if (a.risk_out) a.risk_out[p] = accepted ? risk : nanv;
            if (a.sharpe_out) a.sharpe_out[p] = accepted ? sharpe : nanv;
            if (a.acc_out) a.acc_out[p] = accepted ? 1 : 0;
        }
        if (a.w_out) {
            const T inv = supplied ? (T)1 : Math<T>::rcp(s);
            for (int i = lane; i < a.n; i += 32) a.w_out[p * (uint64_t)a.n + i] = e[i] * inv;
        }
    }
    // all lanes of a warp hold identical candidates; n_acc is per warp (count lane 0 only)
    __shared__ PfCand wc[GEN_WARPS];
    __shared__ unsigned int wacc[GEN_WARPS];
    if (lane == 0) {
        wc[warp] = PfCand{(double)best_s, idx_s, (double)best_d, idx_d, (double)rmin, (double)rmax};
        wacc[warp] = n_acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        PfCand b = wc[0];
        unsigned long long acc = wacc[0];
        for (int w = 1; w < GEN_WARPS; ++w) {
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

// record packing for the replay of a selected portfolio (see small_replay)
template <typename T>
__global__ void pack_record(uint64_t gidx, int which, T target, const T* w, const T* ret, const T* risk, const T* sharpe,
                            int n, double* rec) {
    if (threadIdx.x == 0) {
        rec[0] = __longlong_as_double((long long)gidx);
        rec[1] = which == 0 ? (double)sharpe[0] : (double)Math<T>::abs(risk[0] - target);
        rec[2] = (double)ret[0];
        rec[3] = (double)risk[0];
        rec[4] = (double)sharpe[0];
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) rec[PF_REC_HEADER + i] = (double)w[i];
}

// ---- host side -------------------------------------------------------------------------------

static bool use_tiled(const PfJob& job) { return job.dtype == MCP_F32 && job.n <= LG_MAX_N; }

static int large_launch_tiled(mcp_context* h, PfJob& job) {
    const int n = job.n, np = (n + 31) / 32 * 32, G = np / LG_CG;
    const int st_floats = lg_row_offset(np, np);
    // S' k-major padded-triangular: row k holds columns j = 16*floor(k/16) .. np-1, S'[j][k] for j >= k else 0
    std::vector<float> host((size_t)st_floats + 3 * np, 0.f);
    for (int k = 0; k < n; ++k) {
        const int base = lg_row_offset(k, np), j0 = LG_CG * (k / LG_CG);
        for (int j = std::max(k, j0); j < n; ++j) {
            const double v = j == k ? job.sigma[(size_t)k * n + k] : job.sigma[(size_t)j * n + k] + job.sigma[(size_t)k * n + j];
            host[(size_t)base + (j - j0)] = (float)v;
        }
    }
    float* hmu = host.data() + st_floats;
    float* hlo = hmu + np;
    float* hhi = hlo + np;
    for (int i = 0; i < np; ++i) {
        hmu[i] = i < n ? (float)job.mu[i] : 0.f;
        hlo[i] = (i < n && job.lo) ? (float)job.lo[i] : -1e30f;
        hhi[i] = (i < n && job.hi) ? (float)job.hi[i] : 1e30f;
    }
    float* dev = nullptr;
    // next to the tcgen05 sweep (bounded route) this kernel's constants live in their own slot: slot 6 holds that kernel's table
    const bool beside_tc = job.tc_bounds_route != 0 || job.idx_list != nullptr;
    MCP_CHECK(mcp_dev_reserve(h, beside_tc ? 20 : 6, host.size() * sizeof(float), (void**)&dev));
    // once per mcp_portfolios call: later chunks (other pipeline streams included) and the replays reuse the table
    if (beside_tc ? !job.lgl_uploaded : (job.lg_table_epoch == 0 || job.lg_table_epoch != h->const_epoch || job.lg_table_kind != 0)) {
        MCP_CUDA(h, cudaDeviceSynchronize());                  // another pipeline stream may still read the previous table
        MCP_CUDA(h, cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, job.stream));
        MCP_CUDA(h, cudaStreamSynchronize(job.stream));        // `host` is pageable and dies at scope exit
        if (beside_tc) job.lgl_uploaded = true;
        else {
            job.lg_table_epoch = ++h->const_epoch;
            job.lg_table_kind = 0;
        }
    }

    LargeArgs a;
    a.st = dev; a.mu = dev + st_floats; a.lo = a.mu + np; a.hi = a.lo + np;
    a.w_in = (const float*)job.w_in; a.w_out = (float*)job.w_out;
    a.ret_out = (float*)job.ret_out; a.risk_out = (float*)job.risk_out; a.sharpe_out = (float*)job.sharpe_out;
    a.acc_out = job.acc_out; a.cands = job.cands; a.n_accepted = job.n_accepted;
    a.first = job.first; a.P = job.P; a.n = n; a.np = np; a.st_floats = st_floats;
    a.list = job.idx_list; a.list_count = job.idx_count; a.list_att = job.idx_attempt;
    a.max_tries = job.max_tries; a.keep_last = job.keep_last; a.bounds = job.bounds ? 1 : 0;
    a.k0 = (uint32_t)job.seed; a.k1 = (uint32_t)(job.seed >> 32);
    a.rf = (float)job.rf; a.target = (float)job.target;
    (void)G;
    const size_t smem = ((size_t)st_floats + (size_t)np * LG_WSTRIDE + np + 3 * LG_WARPS * LG_TP + 2 * LG_TP) * sizeof(float) +
                        (2 * LG_TP + 4) * sizeof(int);
    if (smem > h->prop.sharedMemPerBlockOptin)
        return mcp_fail(h, MCP_ERR_INVALID, "large_sweep: N=%d needs %zu B of shared memory (max %zu)", n, smem,
                        (size_t)h->prop.sharedMemPerBlockOptin);
    auto kern = job.rounds == 7 ? large_sweep<7> : large_sweep<10>;
    MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t n_tiles = (job.P + LG_TP - 1) / LG_TP;
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount, n_tiles);
    grid = std::max<uint64_t>(1, std::min<uint64_t>(grid, job.max_blocks));
    job.blocks_used = (int)grid;
    kern<<<(unsigned)grid, LG_THREADS, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <typename T>
static int large_launch_generic(mcp_context* h, PfJob& job) {
    const int n = job.n, np = (n + 3) / 4 * 4;
    if (n > GEN_MAX_N) return mcp_fail(h, MCP_ERR_INVALID, "n_assets=%d exceeds the supported maximum of %d", n, GEN_MAX_N);
    std::vector<T> host((size_t)n * n + 3 * np);
    for (size_t i = 0; i < (size_t)n * n; ++i) host[i] = (T)job.sigma[i];
    T* hmu = host.data() + (size_t)n * n;
    T* hlo = hmu + np;
    T* hhi = hlo + np;
    for (int i = 0; i < np; ++i) {
        hmu[i] = i < n ? (T)job.mu[i] : (T)0;
        hlo[i] = (i < n && job.lo) ? (T)job.lo[i] : (T)-1e30;
        hhi[i] = (i < n && job.hi) ? (T)job.hi[i] : (T)1e30;
    }
    T* dev = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 6, host.size() * sizeof(T), (void**)&dev));
    const int kind = sizeof(T) == 8 ? 2 : 1;
    if (job.lg_table_epoch == 0 || job.lg_table_epoch != h->const_epoch || job.lg_table_kind != kind) {
        MCP_CUDA(h, cudaDeviceSynchronize());
        MCP_CUDA(h, cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, job.stream));
        MCP_CUDA(h, cudaStreamSynchronize(job.stream));
        job.lg_table_epoch = ++h->const_epoch;
        job.lg_table_kind = kind;
    }
    GenArgs<T> a;
    a.sigma = dev; a.mu = dev + (size_t)n * n; a.lo = a.mu + np; a.hi = a.lo + np;
    a.w_in = (const T*)job.w_in; a.w_out = (T*)job.w_out;
    a.ret_out = (T*)job.ret_out; a.risk_out = (T*)job.risk_out; a.sharpe_out = (T*)job.sharpe_out;
    a.acc_out = job.acc_out; a.cands = job.cands; a.n_accepted = job.n_accepted;
    a.first = job.first; a.P = job.P; a.n = n; a.np = np;
    a.max_tries = job.max_tries; a.keep_last = job.keep_last; a.bounds = job.bounds ? 1 : 0;
    a.k0 = (uint32_t)job.seed; a.k1 = (uint32_t)(job.seed >> 32);
    a.rf = (T)job.rf; a.target = (T)job.target;
    const size_t smem = (size_t)GEN_WARPS * np * sizeof(T);
    auto kern = job.rounds == 7 ? generic_sweep<T, 7> : generic_sweep<T, 10>;
    if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MCP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GEN_THREADS, smem));
    if (per_sm < 1) return mcp_fail(h, MCP_ERR_CUDA, "generic_sweep: zero occupancy (smem %zu B)", smem);
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount * per_sm, (job.P + GEN_WARPS - 1) / GEN_WARPS);
    grid = std::max<uint64_t>(1, std::min<uint64_t>(grid, job.max_blocks));
    job.blocks_used = (int)grid;
    kern<<<(unsigned)grid, GEN_THREADS, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

int pf_large_launch_list(mcp_context* h, PfJob& job) {
    if (!use_tiled(job)) return mcp_fail(h, MCP_ERR_INVALID, "list mode needs the tiled FP32 sweep (n_assets <= %d)", LG_MAX_N);
    return large_launch_tiled(h, job);
}

int pf_large_launch(mcp_context* h, PfJob& job) {
    if (pf_large_tc_eligible(job)) return job.bounds ? pf_large_launch_tc_bounded(h, job) : pf_large_launch_tc(h, job);
    if (use_tiled(job)) return large_launch_tiled(h, job);
    return job.dtype == MCP_F64 ? large_launch_generic<double>(h, job) : large_launch_generic<float>(h, job);
}

// Replay = the same kernel on the one-portfolio range [idx, idx+1) with write-back into scratch,
// so the record is bit-identical to what the sweep computed for that portfolio.
int pf_large_replay(mcp_context* h, const PfJob& job, const PfReplay& rp) {
    const size_t es = job.dtype == MCP_F64 ? 8 : 4;
    const int n = job.n;
    unsigned char* scratch = nullptr;
    const size_t per = ((size_t)(n + 3) * es + 255) / 256 * 256 + 256;
    MCP_CHECK(mcp_dev_reserve(h, 7, 2 * per + 4 * sizeof(PfCand) + 64, (void**)&scratch));
    for (int k = 0; k < rp.n_sel; ++k) {
        double* rec = rp.rec + (size_t)k * (PF_REC_HEADER + n);
        unsigned char* base = scratch + (size_t)k * per;
        PfJob j = job;
        j.first = rp.idx[k];
        j.P = 1;
        j.w_in = rp.rows ? (const unsigned char*)rp.rows + (size_t)k * n * es : nullptr;
        j.w_out = base;
        j.ret_out = base + (size_t)n * es;
        j.risk_out = base + (size_t)(n + 1) * es;
        j.sharpe_out = base + (size_t)(n + 2) * es;
        j.acc_out = nullptr;
        j.cands = (PfCand*)(scratch + 2 * per);
        j.n_accepted = (unsigned long long*)(scratch + 2 * per + 4 * sizeof(PfCand));
        j.max_blocks = 1;
        if (j.bounds) j.tc_bounds_route = 2;      // one row: the tiled SIMT kernel replays the rejection loop (its decisions are the sweep's)
        if (rp.idx[k] == MCP_NO_INDEX) {
            std::vector<double> nanrec(PF_REC_HEADER + n, NAN);
            uint64_t none = MCP_NO_INDEX;
            memcpy(&nanrec[0], &none, 8);
            MCP_CUDA(h, cudaMemcpyAsync(rec, nanrec.data(), nanrec.size() * 8, cudaMemcpyHostToDevice, job.stream));
            MCP_CUDA(h, cudaStreamSynchronize(job.stream));
            continue;
        }
        MCP_CHECK(pf_large_launch(h, j));
        if (job.dtype == MCP_F64)
            pack_record<double><<<1, 128, 0, job.stream>>>(rp.idx[k], k, (double)job.target, (const double*)j.w_out, (const double*)j.ret_out,
                                                          (const double*)j.risk_out, (const double*)j.sharpe_out, n, rec);
        else
            pack_record<float><<<1, 128, 0, job.stream>>>(rp.idx[k], k, (float)job.target, (const float*)j.w_out, (const float*)j.ret_out,
                                                         (const float*)j.risk_out, (const float*)j.sharpe_out, n, rec);
        MCP_CUDA(h, cudaGetLastError());
        h->launches++;
    }
    return MCP_OK;
}

}  // namespace mcp
