// Tensor-core portfolio sweep for 32 < N <= 256 assets, FP32, no bounds (C5: N = 256, 1e9 portfolios; Philox or supplied weights).
//
// The quadratic form of 128 portfolios at a time is a GEMM: Y' = E S' with E [128 x N] the tile's
// un-normalised exponentials and S' the lower triangle of Sigma with doubled off-diagonals, then
// q_p = sum_j Y'_pj E_pj (app.py:709 restated on the triangle).  E never exists in shared or global
// memory: it is produced by Philox in the registers of the thread that owns the row, written to
// TENSOR MEMORY as the MMA's A operand (tcgen05.st), multiplied on the 5th-generation tensor cores
// (tcgen05.mma, A from TMEM, B = S' from shared memory, FP32 accumulators in TMEM); an epilogue thread of
// the same TMEM lane reads the accumulator row AND the staged E back (tcgen05.ld) for the row-dot.
//
// FP32-class accuracy from narrow tensor-core operands (split operands, all accumulation in FP32).  Two splits share the kernel:
//
//   FP16 split (template F16 = true; Philox rows, where |l| = |lg2 U| <= 24 is inside FP16's range):
//     E  = h1 + h2             h1 = FP16(E), h2 = FP16(E - h1): 22 significant bits
//     S' = (S1 + S2) 2^-e      S1 = FP16(S' 2^e), S2 = FP16(S' 2^e - S1), 2^e puts the largest entry in [2^13, 2^14)
//     Y' = h1 S1 + h2 S1 + h1 S2      three FP16 MMA sets at the full 16-bit rate; q is scaled back by 2^-e
//     stage = h1 (16 columns) + h2 (16) + E itself in FP32 (32, for the epilogue's row-dot) = 64 columns -> 4 stages, 4 generator groups
//     S1 + S2 = 2 x 72 KB of shared memory
//   TF32 split (F16 = false; supplied weights, which may be any FP32 number):
//     E  = Ehi + Elo           Ehi = E with the low 13 mantissa bits cleared (a TF32 number), Elo = E - Ehi exact
//     S' = Shi + Slo           Shi = TF32 round-to-nearest of S', Slo = BF16(S' - Shi)
//     Y' = Ehi Shi (tf32) + Elo Shi (tf32) + bf16(E) Slo (bf16)       relative error per product ~ 2^-19
//     stage = 80 columns -> 3 stages / groups; Shi (144 KB) + Slo in BF16 (72 KB)
// In both, S' is stored as a triangle in 32-row K chunks: chunk c holds rows k in [32c, 32c+32) and only columns j < 32(c+1).
//
// Roles (one CTA per SM, persistent over tiles of 128 portfolios; thread = portfolio row = TMEM lane; G = 4 (FP16) or 3 (TF32) groups):
//   warps 0..4G-1   generator groups of 128 threads, one A stage each.  Group g produces the CTA's K chunks n = g (mod G):
//                   6 Philox calls (32 uniforms as 24-bit fields) -> l = lg2(U) = -e -> split -> tcgen05.st into its stage -> a_full[g].
//                   It never waits for the tensor core, only for its stage to have been read back (a_free[g]).
//   warp 4G+4       TMEM allocation and the MMA issue loop: warp-convergent, one elected thread; FP16: the chunk's descriptor words
//                   come from a table in shared memory and the six tcgen05.mma + tcgen05.commit are ONE asm block under one elect.
//                   Chunks run from the widest (c = C-1, all columns, overwrites the accumulator) to the narrowest,
//                   so column block c is final as soon as chunk c's MMAs complete and the epilogue of a tile
//                   overlaps its remaining MMAs.
//   warps 4G..4G+3  epilogue + finaliser: per chunk, tcgen05.ld the 32 finished accumulator columns and the chunk's
//                   stage (FP16: E itself; TF32: l = hi + lo exactly), release the stage, accumulate q, sum l, l.mu; per tile compute
//                   return / risk / Sharpe (app.py:708-711), track the selections, write the arrays.
// Measured on B200 (N = 256, FP16 split): 4.3e9 portfolios/s in a 50 ms launch, 3.8e9 in the 10^9-portfolio envelope step (TF32 split:
// 3.6 / 3.1e9; SIMT kernel 5.9e8).  The kernel is bound by the SIMT issue of the generator warps (ncu: issue 59 %, ALU pipe 49 %,
// tensor pipe 39 %): what moved it were instruction counts -- one-LOP3 field masks, Philox round keys precomputed on the host (TcArgs::rk),
// the fused MMA issue (DESIGN.md section 4 has the history and the ablations).
// The Philox counter layout and the 24-bit uniform fields are those of every other FP32 sweep kernel (global index /
// attempt 0 / block; mcp_device.cuh), so the weights are the same portfolios the SIMT kernels and oracle/philox_np.py produce.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "mcp_device.cuh"
#include "mcp_portfolio.h"
#include "mcp_tcgen05.cuh"

namespace mcp {

constexpr int TC_ROWS = 128;                 // portfolios per tile = TMEM lanes
constexpr int TC_KC = 32;                    // K (assets) per chunk
constexpr int TC_MAX_N = 256;
constexpr uint32_t TC_COL_A = 256;           // first TMEM column of the A stages (accumulator = columns 0..255)
constexpr int TC_MAX_GROUPS = 4;
constexpr uint64_t TC_BOUNDED_MAX_P = 1ull << 31;     // rows per launch of the bounded route (the index lists are 8 B per row of a 2^26-row sub-range)
// Two operand splits share the kernel (template parameter F16):
//   TF32 split (any FP32 input, so supplied weights use it): stage = hi 32 + lo 32 + bf16 16 = 80 columns, 3 stages / generator groups
//   FP16 split (Philox rows: |lg2 U| <= 24 fits FP16):        stage = h1 16 + h2 16 + l 32 = 64 columns, 4 stages / generator groups
template <bool F16> struct TcCfg {
    static constexpr int GROUPS = F16 ? 4 : 3;                      // row groups = A stages
    static constexpr uint32_t STAGE_COLS = F16 ? 64 : 80;
    static constexpr int EPI_WARP0 = 4 * GROUPS;                    // four epilogue warps (one per TMEM lane quadrant)
    static constexpr int MMA_WARP = 4 * GROUPS + 4;
    static constexpr int THREADS = (4 * GROUPS + 5) * 32;
};

struct TcArgs {
    const unsigned char* table;              // global: Shi image, Slo image (canonical UMMA layout), mu[np]
    const float* w_in;                       // supplied weights [P, n] (parity mode) or null (Philox)
    float* w_out;                            // raw exponentials (scaled by inv_out afterwards) or null
    float* inv_out;                          // 1 / sum(e) per portfolio (only with w_out)
    float* ret_out;
    float* risk_out;
    float* sharpe_out;
    uint8_t* acc_out;
    PfCand* cands;
    unsigned long long* n_accepted;
    uint64_t first, P;
    int n, np;
    uint32_t k0, k1;
    PhiloxKeys rk;
    float rf, target;
    float qscale;                            // FP16 split: S' is stored times a power of two, q comes back times that; this undoes it
    // BOUNDS instances (Philox rows, app.py:700-707): the table carries three more float[np] arrays behind mu -- 1 / max_weight
    // (0 where the bound cannot bite), 1 / min_weight and an additive term that takes unconstrained / padded assets out of the
    // minimum.  One launch evaluates ONE attempt of its rows (the contiguous range, or in_list[0 .. *in_count) in the later
    // rounds): rows inside the bounds with a margin are accepted; rows outside with a margin go to retry_list (next round,
    // attempt + 1) or are skipped / kept when this was the last attempt; rows too close to a bound to call in this kernel's
    // summation order go to simt_list with the attempt they are at, and the tiled SIMT kernel decides them.
    const uint64_t* in_list;
    const unsigned long long* in_count;
    uint64_t* retry_list;
    unsigned long long* retry_count;
    uint64_t* simt_list;
    uint16_t* simt_att;
    unsigned long long* simt_count;
    uint32_t attempt;
    int max_tries, keep_last;
    int has_lo;                              // some min_weight > 0
};

// warp-aggregated append of `v` to list[(*count)++] for the lanes with `take` (called by all 32 lanes)
__device__ __forceinline__ unsigned long long warp_append(bool take, unsigned long long* count, int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m == 0u) return 0ull;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned)__popc(m & ((1u << lane) - 1u));
}

// min / max of three (sm_100: one FMNMX3)
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

__host__ __device__ inline uint32_t tc_hi_off(int c) { return 2048u * (uint32_t)(c * (c + 1)); }     // bytes before chunk c
__host__ __device__ inline uint32_t tc_lo_off(int c) { return 1024u * (uint32_t)(c * (c + 1)); }

template <bool F16, int ROUNDS, bool BOUNDS>
__global__ void __launch_bounds__(TcCfg<F16>::THREADS, 1) large_sweep_tc(const TcArgs a) {
    static_assert(!BOUNDS || F16, "bounds rejection runs on the Philox (FP16-split) instance");
    constexpr int TC_GROUPS = TcCfg<F16>::GROUPS, TC_EPI_WARP0 = TcCfg<F16>::EPI_WARP0, TC_MMA_WARP = TcCfg<F16>::MMA_WARP;
    constexpr uint32_t TC_STAGE_COLS = TcCfg<F16>::STAGE_COLS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int C = a.np / TC_KC;                                         // K chunks per tile (2..8)
    // TF32 split: Shi image (32-bit) | Slo image (16-bit);  FP16 split: S1 image | S2 image (both 16-bit)
    const uint32_t hi_bytes = F16 ? tc_lo_off(C) : tc_hi_off(C), lo_bytes = tc_lo_off(C);
    unsigned char* sHi = smem;
    unsigned char* sLo = smem + hi_bytes;
    float* sMu = reinterpret_cast<float*>(sLo + lo_bytes);                                       // [np]
    const float* sInvHi = sMu + a.np;                                                            // [np] each, BOUNDS only
    const float* sInvLo = sInvHi + a.np;
    const float* sBiasLo = sInvLo + a.np;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sMu + (BOUNDS ? 4 : 1) * a.np);
    uint64_t* a_full = bars;                           // [G]  128 arrivals: the stage's A operand is in TMEM
    uint64_t* d_done = bars + TC_MAX_GROUPS;           // [G]  tcgen05.commit: the stage's MMAs are complete
    uint64_t* a_free = bars + 2 * TC_MAX_GROUPS;       // [G]  128 arrivals: the epilogue has read the stage back
    uint64_t* drained = bars + 3 * TC_MAX_GROUPS;      // [1]  128 arrivals: the tile's last accumulator columns were read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_MAX_GROUPS + 1);
    PfCand* sCand = reinterpret_cast<PfCand*>(bars + 3 * TC_MAX_GROUPS + 3);               // [4] per epilogue warp
    unsigned int* sAcc = reinterpret_cast<unsigned int*>(sCand + 4);    // [4]
    uint4* sDesc = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(sAcc + 4) + 15) & ~(uintptr_t)15);   // [8] per-chunk descriptor words (FP16 split)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // S' images + mu arrive by TMA bulk copies (cp.async.bulk: global -> shared through the async proxy, completion counted
    // in bytes on an mbarrier) issued by one thread while the other warps set up barriers and tensor memory; the table in
    // global memory is laid out exactly like this shared-memory block.
    uint64_t* table_bar = bars + 3 * TC_MAX_GROUPS + 2;
    if (tid == 0) {
        for (int g = 0; g < TC_GROUPS; ++g) { mbar_init(&a_full[g], TC_ROWS); mbar_init(&d_done[g], 1); mbar_init(&a_free[g], TC_ROWS); }
        mbar_init(drained, TC_ROWS);
        mbar_init(table_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = hi_bytes + lo_bytes + (uint32_t)a.np * (BOUNDS ? 16u : 4u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(table_bar)), "r"(bytes) : "memory");
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t part = bytes - off < 32768u ? bytes - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + off)), "l"(a.table + off), "r"(part), "r"(smem_u32(table_bar)) : "memory");
        }
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(table_bar, 0u);                                          // every consumer of S' / mu waits for the bytes to land
    const uint32_t tmem = *tmem_slot;

    const bool listed = BOUNDS && a.in_list != nullptr;
    const uint64_t P = listed ? (uint64_t)*a.in_count : a.P;
    const uint64_t n_tiles = (P + TC_ROWS - 1) / TC_ROWS;

    if (warp < TC_EPI_WARP0) {
        // ================= generators: Philox -> lg2 -> split -> A stage =================
        // The group works on l = lg2(U) = -e (the quadratic form is even in the sign; the epilogue flips the two
        // linear sums).  It never waits for the tensor core, only for its stage to be read back (a_free).
        const int g = warp >> 2, row = tid & (TC_ROWS - 1);
        const uint32_t stage = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + TC_COL_A + TC_STAGE_COLS * (uint32_t)g;
        const uint32_t one_bits = opaque_u32(0x3f800000u);   // (x & 0x7fffff) | 1.0f stays ONE LOP3 (mcp_device.cuh)
        uint32_t k = 0;                                  // chunks this group has produced (barrier parity)
        uint32_t first_mod = 0;                          // (tl * C) mod G
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint64_t p0 = tile * TC_ROWS;
            const bool live = p0 + (uint64_t)row < P;
            const uint64_t gidx = listed ? (live ? a.in_list[p0 + (uint64_t)row] : 0ull) : a.first + p0 + (uint64_t)row;
            const uint64_t out_row = BOUNDS ? gidx - a.first : p0 + (uint64_t)row;      // row in the output arrays
            const uint32_t c0 = (uint32_t)gidx, c1 = (uint32_t)(gidx >> 32);
            for (int ci = (int)((g + (uint32_t)TC_GROUPS - first_mod) % (uint32_t)TC_GROUPS); ci < C; ci += TC_GROUPS, ++k) {
                const int c = C - 1 - ci;
                float l[TC_KC];
                const int i0 = TC_KC * c;
                if (a.w_in != nullptr) {
                    // ---- supplied weights (parity mode): this row's 32 values of chunk c from global memory ----
                    const float* src = a.w_in + (p0 + (uint64_t)row) * (uint64_t)a.n + (uint64_t)i0;
                    if ((a.n & 3) == 0) {
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            const float4 v = (live && i0 + 4 * m < a.n) ? __ldg(reinterpret_cast<const float4*>(src) + m) : make_float4(0.f, 0.f, 0.f, 0.f);
                            l[4 * m] = v.x; l[4 * m + 1] = v.y; l[4 * m + 2] = v.z; l[4 * m + 3] = v.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < TC_KC; ++j) l[j] = (live && i0 + j < a.n) ? __ldg(src + j) : 0.f;
                    }
                } else {
                    // ---- Philox: 24-bit fields 32c .. 32c+31 = blocks 6c .. 6c+5; l = lg2(U) = -e ----
                    uint32_t f[TC_KC];
                    philox_fields<TC_KC, ROUNDS>(c0, c1, BOUNDS ? a.attempt : 0u, STREAM_WEIGHTS | (uint32_t)(6 * c), a.rk, f);
#pragma unroll
                    for (int j = 0; j < TC_KC; j += 2) {          // U = 2 - f in (0, 1], two per FFMA2 (same values as unit_open0)
                        const float2 m2 = make_float2(__uint_as_float(mant_or(f[j], one_bits)), __uint_as_float(mant_or(f[j + 1], one_bits)));
                        const float2 u2 = fma2(m2, bcast2(-1.0f), bcast2(2.0f));
                        l[j] = Math<float>::lg2(u2.x);
                        l[j + 1] = Math<float>::lg2(u2.y);
                    }
                    if (i0 + TC_KC > a.n) {              // only the last chunk can reach past n (uniform branch)
#pragma unroll
                        for (int j = 0; j < TC_KC; ++j)
                            if (i0 + j >= a.n) l[j] = 0.f;
                    }
                }
                if (a.w_out != nullptr && live) {         // raw values (e = -l, or the supplied weight); tc_scale_rows normalises them
                    const float sg = a.w_in != nullptr ? 1.f : -1.f;
                    float* dst = a.w_out + out_row * (uint64_t)a.n + (uint64_t)i0;
                    if ((a.n & 3) == 0) {
#pragma unroll
                        for (int m = 0; m < 8; ++m)
                            if (i0 + 4 * m < a.n) reinterpret_cast<float4*>(dst)[m] = make_float4(sg * l[4 * m], sg * l[4 * m + 1], sg * l[4 * m + 2], sg * l[4 * m + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < TC_KC; ++j)
                            if (i0 + j < a.n) dst[j] = sg * l[j];
                    }
                }
                if constexpr (F16) {
                    // ---- split: h1 = FP16(l), h2 = FP16(l - h1): 22 significant bits in two FP16 operands; l itself rides along in
                    //      the stage for the epilogue's row-dot.  The split is done BEFORE waiting for the stage: the wait -> tcgen05.st
                    //      -> a_full segment is on the ring every stage travels (generator -> MMA -> epilogue -> generator), and
                    //      only the four stores have to be on it ----
                    uint32_t h1[16], h2[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 l2 = make_float2(l[2 * j], l[2 * j + 1]);
                        const __half2 p1 = __floats2half2_rn(l2.x, l2.y);                                  // low half = even k
                        const float2 b = __half22float2(p1);
                        const float2 rem = fma2(b, bcast2(-1.0f), l2);                                     // exact
                        const __half2 p2 = __floats2half2_rn(rem.x, rem.y);
                        h1[j] = *reinterpret_cast<const uint32_t*>(&p1);
                        h2[j] = *reinterpret_cast<const uint32_t*>(&p2);
                    }
                    if (k > 0) {                         // the stage's previous contents must have been read back
                        mbar_wait(&a_free[g], (k - 1u) & 1u);
                        tc_fence_after();
                    }
                    tmem_st16(stage, h1);
                    tmem_st16(stage + 16u, h2);
                    tmem_st16(stage + 32u, *reinterpret_cast<const uint32_t(*)[16]>(&l[0]));
                    tmem_st16(stage + 48u, *reinterpret_cast<const uint32_t(*)[16]>(&l[16]));
                } else {
                    if (k > 0) {                         // the stage's previous contents must have been read back
                        mbar_wait(&a_free[g], (k - 1u) & 1u);
                        tc_fence_after();
                    }
                    // ---- split: hi = TF32 truncation, lo = exact remainder, bf = BF16 copy for the Slo term ----
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t hi[16], lo[16], bf[8];
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {            // lo = l - hi exactly (one FFMA2 per pair)
                            hi[j] = __float_as_uint(l[16 * h + j]) & 0xffffe000u;
                            hi[j + 1] = __float_as_uint(l[16 * h + j + 1]) & 0xffffe000u;
                            const float2 lo2 = fma2(make_float2(__uint_as_float(hi[j]), __uint_as_float(hi[j + 1])), bcast2(-1.0f),
                                                    make_float2(l[16 * h + j], l[16 * h + j + 1]));
                            lo[j] = __float_as_uint(lo2.x);
                            lo[j + 1] = __float_as_uint(lo2.y);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const __nv_bfloat162 p2 = __floats2bfloat162_rn(l[16 * h + 2 * j], l[16 * h + 2 * j + 1]);        // low half = even k
                            bf[j] = *reinterpret_cast<const uint32_t*>(&p2);
                        }
                        tmem_st16(stage + 16u * h, hi);
                        tmem_st16(stage + 32u + 16u * h, lo);
                        tmem_st8(stage + 64u + 8u * h, bf);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive(&a_full[g]);
            }
            first_mod = (first_mod + (uint32_t)C) % (uint32_t)TC_GROUPS;
        }
    } else if (warp == TC_MMA_WARP) {
        // ================= MMA issue: one elected thread, warp-convergent =================
        // The ten tcgen05.mma of a chunk go out back to back once its A stage is full: the descriptors are plain
        // functions of the chunk index, computed on warp-uniform values before the barrier waits, and the issue is
        // predicated on elect.sync in convergent code (an `if (lane == 0)` makes ptxas wrap every UTCHMMA in a lane loop).
        const uint32_t sbo = 128u, hi_base = smem_u32(sHi), lo_base = smem_u32(sLo);
        const uint32_t id32_base = tc_idesc(2u, 0u), id16_base = tc_idesc(1u, 0u), idf16_base = tc_idesc(0u, 0u);   // TF32 / BF16 / FP16 operands
        if constexpr (F16) {
            if (lane < C) {                              // low descriptor words of chunk `lane`: start address >> 4 | (LBO >> 4) << 16
                const uint32_t Nl = (uint32_t)(TC_KC * (lane + 1)), lbol = (Nl / 8u) * 128u;
                const uint32_t b1 = hi_base + tc_lo_off(lane), b2 = lo_base + tc_lo_off(lane);
                sDesc[lane] = make_uint4((uint32_t)tc_sdesc(b1, lbol, sbo), (uint32_t)tc_sdesc(b1 + 2u * lbol, lbol, sbo),
                                         (uint32_t)tc_sdesc(b2, lbol, sbo), (uint32_t)tc_sdesc(b2 + 2u * lbol, lbol, sbo));
            }
            __syncwarp();
        }
        uint32_t g = 0, cyc = 0;                         // chunk n = 3 * cyc + g uses stage g for the cyc-th time
        uint32_t tl = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            for (int ci = 0; ci < C; ++ci) {
                const int c = C - 1 - ci;
                const uint32_t N = (uint32_t)(TC_KC * (c + 1)), lbo = (N / 8u) * 128u;
                const uint32_t a_st = tmem + TC_COL_A + TC_STAGE_COLS * g;
                if constexpr (F16) {
                    // six FP16 MMAs (K = 16 each): h1 S1 + h2 S1 + h1 S2 for the two halves of the chunk, one elect, one asm block;
                    // the chunk's descriptor words come from the table this warp built in shared memory before the loop (this warp
                    // shares its scheduler with five busy ones: every instruction between a_full and the first MMA is ring latency)
                    const uint4 dw = sDesc[c];
                    const uint32_t idh = idf16_base | ((N >> 3) << 17);
                    mbar_wait(&a_full[g], cyc & 1u);
                    if (ci == 0 && tl > 0) mbar_wait(drained, (tl - 1u) & 1u);          // chunk C-1 overwrites the whole accumulator
                    tc_fence_after();
                    mma6_f16_commit_elect(tmem, a_st, a_st + 16u, dw.x, dw.y, dw.z, dw.w, 0x4008u, idh, ci > 0 ? 1u : 0u, smem_u32(&d_done[g]));
                } else {
                    const uint32_t bhi = hi_base + tc_hi_off(c), blo = lo_base + tc_lo_off(c);
                    const uint64_t bd0 = tc_sdesc(bhi, lbo, sbo), bd1 = tc_sdesc(bhi + 2u * lbo, lbo, sbo), bd2 = tc_sdesc(bhi + 4u * lbo, lbo, sbo),
                                   bd3 = tc_sdesc(bhi + 6u * lbo, lbo, sbo), bl0 = tc_sdesc(blo, lbo, sbo), bl1 = tc_sdesc(blo + 2u * lbo, lbo, sbo);
                    const uint32_t id32 = id32_base | ((N >> 3) << 17), id16 = id16_base | ((N >> 3) << 17);
                    const uint32_t a_hi = a_st, a_lo = a_hi + 32u, a_bf = a_hi + 64u;
                    mbar_wait(&a_full[g], cyc & 1u);
                    if (ci == 0 && tl > 0) mbar_wait(drained, (tl - 1u) & 1u);          // chunk C-1 overwrites the whole accumulator
                    tc_fence_after();
                    mma_tf32_ts_elect(tmem, a_hi, bd0, id32, ci > 0 ? 1u : 0u);
                    mma_tf32_ts_elect(tmem, a_lo, bd0, id32, 1u);
                    mma_tf32_ts_elect(tmem, a_hi + 8u, bd1, id32, 1u);
                    mma_tf32_ts_elect(tmem, a_lo + 8u, bd1, id32, 1u);
                    mma_tf32_ts_elect(tmem, a_hi + 16u, bd2, id32, 1u);
                    mma_tf32_ts_elect(tmem, a_lo + 16u, bd2, id32, 1u);
                    mma_tf32_ts_elect(tmem, a_hi + 24u, bd3, id32, 1u);
                    mma_tf32_ts_elect(tmem, a_lo + 24u, bd3, id32, 1u);
                    mma_bf16_ts_elect(tmem, a_bf, bl0, id16, 1u);
                    mma_bf16_ts_elect(tmem, a_bf + 8u, bl1, id16, 1u);
                    tc_commit_elect(&d_done[g]);
                }
                if (++g == TC_GROUPS) { g = 0; ++cyc; }
            }
        }
    } else {
        // ================= epilogue + finaliser: thread = portfolio row =================
        // Chunk by chunk, as soon as its MMAs completed: read the 32 final accumulator columns and the chunk's A stage
        // (l = hi + lo exactly) back from TMEM, release the stage, accumulate q = sum Y'_j l_j, sum l and l.mu; at the
        // end of the tile compute return / risk / Sharpe (app.py:708-711), track the selections, write the arrays.
        const int qd = warp - TC_EPI_WARP0, row = 32 * qd + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(32 * qd) << 16);
        float best_s = -Math<float>::inf(), best_d = -Math<float>::inf();
        uint64_t idx_s = MCP_NO_INDEX, idx_d = MCP_NO_INDEX;
        float rmin = Math<float>::inf(), rmax = -Math<float>::inf();
        unsigned int n_acc = 0;
        uint32_t g = 0, cyc = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            float2 q2 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f), r2 = make_float2(0.f, 0.f);     // .x: even assets, .y: odd assets
            // BOUNDS: w_i <= hi_i  <=>  e_i / hi_i <= s and w_i >= lo_i  <=>  e_i / lo_i >= s, on l = -e: running min of l / hi and
            // running max of l / lo (+ a term that removes unconstrained assets) -- one FFMA2 + one three-input min / max per pair
            float mn_hi = 0.f, mx_lo = -Math<float>::inf();
            for (int ci = 0; ci < C; ++ci) {
                const int c = C - 1 - ci;
                const uint32_t st = lane_base + TC_COL_A + TC_STAGE_COLS * g;
                mbar_wait(&d_done[g], cyc & 1u);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t y[16], hi[16], lo[16];
                    tmem_ld16(lane_base + (uint32_t)(TC_KC * c + 16 * h), y);
                    if constexpr (F16) {
                        tmem_ld16(st + 32u + 16u * h, hi);                  // the FP16 stage carries l itself
                    } else {
                        tmem_ld16(st + 16u * h, hi);
                        tmem_ld16(st + 32u + 16u * h, lo);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (h == 1) {                         // everything of this chunk is in registers: release stage and columns
                        tc_fence_before();
                        mbar_arrive(&a_free[g]);
                        if (ci == C - 1) mbar_arrive(drained);
                    }
                    const float4* mu4 = reinterpret_cast<const float4*>(sMu + TC_KC * c + 16 * h);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const float4 u = mu4[m];
                        float2 la, lb;
                        if constexpr (F16) {
                            la = make_float2(__uint_as_float(hi[4 * m]), __uint_as_float(hi[4 * m + 1]));
                            lb = make_float2(__uint_as_float(hi[4 * m + 2]), __uint_as_float(hi[4 * m + 3]));
                        } else {
                            // l = hi + lo exactly; packed FP32x2 throughout (one FFMA2 = two assets)
                            la = fma2(make_float2(__uint_as_float(hi[4 * m]), __uint_as_float(hi[4 * m + 1])), bcast2(1.0f),
                                      make_float2(__uint_as_float(lo[4 * m]), __uint_as_float(lo[4 * m + 1])));
                            lb = fma2(make_float2(__uint_as_float(hi[4 * m + 2]), __uint_as_float(hi[4 * m + 3])), bcast2(1.0f),
                                      make_float2(__uint_as_float(lo[4 * m + 2]), __uint_as_float(lo[4 * m + 3])));
                        }
                        q2 = fma2(make_float2(__uint_as_float(y[4 * m]), __uint_as_float(y[4 * m + 1])), la, q2);
                        q2 = fma2(make_float2(__uint_as_float(y[4 * m + 2]), __uint_as_float(y[4 * m + 3])), lb, q2);
                        s2 = fma2(la, bcast2(1.0f), s2);
                        s2 = fma2(lb, bcast2(1.0f), s2);
                        r2 = fma2(la, make_float2(u.x, u.y), r2);
                        r2 = fma2(lb, make_float2(u.z, u.w), r2);
                        if constexpr (BOUNDS) {
                            const float4 ih = reinterpret_cast<const float4*>(sInvHi + TC_KC * c + 16 * h)[m];
                            const float2 ta = fma2(la, make_float2(ih.x, ih.y), bcast2(0.f)), tb = fma2(lb, make_float2(ih.z, ih.w), bcast2(0.f));
                            mn_hi = fmin3(mn_hi, ta.x, ta.y);
                            mn_hi = fmin3(mn_hi, tb.x, tb.y);
                            if (a.has_lo) {
                                const float4 il = reinterpret_cast<const float4*>(sInvLo + TC_KC * c + 16 * h)[m];
                                const float4 bl = reinterpret_cast<const float4*>(sBiasLo + TC_KC * c + 16 * h)[m];
                                const float2 va = fma2(la, make_float2(il.x, il.y), make_float2(bl.x, bl.y));
                                const float2 vb = fma2(lb, make_float2(il.z, il.w), make_float2(bl.z, bl.w));
                                mx_lo = fmax3(mx_lo, va.x, va.y);
                                mx_lo = fmax3(mx_lo, vb.x, vb.y);
                            }
                        }
                    }
                }
                if (++g == TC_GROUPS) { g = 0; ++cyc; }
            }
            const uint64_t lrow = tile * TC_ROWS + (uint64_t)row;                // row of this launch
            const bool live = lrow < P;
            const uint64_t gi = (BOUNDS && listed) ? (live ? a.in_list[lrow] : 0ull) : a.first + lrow;
            const uint64_t local = BOUNDS ? gi - a.first : lrow;                   // row in the output arrays
            bool take = live;
            if constexpr (BOUNDS) {
                // w_i <= hi_i for all i  <=>  max e_i / hi_i <= sum(e), and the mirror image for the lower bounds.  This kernel's
                // sum(e) differs from the SIMT kernel's in the last bits, so only verdicts with a margin far above that (1e-5
                // relative against ~1e-7) are taken here; the rest is decided by the SIMT kernel, whose accept / skip decisions
                // and attempt numbers therefore hold for every row.
                const float s = -(s2.x + s2.y), up = s * (1.f + 1e-5f), dn = s * (1.f - 1e-5f);
                const float mx = -mn_hi, mn = -mx_lo;
                const bool inside = mx <= dn && (!a.has_lo || mn >= up);
                const bool outside = mx > up || (a.has_lo && mn < dn);
                const bool last = (int)a.attempt + 1 >= a.max_tries;
                const bool retry = live && outside && !last;
                const bool unsure = live && !inside && !outside;
                take = live && (inside || (outside && last && a.keep_last != 0));
                const bool skip = live && outside && last && a.keep_last == 0;     // app.py:706-707
                const unsigned long long rpos = warp_append(retry, a.retry_count, lane);
                if (retry) a.retry_list[rpos] = gi;
                const unsigned long long spos = warp_append(unsure, a.simt_count, lane);
                if (unsure) { a.simt_list[spos] = gi; a.simt_att[spos] = (uint16_t)a.attempt; }
                if (live && !take && a.inv_out) a.inv_out[local] = 0.f;            // not this row's final draw
                if (skip) {
                    const float nanv = Math<float>::nan();
                    if (a.ret_out) a.ret_out[local] = nanv;
                    if (a.risk_out) a.risk_out[local] = nanv;
                    if (a.sharpe_out) a.sharpe_out[local] = nanv;
                    if (a.acc_out) a.acc_out[local] = 0;
                }
            }
            if (take) {
                const bool supplied = a.w_in != nullptr;
                const float q = F16 ? (q2.x + q2.y) * a.qscale : q2.x + q2.y;
                const float s = supplied ? 1.f : -(s2.x + s2.y), r = supplied ? (r2.x + r2.y) : -(r2.x + r2.y);         // Philox rows hold l = -e
                float ret, risk, sharpe;
                metrics_from<float>(q, r, s, a.rf, supplied, ret, risk, sharpe);
                ++n_acc;
                // tiles ascend per thread: the first occurrence is kept (a list is in arrival order: ties compare indices there)
                if (sharpe > best_s || (BOUNDS && sharpe == best_s && gi < idx_s)) { best_s = sharpe; idx_s = gi; }
                const float d = -fabsf(risk - a.target);
                if (d > best_d || (BOUNDS && d == best_d && gi < idx_d)) { best_d = d; idx_d = gi; }
                rmin = fminf(rmin, risk);
                rmax = fmaxf(rmax, risk);
                if (a.ret_out) a.ret_out[local] = ret;
                if (a.risk_out) a.risk_out[local] = risk;
                if (a.sharpe_out) a.sharpe_out[local] = sharpe;
                if (a.acc_out) a.acc_out[local] = 1;
                if (a.inv_out) a.inv_out[local] = supplied ? 1.f : Math<float>::rcp(s);
            }
        }
        warp_argmax<float>(best_s, idx_s);
        warp_argmax<float>(best_d, idx_d);
        rmin = warp_min<float>(rmin);
        rmax = warp_max<float>(rmax);
        n_acc = __reduce_add_sync(0xffffffffu, n_acc);
        if (lane == 0) {
            sCand[qd] = PfCand{(double)best_s, idx_s, (double)best_d, idx_d, (double)rmin, (double)rmax};
            sAcc[qd] = n_acc;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    if (tid == 0) {
        PfCand b = sCand[0];
        unsigned long long acc = sAcc[0];
        for (int w = 1; w < 4; ++w) {
            const PfCand o = sCand[w];
            if (cand_better<double>(o.key_s, o.idx_s, b.key_s, b.idx_s)) { b.key_s = o.key_s; b.idx_s = o.idx_s; }
            if (cand_better<double>(o.key_d, o.idx_d, b.key_d, b.idx_d)) { b.key_d = o.key_d; b.idx_d = o.idx_d; }
            b.rmin = o.rmin < b.rmin ? o.rmin : b.rmin;
            b.rmax = o.rmax > b.rmax ? o.rmax : b.rmax;
            acc += sAcc[w];
        }
        a.cands[blockIdx.x] = b;
        if (acc) atomicAdd(a.n_accepted, acc);
    }
}

// w[p][i] *= inv[p]: turns the raw exponentials the sweep stored into weights (same e * rcp(sum e) as the SIMT kernels)
__global__ void __launch_bounds__(256) tc_scale_rows(float* w, const float* inv, uint64_t P, int n) {
    const uint64_t total = P * (uint64_t)n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) w[i] *= inv[i / (uint64_t)n];
}

// the same for the rows of one retry round: row = list[r] - first
__global__ void __launch_bounds__(256) tc_scale_rows_list(float* w, const float* inv, const uint64_t* list, const unsigned long long* count,
                                                           uint64_t first, int n) {
    const uint64_t total = (uint64_t)*count * (uint64_t)n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = list[i / (uint64_t)n] - first;
        w[row * (uint64_t)n + i % (uint64_t)n] *= inv[row];
    }
}

// ---- host side -------------------------------------------------------------------------------

static inline float tf32_round(float x) {          // round-to-nearest-even onto 10 explicit mantissa bits
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x00000fffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}
static inline uint16_t bf16_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x00007fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

bool pf_large_tc_eligible(const PfJob& job) {
    const char* v = getenv("MCP_LARGE_TC");               // "0" forces the SIMT kernel (A/B tests, benchmarks)
    if (v && v[0] == '0') return false;
    if (!(job.dtype == MCP_F32 && job.n > PF_SMALL_MAX_N && job.n <= TC_MAX_N)) return false;
    if (!job.bounds) return true;
    // bounds: Philox rows only.  The choice must not depend on the size of the launch: a call is cut into chunks and shards, and
    // which kernel evaluates a row has to be a property of the row (the two kernels agree to ~1e-7, not to the bit).  Replays of the
    // selected rows ask for the SIMT kernel explicitly (tc_bounds_route == 2).
    const char* b = getenv("MCP_LARGE_TC_BOUNDS");        // "0": bounded sweeps stay on the SIMT kernel (A/B tests)
    if (b && b[0] == '0') return false;
    return job.w_in == nullptr && job.idx_list == nullptr && job.tc_bounds_route != 2 && job.P <= TC_BOUNDED_MAX_P;
}

// FP16 split: Philox rows only (|lg2 U| <= 24 is inside FP16's range; supplied weights may be any FP32 value), MCP_LARGE_TC_F16=0 disables
static bool tc_use_f16(const PfJob& job) {
    const char* v = getenv("MCP_LARGE_TC_F16");
    if (v && v[0] == '0') return false;
    return job.w_in == nullptr;
}

template <bool F16, bool BOUNDS = false>
static int tc_launch(mcp_context* h, PfJob& job, const TcArgs& a, size_t table_bytes) {
    auto kern = job.rounds == 7 ? large_sweep_tc<F16, 7, BOUNDS> : large_sweep_tc<F16, 10, BOUNDS>;
    const size_t smem = table_bytes + (3 * TC_MAX_GROUPS + 3) * sizeof(uint64_t) + 4 * sizeof(PfCand) + 4 * sizeof(unsigned int) + 8 * sizeof(uint4) + 32;
    if (smem > h->prop.sharedMemPerBlockOptin)
        return mcp_fail(h, MCP_ERR_INVALID, "large_sweep_tc: N=%d needs %zu B of shared memory (max %zu)", job.n, smem, (size_t)h->prop.sharedMemPerBlockOptin);
    MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t n_tiles = (job.P + TC_ROWS - 1) / TC_ROWS;
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount, n_tiles);
    grid = std::max<uint64_t>(1, std::min<uint64_t>(grid, job.max_blocks));
    job.blocks_used = (int)grid;
    kern<<<(unsigned)grid, TcCfg<F16>::THREADS, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

// Lists of one bounded round (device memory): input rows (null = the contiguous range), rows to redraw, rows for the SIMT kernel
struct TcRound {
    const uint64_t* in_list = nullptr;
    const unsigned long long* in_count = nullptr;
    uint64_t* retry_list = nullptr;
    unsigned long long* retry_count = nullptr;
    uint64_t* simt_list = nullptr;
    uint16_t* simt_att = nullptr;
    unsigned long long* simt_count = nullptr;
    int attempt = 0;
    float* inv_out = nullptr;          // 1 / sum(e) scratch of the whole sub-range (allocated by the first round)
};

// One tcgen05 sweep over job's range.  round != null: the BOUNDS instance (FP16 split, Philox rows), one attempt of the round's rows.
static int tc_launch_range(mcp_context* h, PfJob& job, TcRound* round) {
    const bool bounded = round != nullptr;
    const int n = job.n, np = std::max(64, (n + TC_KC - 1) / TC_KC * TC_KC), C = np / TC_KC;
    const bool f16 = bounded || tc_use_f16(job);
    const int table_kind = bounded ? 2 : (f16 ? 1 : 0);
    const uint32_t lo_bytes = tc_lo_off(C), hi_bytes = f16 ? lo_bytes : tc_hi_off(C);
    const size_t table_bytes = (size_t)hi_bytes + lo_bytes + (size_t)np * (bounded ? 16 : 4);
    bool has_lo = false;
    if (bounded && job.lo)
        for (int i = 0; i < n; ++i) has_lo = has_lo || job.lo[i] > 0.0;
    // S'[k][j], j <= k: Sigma_kk on the diagonal, Sigma_kj + Sigma_jk below it (w' Sigma w = sum_k sum_{j<=k} w_k S'_kj w_j)
    auto s_prime = [&](int k, int j) { return j == k ? job.sigma[(size_t)k * n + k] : job.sigma[(size_t)k * n + j] + job.sigma[(size_t)j * n + k]; };
    // FP16 split: S' is stored times 2^e with the largest entry in [2^13, 2^14) (FP16 keeps 11 significant bits down to
    // 2^-14, i.e. over 28 binades below the largest entry; the second image holds the next 11 bits); q is multiplied by 2^-e
    double qscale = 1.0, sscale = 1.0;
    if (f16) {
        double smax = 0;
        for (int k = 0; k < n; ++k)
            for (int j = 0; j <= k; ++j) smax = std::max(smax, std::fabs(s_prime(k, j)));
        if (smax > 0 && std::isfinite(smax)) {
            int e = 0;
            std::frexp(smax, &e);                       // smax = m 2^e, m in [0.5, 1)
            e = std::min(120, std::max(-120, 14 - e));        // 2^e and 2^-e both normal FP32 numbers
            sscale = std::ldexp(1.0, e);
            qscale = std::ldexp(1.0, -e);
        }
    }
    unsigned char* dev = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 6, table_bytes, (void**)&dev));
    // once per mcp_portfolios call and operand split: later chunks and the replays reuse it
    if (job.tc_table_epoch == 0 || job.tc_table_epoch != h->const_epoch || job.tc_table_f16 != table_kind) {
        std::vector<unsigned char> host(table_bytes, 0);
        // Canonical K-major no-swizzle layout per chunk: [K core (4 tf32 / 8 16-bit)][N group of 8][8 rows x 16 bytes].
        for (int k = 0; k < n; ++k) {
            const int c = k / TC_KC, kk = k % TC_KC, Nc = TC_KC * (c + 1);
            for (int j = 0; j <= k; ++j) {
                const double v = s_prime(k, j);
                const size_t o16 = (size_t)tc_lo_off(c) + ((size_t)(kk / 8) * (Nc / 8) + j / 8) * 128 + (j % 8) * 16 + (kk % 8) * 2;
                if (f16) {
                    const double vs = v * sscale;
                    const __half s1 = __float2half_rn((float)vs);
                    const __half s2 = __float2half_rn((float)(vs - (double)__half2float(s1)));
                    memcpy(&host[o16], &s1, 2);
                    memcpy(&host[(size_t)hi_bytes + o16], &s2, 2);
                } else {
                    const float shi = tf32_round((float)v);
                    const uint16_t slo = bf16_round((float)(v - (double)shi));
                    const size_t oh = (size_t)tc_hi_off(c) + ((size_t)(kk / 4) * (Nc / 8) + j / 8) * 128 + (j % 8) * 16 + (kk % 4) * 4;
                    memcpy(&host[oh], &shi, 4);
                    memcpy(&host[(size_t)hi_bytes + o16], &slo, 2);
                }
            }
        }
        float* hmu = reinterpret_cast<float*>(host.data() + hi_bytes + lo_bytes);
        for (int i = 0; i < n; ++i) hmu[i] = (float)job.mu[i];
        if (bounded) {
            // w_i <= hi_i <=> e_i / hi_i <= sum(e);  w_i >= lo_i <=> e_i / lo_i >= sum(e).  A bound that cannot bite (hi >= 1, lo <= 0)
            // and the padded assets drop out: factor 0 for the maximum, factor 0 with a -1e30 term for the minimum.
            float* inv_hi = hmu + np;
            float* inv_lo = inv_hi + np;
            float* bias_lo = inv_lo + np;
            for (int i = 0; i < np; ++i) {
                inv_hi[i] = 0.f; inv_lo[i] = 0.f; bias_lo[i] = -1e30f;
                if (i >= n) continue;
                if (job.hi && job.hi[i] < 1.0) inv_hi[i] = job.hi[i] > 1e-30 ? (float)(1.0 / job.hi[i]) : 1e30f;
                if (job.lo && job.lo[i] > 0.0) { inv_lo[i] = (float)std::min(1.0 / job.lo[i], 1e30); bias_lo[i] = 0.f; }
            }
        }
        MCP_CUDA(h, cudaMemcpyAsync(dev, host.data(), table_bytes, cudaMemcpyHostToDevice, job.stream));
        MCP_CUDA(h, cudaStreamSynchronize(job.stream));        // `host` is pageable and dies at scope exit; other streams may read next
        job.tc_table_epoch = ++h->const_epoch;
        job.tc_table_f16 = table_kind;
    }

    TcArgs a;
    a.table = dev;
    a.w_in = (const float*)job.w_in;
    a.w_out = (float*)job.w_out;
    a.inv_out = nullptr;
    if (a.w_out) {    // two sweeps can be in flight in the HOST-space pipeline (one per side stream): one scratch each
        MCP_CHECK(mcp_dev_reserve(h, job.stream == h->side_stream[1] ? 13 : 12, (size_t)job.P * sizeof(float), (void**)&a.inv_out));
        if (bounded) round->inv_out = a.inv_out;
    }
    a.ret_out = (float*)job.ret_out; a.risk_out = (float*)job.risk_out; a.sharpe_out = (float*)job.sharpe_out;
    a.acc_out = job.acc_out; a.cands = job.cands; a.n_accepted = job.n_accepted;
    a.first = job.first; a.P = job.P; a.n = n; a.np = np;
    a.k0 = (uint32_t)job.seed; a.k1 = (uint32_t)(job.seed >> 32);
    philox_keys_fill(a.rk, job.seed);
    a.rf = (float)job.rf; a.target = (float)job.target;
    a.qscale = (float)qscale;
    a.in_list = nullptr; a.in_count = nullptr; a.retry_list = nullptr; a.retry_count = nullptr;
    a.simt_list = nullptr; a.simt_att = nullptr; a.simt_count = nullptr;
    a.attempt = 0; a.max_tries = job.max_tries; a.keep_last = job.keep_last;
    a.has_lo = has_lo ? 1 : 0;
    if (bounded) {
        a.in_list = round->in_list; a.in_count = round->in_count;
        a.retry_list = round->retry_list; a.retry_count = round->retry_count;
        a.simt_list = round->simt_list; a.simt_att = round->simt_att; a.simt_count = round->simt_count;
        a.attempt = (uint32_t)round->attempt;
        MCP_CHECK((tc_launch<true, true>(h, job, a, table_bytes)));
    } else {
        MCP_CHECK(f16 ? tc_launch<true>(h, job, a, table_bytes) : tc_launch<false>(h, job, a, table_bytes));
    }
    if (a.w_out && bounded && round->in_list) {
        tc_scale_rows_list<<<(unsigned)h->prop.multiProcessorCount * 4, 256, 0, job.stream>>>(a.w_out, a.inv_out, round->in_list, round->in_count, job.first, n);
        MCP_CUDA(h, cudaGetLastError());
        h->launches++;
    } else if (a.w_out) {
        const uint64_t total = job.P * (uint64_t)n;
        const unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)h->prop.multiProcessorCount * 8);
        tc_scale_rows<<<blocks, 256, 0, job.stream>>>(a.w_out, a.inv_out, job.P, n);
        MCP_CUDA(h, cudaGetLastError());
        h->launches++;
    }
    return MCP_OK;
}

int pf_large_launch_tc(mcp_context* h, PfJob& job) { return tc_launch_range(h, job, nullptr); }

// Bounds rejection (app.py:700-707) with the quadratic forms on the tensor cores.  Per sub-range of the launch, in rounds: round k
// evaluates attempt k of the rows still pending (round 0: the contiguous range, later: the previous round's retry list) on the
// BOUNDS instance, which accepts, skips (last attempt) or queues each row for the next round -- and hands the rows it cannot
// call (within 1e-5 of a bound) to the tiled SIMT kernel together with the attempt they are at.  Which kernel evaluates a row
// depends on that row's draws only -- never on how the range is chunked or sharded -- so results are reproducible across chunk
// sizes and GPU counts.  One 8-byte read-back per round (the retry count) ends the rounds when nothing is pending.
int pf_large_launch_tc_bounded(mcp_context* h, PfJob& job) {
    const uint64_t P = job.P, first = job.first;
    const size_t es = 4;
    constexpr uint64_t SUB = 1ull << 26;                  // rows per sub-range: three index lists of at most 512 MB each
    constexpr int MAX_SUBS = 40;
    const int side = job.stream == h->side_stream[1] ? 1 : 0;
    const uint64_t cap = std::min<uint64_t>(P, SUB);
    unsigned char* lbuf = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 21 + side, 256 + cap * (3 * 8 + 2) + 64, (void**)&lbuf));
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(lbuf);      // [0], [1]: ping-pong retry counts, [2]: SIMT rows
    uint64_t* d_retry[2] = {reinterpret_cast<uint64_t*>(lbuf + 256), reinterpret_cast<uint64_t*>(lbuf + 256) + cap};
    uint64_t* d_simt = d_retry[1] + cap;
    uint16_t* d_simt_att = reinterpret_cast<uint16_t*>(d_simt + cap);
    PfCand* const cands0 = job.cands;
    const int max_blocks0 = job.max_blocks;
    if (max_blocks0 < MAX_SUBS + h->prop.multiProcessorCount) return mcp_fail(h, MCP_ERR_INVALID, "bounded tcgen05 sweep: candidate scratch too small");
    if ((P + SUB - 1) / SUB > (uint64_t)MAX_SUBS) return mcp_fail(h, MCP_ERR_INVALID, "bounded tcgen05 sweep: %llu rows in one launch", (unsigned long long)P);
    void* const w0 = job.w_out; void* const r0p = job.ret_out; void* const k0p = job.risk_out; void* const s0p = job.sharpe_out;
    uint8_t* const a0 = job.acc_out;
    auto at = [&](void* q, uint64_t off, size_t row_bytes) -> void* { return q ? (unsigned char*)q + off * row_bytes : nullptr; };
    job.tc_bounds_route = 1;
    int n_sub = 0;
    int rc = MCP_OK;
    PfCand* const region = cands0 + MAX_SUBS;
    for (uint64_t off = 0; off < P && rc == MCP_OK; off += SUB, ++n_sub) {
        const uint64_t rows = std::min<uint64_t>(SUB, P - off);
        job.first = first + off;
        job.P = rows;
        job.w_out = at(w0, off, (size_t)job.n * es); job.ret_out = at(r0p, off, es); job.risk_out = at(k0p, off, es); job.sharpe_out = at(s0p, off, es);
        job.acc_out = (uint8_t*)at(a0, off, 1);
        job.cands = region;
        job.max_blocks = h->prop.multiProcessorCount;
        MCP_CUDA(h, cudaMemsetAsync(d_counts, 0, 24, job.stream));
        TcRound rd;
        rd.simt_list = d_simt; rd.simt_att = d_simt_att; rd.simt_count = d_counts + 2;
        unsigned long long pending = rows;
        int folded = 0;
        for (int attempt = 0; attempt < job.max_tries && pending > 0 && rc == MCP_OK; ++attempt) {
            const int cur = attempt & 1, nxt = cur ^ 1;
            rd.attempt = attempt;
            rd.in_list = attempt == 0 ? nullptr : d_retry[cur];
            rd.in_count = attempt == 0 ? nullptr : d_counts + cur;
            rd.retry_list = d_retry[nxt];
            rd.retry_count = d_counts + nxt;
            MCP_CUDA(h, cudaMemsetAsync(d_counts + nxt, 0, 8, job.stream));
            rc = tc_launch_range(h, job, &rd);
            if (rc != MCP_OK) break;
            rc = pf_reduce_launch(h, region, job.blocks_used, cands0 + n_sub, folded, job.stream);
            folded = 1;
            if (rc != MCP_OK) break;
            MCP_CUDA(h, cudaMemcpyAsync(&pending, d_counts + nxt, 8, cudaMemcpyDeviceToHost, job.stream));
            MCP_CUDA(h, cudaStreamSynchronize(job.stream));
        }
        if (rc != MCP_OK) break;
        job.idx_list = d_simt;
        job.idx_count = d_counts + 2;
        job.idx_attempt = d_simt_att;
        rc = pf_large_launch_list(h, job);
        job.idx_list = nullptr; job.idx_count = nullptr; job.idx_attempt = nullptr;
        if (rc != MCP_OK) break;
        rc = pf_reduce_launch(h, region, job.blocks_used, cands0 + n_sub, folded, job.stream);
    }
    job.first = first; job.P = P;
    job.w_out = w0; job.ret_out = r0p; job.risk_out = k0p; job.sharpe_out = s0p; job.acc_out = a0;
    job.cands = cands0; job.max_blocks = max_blocks0;
    job.blocks_used = n_sub;
    return rc;
}

}  // namespace mcp
