// Correlated-path simulator on the tensor cores for wide universes (FP32, Philox normals, 32 < N <= 256): two-stage variant with
// 16-bit split operands.  Same scheme as mcp_paths_tc.cu -- thread = path = TMEM lane, the step's normals go from registers
// straight into tensor memory as the MMA's A operand, the thread reads its row of R = Z Lp' back and compounds V in registers
// (np.cumprod(1 + returns), app.py:253) -- with the two changes the width asks for:
//
//   * FP16 operand split instead of TF32:  z = h1 + h2 (h1 = FP16(z), h2 = FP16(z - h1): 22 significant bits, |z| < 6 sits well
//     inside FP16's range), Lp 2^e = B1 + B2 (two FP16 images, 2^e puts the largest entry in [2^13, 2^14)), R 2^e = h1 B1 + h2 B1 +
//     h1 B2 with FP32 accumulation.  A 16-bit A operand packs two K elements per tensor-memory column, so a stage
//     [D: N | h1: N/2 | h2: N/2] is 2 N columns and TWO stages fit the 512 columns at N = 128 (one row tile) and at N = 64 (two
//     tiles); kind::f16 MMAs take K = 16 per instruction, 3 N / 16 per round instead of the TF32 kernel's 3 N / 8 + 1.
//   * two stages: the MMAs of round k run while the thread draws round k + 1, and the R row of step s is read back two steps
//     later, so a generator thread never waits for the tensor core in the steady state.  With one stage (the TF32 wide kernel)
//     the step time at N = 128 was the SUM of the generation and of the 49 MMAs of the tile-step.
//   The drift mu dt is added on the SIMT side (it no longer has a column to ride in), together with the 2^-e rescaling:
//   V <- V + V (d 2^-e + mu dt), two FFMA2 per asset pair.
//
// The normals are the ones every other path kernel and oracle/philox_np.py produce (chunk c of a step = Philox blocks 3c .. 3c+2,
// 24-bit fields, Box-Muller); K rounds and the block decomposition of 128 < N <= 256 work as in mcp_paths_tc.cu.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>

#include "mcp_device.cuh"
#include "mcp_paths.h"
#include "mcp_tcgen05.cuh"

namespace mcp {

constexpr int P16_ROWS = 128;
constexpr int P16_MAX_TILES = 2;

template <int NP> struct P16Cfg {
    static_assert(NP == 64 || NP == 128, "padded asset counts of the 16-bit wide path kernel");
    static constexpr uint32_t STAGE_COLS = 2 * NP, COL_D = 0, COL_H1 = NP, COL_H2 = NP + NP / 2, TILE_COLS = 2 * STAGE_COLS;
    static constexpr uint32_t IMG_BYTES = NP * NP * 2;             // one FP16 image of one K round: K = NP rows x N = NP columns
    static constexpr uint32_t ROUND_BYTES = 2 * IMG_BYTES;         // B1 | B2
    static constexpr uint32_t LBO = (NP / 8) * 128, SBO = 128;     // canonical K-major no-swizzle core matrices (8 rows x 16 bytes)
    static constexpr uint32_t DESC_STEP = (2 * LBO) >> 4;          // descriptor start-address increment per K = 16 step
    static constexpr int TILES = 512 / TILE_COLS;
};

template <int NP>
struct P16Args {
    const unsigned char* table;          // global: per K round B1 image | B2 image
    float w[NP];                         // portfolio weights of the tile's assets (0 for padded ones)
    float drift[NP];                     // mu dt of the tile's assets
    float qs;                            // 2^-e: undoes the scaling of the B images
    float x0;                            // constant of the terminal sum (-1 for a whole universe, 0 for a block's part)
    float* terminal;
    unsigned long long* hist0;
    uint64_t first, M;
    int n_steps, k_rounds;
    PhiloxKeys rk;
};

__device__ __forceinline__ void p16_wait_idle(uint64_t* b, uint32_t parity) {      // long suspend hint: a waiting warp must not compete for issue slots
    const uint32_t addr = smem_u32(b);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
    }
}

// GG generator groups of 128 threads per row tile: thread (g, row) -- the same TMEM lane in every group -- draws the chunks
// g CPG .. g CPG + CPG - 1 of each K round and owns the assets g VS .. g VS + VS - 1 of the row (it reads that slice of the
// accumulator back and compounds it in VS registers).  With the MMAs off the critical path (two stages) the step time is the
// generation, and one thread per path leaves an SM with four generator warps at N = 128: more groups = more warps to hide the
// Philox / MUFU latencies with.  The groups' parts of w . V_T meet in shared memory; group 0 adds them in a fixed order.
template <int NP, int GG, int ROUNDS>
__global__ void __launch_bounds__(P16Cfg<NP>::TILES* GG* P16_ROWS + 32 * P16Cfg<NP>::TILES, 1) path_kernel_tc16(const __grid_constant__ P16Args<NP> a) {
    using Cfg = P16Cfg<NP>;
    constexpr int TILES = Cfg::TILES, GEN_WARPS = 4 * TILES * GG, NCH = NP / 16, CPG = NCH / GG, VS = NP / GG;
    constexpr int NBINS = 1 << MCP_SEL_BITS;
    static_assert(NCH % GG == 0 && VS % 32 == 0, "a thread owns a whole number of chunks and 32-column pieces of the accumulator");
    extern __shared__ __align__(128) unsigned char smem[];
    const int KR = a.k_rounds;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)KR * Cfg::ROUND_BYTES);
    uint64_t* full = bars;                               // [TILES][2] 128 arrivals: the A stage holds a round's normals
    uint64_t* done = bars + 2 * P16_MAX_TILES;           // [TILES][2] commit: the MMAs that read the A stage are complete
    uint64_t* dready = bars + 4 * P16_MAX_TILES;         // [TILES][2] commit after a step's last round: its R rows are complete
    uint64_t* table_bar = bars + 6 * P16_MAX_TILES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 * P16_MAX_TILES + 1);
    float* sPart = reinterpret_cast<float*>(bars + 6 * P16_MAX_TILES + 2);          // [TILES][GG][128] parts of w . V_T (GG > 1)
    unsigned int* sHist = reinterpret_cast<unsigned int*>(sPart + (GG > 1 ? TILES * GG * P16_ROWS : 0));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = (int)blockDim.x;
    if (tid == 0) {
        for (int t = 0; t < 2 * TILES; ++t) { mbar_init(&full[t], P16_ROWS * GG); mbar_init(&done[t], 1); mbar_init(&dready[t], 1); }
        mbar_init(table_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)KR * Cfg::ROUND_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(table_bar)), "r"(bytes) : "memory");
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t part = bytes - off < 32768u ? bytes - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + off)), "l"(a.table + off), "r"(part), "r"(smem_u32(table_bar)) : "memory");
        }
    }
    if (warp == GEN_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (a.hist0 != nullptr)
        for (int i = tid; i < NBINS; i += nthreads) sHist[i] = 0u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(table_bar, 0u);
    const uint32_t tmem = *tmem_slot;

    const uint64_t n_tiles = (a.M + P16_ROWS - 1) / P16_ROWS;
    const uint64_t slots = (uint64_t)gridDim.x * TILES;
    const uint32_t S = (uint32_t)a.n_steps;

    if (warp < GEN_WARPS) {
        // generator warp w: tile t0 = w / (4 GG), group g = (w / 4) % GG, TMEM lane quadrant w % 4
        const int t0 = warp / (4 * GG), g = (warp >> 2) % GG, row = 32 * (warp & 3) + lane;
        const uint32_t tile_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + Cfg::TILE_COLS * (uint32_t)t0;
        float* part = sPart + (size_t)t0 * GG * P16_ROWS;
        const float kPi = 3.14159265358979323846f;
        const uint32_t one_bits = opaque_u32(0x3f800000u), two_bits = opaque_u32(0x40000000u);
        uint32_t ph = 0, gs = 0;                          // rounds / steps this slot has published (across its tiles)
        for (uint64_t tile = (uint64_t)blockIdx.x * TILES + t0; tile < n_tiles; tile += slots) {
            const uint64_t m = tile * P16_ROWS + (uint64_t)row;
            const uint64_t gidx = a.first + m;
            const uint32_t c0 = (uint32_t)gidx, c1 = (uint32_t)(gidx >> 32);
            float2 V[VS / 2];                             // this thread's assets g VS .. g VS + VS - 1
#pragma unroll
            for (int i = 0; i < VS / 2; ++i) V[i] = make_float2(1.f, 1.f);
            // chunk c of step s: 16 normals as 8 + 8 packed FP16 pairs.  Packed column 2q holds normals (4q, 4q + 2), column 2q + 1
            // normals (4q + 1, 4q + 3): the K order of ptc_normal_of_col, which the B images follow.
            auto draw = [&](uint32_t s, int c, uint32_t (&h1)[8], uint32_t (&h2)[8]) {
                uint32_t f[16];
                philox_fields<16, ROUNDS>(c0, c1, s, STREAM_NORMALS + (uint32_t)(3 * c), a.rk, f);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f1 = make_float2(__uint_as_float(mant_or(f[4 * q], one_bits)), __uint_as_float(mant_or(f[4 * q + 2], one_bits)));
                    const float2 u1 = fma2(f1, bcast2(-1.0f), bcast2(2.0f));
                    const float2 r = make_float2(Math<float>::sqrt(-Math<float>::lg2(u1.x)), Math<float>::sqrt(-Math<float>::lg2(u1.y)));
                    const float2 f2 = make_float2(__uint_as_float(mant_or(f[4 * q + 1], two_bits)), __uint_as_float(mant_or(f[4 * q + 3], two_bits)));
                    const float2 th = fma2(f2, bcast2(kPi), bcast2(-3.0f * kPi));
                    const float2 cs = make_float2(Math<float>::cosf_(th.x), Math<float>::cosf_(th.y));
                    const float2 sn = make_float2(Math<float>::sinf_(th.x), Math<float>::sinf_(th.y));
                    const float2 zc = fma2(r, cs, bcast2(0.0f)), zs = fma2(r, sn, bcast2(0.0f));
                    const __half2 pc = __floats2half2_rn(zc.x, zc.y), ps = __floats2half2_rn(zs.x, zs.y);        // low half = even K element
                    const float2 rc = fma2(__half22float2(pc), bcast2(-1.0f), zc), rs = fma2(__half22float2(ps), bcast2(-1.0f), zs);   // exact
                    const __half2 qc = __floats2half2_rn(rc.x, rc.y), qz = __floats2half2_rn(rs.x, rs.y);
                    h1[2 * q] = *reinterpret_cast<const uint32_t*>(&pc);
                    h1[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&ps);
                    h2[2 * q] = *reinterpret_cast<const uint32_t*>(&qc);
                    h2[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&qz);
                }
            };
            auto compound = [&](uint32_t gstep) {        // this thread's slice of the R row of (global) step gstep: V <- V + V (d 2^-e + mu dt)
                mbar_wait(&dready[2 * t0 + (gstep & 1u)], (gstep >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int b = 0; b < VS / 32; ++b) {
                    uint32_t d[32];
                    tmem_ld32(tile_base + (gstep & 1u) * Cfg::STAGE_COLS + Cfg::COL_D + (uint32_t)(VS * g + 32 * b), d);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int k = 16 * b + i;
                        float2 dr;
                        if constexpr (GG == 1) dr = make_float2(a.drift[2 * k], a.drift[2 * k + 1]);        // compile-time index: constant bank operand
                        else dr = make_float2(a.drift[VS * g + 2 * k], a.drift[VS * g + 2 * k + 1]);
                        const float2 r = fma2(make_float2(__uint_as_float(d[2 * i]), __uint_as_float(d[2 * i + 1])), bcast2(a.qs), dr);
                        V[k] = fma2(V[k], r, V[k]);
                    }
                }
            };
            for (uint32_t s = 0; s < S; ++s, ++gs) {
                if (s >= 2u) compound(gs - 2u);          // long complete: frees the accumulator this step's first round overwrites
                for (int r = 0; r < KR; ++r, ++ph) {
                    const uint32_t st = ph & 1u;
                    if (ph >= 2u) {                      // the A stage was last read by the MMAs of round ph - 2
                        mbar_wait(&done[2 * t0 + st], ((ph >> 1) - 1u) & 1u);
                        tc_fence_after();
                    }
                    const uint32_t acol = tile_base + st * Cfg::STAGE_COLS;
#pragma unroll
                    for (int c = g * CPG; c < g * CPG + CPG; ++c) {
                        uint32_t h1[8], h2[8];
                        draw(s, r * NCH + c, h1, h2);
                        tmem_st8(acol + Cfg::COL_H1 + 8u * (uint32_t)c, h1);
                        tmem_st8(acol + Cfg::COL_H2 + 8u * (uint32_t)c, h2);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    mbar_arrive(&full[2 * t0 + st]);
                }
            }
            if (S >= 2u) compound(gs - 2u);
            compound(gs - 1u);
            float2 x2 = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < VS / 2; ++i) {
                float2 wv;
                if constexpr (GG == 1) wv = make_float2(a.w[2 * i], a.w[2 * i + 1]);
                else wv = make_float2(a.w[VS * g + 2 * i], a.w[VS * g + 2 * i + 1]);
                x2 = fma2(wv, V[i], x2);
            }
            if constexpr (GG > 1) {                      // the groups' parts meet in shared memory; group 0 adds them in ascending order
                part[g * P16_ROWS + row] = x2.x + x2.y;
                asm volatile("bar.sync %0, %1;" ::"r"(1 + t0), "r"(GG * P16_ROWS) : "memory");
            }
            if (g == 0) {
                float x = a.x0 + (x2.x + x2.y);
                if constexpr (GG > 1) {
#pragma unroll
                    for (int k = 1; k < GG; ++k) x += part[k * P16_ROWS + row];
                }
                if (m < a.M) {
                    a.terminal[m] = x;
                    if (a.hist0 != nullptr) {             // first radix digit of the order-preserving key (mcp_quantile.cu, pass 0)
                        const unsigned digit = f32_to_key(__float_as_uint(x)) >> (32 - MCP_SEL_BITS);
                        const unsigned act = __activemask();
                        const unsigned peers = __match_any_sync(act, digit);
                        if (lane == __ffs(peers) - 1) atomicAdd(&sHist[digit], (unsigned)__popc(peers));
                    }
                }
            }
            if constexpr (GG > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + t0), "r"(GG * P16_ROWS) : "memory");   // parts are read before the next tile writes them
        }
    } else {
        // MMA issue warp of tile t: per round R 2^e (+)= h1 B1 + h2 B1 + h1 B2, K = 16 per tcgen05.mma (kind::f16)
        const int t = warp - GEN_WARPS;
        const uint32_t idesc = tc_idesc(0u, (uint32_t)NP);           // FP16 operands, FP32 accumulate, M = 128, N = NP
        const uint32_t base = tmem + Cfg::TILE_COLS * (uint32_t)t;
        uint32_t ph = 0, gs = 0;
        for (uint64_t tile = (uint64_t)blockIdx.x * TILES + (uint64_t)t; tile < n_tiles; tile += slots) {
            for (uint32_t s = 0; s < S; ++s, ++gs) {
                const uint32_t dcol = base + (gs & 1u) * Cfg::STAGE_COLS + Cfg::COL_D;
                for (int r = 0; r < KR; ++r, ++ph) {
                    const uint32_t st = ph & 1u, acol = base + st * Cfg::STAGE_COLS;
                    const uint32_t img = smem_u32(smem) + (uint32_t)r * Cfg::ROUND_BYTES;
                    const uint64_t b1 = tc_sdesc(img, Cfg::LBO, Cfg::SBO), b2 = tc_sdesc(img + Cfg::IMG_BYTES, Cfg::LBO, Cfg::SBO);
                    p16_wait_idle(&full[2 * t + st], (ph >> 1) & 1u);
                    tc_fence_after();
#pragma unroll 1
                    for (uint32_t k = 0; k < (uint32_t)(NP / 16); ++k)     // a step's first round overwrites the accumulator, later rounds add to it
                        mma_bf16_ts_elect(dcol, acol + Cfg::COL_H1 + 8u * k, b1 + (uint64_t)(k * Cfg::DESC_STEP), idesc, (k > 0 || r > 0) ? 1u : 0u);
#pragma unroll 1
                    for (uint32_t k = 0; k < (uint32_t)(NP / 16); ++k)
                        mma_bf16_ts_elect(dcol, acol + Cfg::COL_H2 + 8u * k, b1 + (uint64_t)(k * Cfg::DESC_STEP), idesc, 1u);
#pragma unroll 1
                    for (uint32_t k = 0; k < (uint32_t)(NP / 16); ++k)
                        mma_bf16_ts_elect(dcol, acol + Cfg::COL_H1 + 8u * k, b2 + (uint64_t)(k * Cfg::DESC_STEP), idesc, 1u);
                    tc_commit_elect(&done[2 * t + st]);
                    if (r == KR - 1) tc_commit_elect(&dready[2 * t + (gs & 1u)]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == GEN_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    if (a.hist0 != nullptr) {
        for (int i = tid; i < NBINS; i += nthreads) {
            const unsigned c = sHist[i];
            if (c) atomicAdd(&a.hist0[i], (unsigned long long)c);
        }
    }
}

// ---- host side -------------------------------------------------------------------------------

__host__ __device__ constexpr int p16_normal_of_col(int kc) { return (kc & ~3) + ((kc & 3) == 1 ? 2 : (kc & 3) == 2 ? 1 : (kc & 3)); }

template <int NP, int GG, int ROUNDS>
static int p16_launch_t(mcp_context* h, PathJob& job, const P16Args<NP>& a) {
    using Cfg = P16Cfg<NP>;
    auto kern = path_kernel_tc16<NP, GG, ROUNDS>;
    const size_t smem = (size_t)a.k_rounds * Cfg::ROUND_BYTES + (6 * P16_MAX_TILES + 2) * sizeof(uint64_t) +
                        (GG > 1 ? (size_t)Cfg::TILES * GG * P16_ROWS * sizeof(float) : 0) + (a.hist0 ? sizeof(unsigned int) << MCP_SEL_BITS : 0) + 128;
    if (smem > h->prop.sharedMemPerBlockOptin)
        return mcp_fail(h, MCP_ERR_INVALID, "path_kernel_tc16: %d K rounds need %zu B of shared memory (max %zu)", a.k_rounds, smem, (size_t)h->prop.sharedMemPerBlockOptin);
    MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t n_tiles = (job.M + P16_ROWS - 1) / P16_ROWS;
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount, (n_tiles + Cfg::TILES - 1) / Cfg::TILES);
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, Cfg::TILES * GG * P16_ROWS + 32 * Cfg::TILES, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

__global__ void __launch_bounds__(256) p16_combine_blocks(const float* __restrict__ partial, int n_blocks, uint64_t M, float* __restrict__ terminal) {
    for (uint64_t m = (uint64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (uint64_t)gridDim.x * 256) {
        float x = -1.0f;
        for (int b = 0; b < n_blocks; ++b) x += partial[(size_t)b * M + m];
        terminal[m] = x;
    }
}

// Blocks of NP output assets (one block when n <= 128).  Block b reaches the normals 0 .. NP (b + 1) - 1: b + 1 K rounds.
template <int NP>
static int p16_launch_blocks(mcp_context* h, PathJob& job) {
    using Cfg = P16Cfg<NP>;
    const int n = job.n, B = (n + NP - 1) / NP;
    const std::vector<double>& L = *job.L;
    const double sdt = std::sqrt(job.dt), c = std::sqrt(2.0 * std::log(2.0));     // the kernel's normals come out divided by sqrt(2 ln 2)
    double vmax = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) vmax = std::max(vmax, std::fabs(L[(size_t)i * n + j] * sdt * c));
    double scale = 1.0, qs = 1.0;
    if (vmax > 0 && std::isfinite(vmax)) {
        int e = 0;
        std::frexp(vmax, &e);                           // vmax = m 2^e, m in [0.5, 1)
        e = std::min(100, std::max(-100, 14 - e));      // largest entry in [2^13, 2^14): FP16 keeps 11 bits over 28 binades below it
        scale = std::ldexp(1.0, e);
        qs = std::ldexp(1.0, -e);
    }
    std::vector<size_t> off(B + 1, 0);
    for (int b = 0; b < B; ++b) off[b + 1] = off[b] + (size_t)(b + 1) * Cfg::ROUND_BYTES;
    std::vector<unsigned char> host(off[B], 0);
    for (int b = 0; b < B; ++b) {
        for (int r = 0; r <= b; ++r) {
            const size_t img = off[b] + (size_t)r * Cfg::ROUND_BYTES;
            for (int i = 0; i < NP; ++i) {
                const int gi = NP * b + i;
                if (gi >= n) continue;
                for (int kc = 0; kc < NP; ++kc) {
                    const int j = NP * r + p16_normal_of_col(kc);
                    if (j > gi || j >= n) continue;                                  // L is lower triangular
                    const double v = L[(size_t)gi * n + j] * sdt * c * scale;
                    const __half b1 = __float2half_rn((float)v);
                    const __half b2 = __float2half_rn((float)(v - (double)__half2float(b1)));
                    const size_t o = ((size_t)(kc / 8) * (NP / 8) + i / 8) * 128 + (i % 8) * 16 + (kc % 8) * 2;
                    memcpy(&host[img + o], &b1, 2);
                    memcpy(&host[img + Cfg::IMG_BYTES + o], &b2, 2);
                }
            }
        }
    }
    unsigned char* dev = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 6, host.size(), (void**)&dev));
    ++h->const_epoch;
    MCP_CUDA(h, cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, job.stream));
    MCP_CUDA(h, cudaStreamSynchronize(job.stream));        // `host` is pageable and dies at scope exit
    float* partial = nullptr;
    if (B > 1) MCP_CHECK(mcp_dev_reserve(h, 23, (size_t)B * job.M * sizeof(float), (void**)&partial));
    for (int b = 0; b < B; ++b) {
        P16Args<NP> a;
        a.table = dev + off[b];
        for (int i = 0; i < NP; ++i) {
            const int gi = NP * b + i;
            a.w[i] = gi < n ? (float)job.w[gi] : 0.f;
            a.drift[i] = gi < n ? (float)(job.mu[gi] * job.dt) : 0.f;
        }
        a.qs = (float)qs;
        a.x0 = B > 1 ? 0.0f : -1.0f;
        a.terminal = B > 1 ? partial + (size_t)b * job.M : (float*)job.terminal;
        a.hist0 = B > 1 ? nullptr : job.hist0;
        a.first = job.first;
        a.M = job.M;
        a.n_steps = job.n_steps;
        a.k_rounds = b + 1;
        philox_keys_fill(a.rk, job.seed);
        static const int gg_env = getenv("MCP_PATHS_TC_GG") ? atoi(getenv("MCP_PATHS_TC_GG")) : 0;      // tuning knob: generator groups per tile
        constexpr int GMAX = NP == 128 ? 4 : 2;
        // measured on B200 (252 steps): N = 128: 1 / 2 / 4 groups 4.76 / 6.12 / 6.95e9 path-steps/s; N = 64: 1 / 2 groups 1.26 / 1.49e10;
        // N = 256 (two blocks of 128): 1.69 / 2.13 / 2.46e9
        const int gg = gg_env == 1 || gg_env == 2 || gg_env == GMAX ? gg_env : GMAX;
        if (gg == 1) MCP_CHECK(job.rounds == 7 ? (p16_launch_t<NP, 1, 7>(h, job, a)) : (p16_launch_t<NP, 1, 10>(h, job, a)));
        else if (gg == 2) MCP_CHECK(job.rounds == 7 ? (p16_launch_t<NP, 2, 7>(h, job, a)) : (p16_launch_t<NP, 2, 10>(h, job, a)));
        else MCP_CHECK(job.rounds == 7 ? (p16_launch_t<NP, GMAX, 7>(h, job, a)) : (p16_launch_t<NP, GMAX, 10>(h, job, a)));
    }
    if (B > 1) {
        const unsigned blocks = (unsigned)std::min<uint64_t>((job.M + 255) / 256, (uint64_t)h->prop.multiProcessorCount * 8);
        p16_combine_blocks<<<blocks, 256, 0, job.stream>>>(partial, B, job.M, (float*)job.terminal);
        MCP_CUDA(h, cudaGetLastError());
        h->launches++;
    }
    job.hist0_filled = B == 1 && job.hist0 != nullptr;
    return MCP_OK;
}

bool path_tc16_enabled() {
    const char* v = getenv("MCP_PATHS_TC_WIDE16");        // "0": the one-stage TF32-split kernel of mcp_paths_tc.cu (A/B measurements)
    return !(v && v[0] == '0');
}

int path_launch_tc16(mcp_context* h, PathJob& job) {
    if (job.n <= 64) return p16_launch_blocks<64>(h, job);
    return p16_launch_blocks<128>(h, job);                 // 128 < N <= 256: two blocks of 128 assets, the second with two K rounds
}

}  // namespace mcp
