// Correlated-path simulator on the tensor cores (FP32, Philox normals, N <= 32 assets): north-star row a10, config C4.
//
//   per step  r = mu dt + sqrt(dt) L z,  z ~ N(0, I_N);  V_i *= (1 + r_i)  (np.cumprod(1 + returns), app.py:253);  terminal = w . V_T - 1
//
// The step's return vector of 128 paths is a GEMM: R [128 x N] = Z [128 x N] Lp' + 1 drift', Lp = L sqrt(dt).  In the SIMT kernel
// (mcp_paths.cu) that contraction is 136 of the 184 FFMA2 a thread issues per step and path pair; here it runs on tcgen05:
//
//   * thread = path = TMEM lane.  The thread draws its step's normals in registers (Philox4x32 -> Box-Muller on the MUFU pipe,
//     two normal pairs per packed FP32x2 instruction) and writes them with tcgen05.st straight into TENSOR MEMORY as the MMA's A
//     operand; they never exist in shared or global memory.
//   * FP32-class accuracy from TF32 operands, FP32 accumulation: the tensor core reads the top 19 bits of a 32-bit A element
//     (truncation, measured by tools/tc_probe), so A = z itself plus zlo = z - trunc(z) (exact), B = Lhi + Llo (TF32 rounding and
//     its remainder):  R = z Lhi + zlo Lhi + z Llo  -- the dropped zlo Llo term is 2^-22 relative.  The drift rides in the same
//     GEMM: a constant A block [1 1 1 0 ...] against three B rows holding mu dt split into three TF32 pieces (exact to FP32).
//   * B (2.5 KB for N = 16) is fetched once per CTA by a TMA bulk copy into the canonical K-major no-swizzle core-matrix layout.
//   * one MMA-issue warp per CTA serves TILES row tiles of 128 paths; per tile and step it waits for the tile's A stage (mbarrier,
//     128 arrivals), issues 3 N/8 + 1 MMAs of shape 128 x N x 8 (kind::tf32) and commits to the tile's `done` barrier.  The
//     generator thread meanwhile draws the NEXT step's normals; then it reads its row of R back (tcgen05.ld), compounds V in
//     registers (8 FFMA2 for 16 assets) and publishes the next A stage.  A and D are single-buffered: the MMA of step s overlaps
//     the generation of step s + 1, which is an order of magnitude longer.
//   * epilogue: terminal value to global memory (4 B per path, the kernel's only HBM traffic) and, optionally, the first radix
//     histogram of the VaR select (warp-aggregated shared-memory atomics), so mcp_paths_stats starts at pass 1.
//
// The normals are the ones the SIMT kernels and oracle/philox_np.py produce (same counter layout and bit -> float construction);
// only the contraction differs (TF32-split products accumulated in FP32 instead of an FFMA chain): ~1e-6 relative on the terminal
// value, tested against the FP64 oracle in tests/test_paths_gpu.py.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mcp_device.cuh"
#include "mcp_paths.h"
#include "mcp_tcgen05.cuh"

namespace mcp {

constexpr int PTC_ROWS = 128;                // paths per tile = TMEM lanes
constexpr int PTC_MAX_TILES = 7;             // 7 x 128 generator threads + up to four MMA-issue warps <= 1024
constexpr int PTC_MAX_MMA_WARPS = 4;

// Tensor-memory layout of one row tile: STAGES x [D (NP accumulator columns) | z (NP) | zlo (NP)] then the constant block (8).
// STAGES = 2 lets a generator publish step s while the MMAs of step s - 1 are still in flight: it only ever waits for the
// MMAs of step s - 2, which finished long ago -- the issue latency of the MMA warp (~0.5 us per tile-step) leaves the
// critical path.  STAGES = 1 fits more tiles (more resident generator warps) but exposes that latency.
template <int NP, int STAGES> struct PtcCfg {
    static_assert(NP == 16 || NP == 32 || NP == 64 || NP == 128, "padded asset counts of the tensor-core path kernels");
    static_assert(STAGES == 1 || STAGES == 2, "one or two A / D stages per tile");
    static constexpr int KB = NP + 8;                        // K extent of the hi image: NP normals + the constant block (drift)
    static constexpr uint32_t STAGE_COLS = 3 * NP, COL_D = 0, COL_Z = NP, COL_ZLO = 2 * NP, COL_ONE = STAGES * STAGE_COLS;
    static constexpr uint32_t TILE_COLS = STAGES * STAGE_COLS + 8;
    static constexpr int MAX_TILES = (512 / TILE_COLS) < PTC_MAX_TILES ? (512 / TILE_COLS) : PTC_MAX_TILES;
    static constexpr uint32_t HI_BYTES = (KB / 4) * (NP / 8) * 128, LO_BYTES = (NP / 4) * (NP / 8) * 128;
    static constexpr uint32_t LBO = (NP / 8) * 128, SBO = 128;
    static constexpr uint32_t DESC_STEP = (2 * LBO) >> 4;    // descriptor start-address increment per K = 8 step
};

template <int NP>
struct PtcArgs {
    const unsigned char* table;          // global: Bhi image | Blo image (canonical UMMA layout), see path_launch_tc
    float w[NP];                         // portfolio weights (0 for padded assets)
    float* terminal;                     // [M]
    unsigned long long* hist0;           // [1 << MCP_SEL_BITS] or null
    uint64_t first, M;
    int n_steps;
    // path_kernel_tc_wide only: the tile's outputs are one BLOCK of NP assets of a wider universe whose K extent (the normals
    // 0 .. k_rounds NP - 1 its rows of L reach) is worked off in k_rounds rounds of NP normals; x0 = the constant of the
    // terminal sum (-1 for a whole universe, 0 for a block's partial sum)
    int k_rounds;
    float x0;
    PhiloxKeys rk;
};

// A column kc of the stage holds normal pi(kc): the packed Box-Muller produces (cos, cos) and (sin, sin) register pairs for two
// normal pairs at a time, so the natural register order is z[4q], z[4q+2], z[4q+1], z[4q+3]; B's K index is permuted to match.
__host__ __device__ constexpr int ptc_normal_of_col(int kc) { return (kc & ~3) + ((kc & 3) == 1 ? 2 : (kc & 3) == 2 ? 1 : (kc & 3)); }

// mbarrier wait for a warp that has nothing else to do: a long suspend hint, so that the wait costs a few instructions and not a
// spin that competes with the generator warps for issue slots
__device__ __forceinline__ void mbar_wait_idle(uint64_t* b, uint32_t parity) {
    const uint32_t addr = smem_u32(b);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
    }
}

// One tile-step under a single elect: R = z Lhi (overwrites) + zlo Lhi + z Llo + [1 1 1 0 ..] drift, K = 8 per tcgen05.mma
// (kind::tf32), then the commit.  bhi0 / blo0: shared-memory descriptors of the first K step of the two images; later K steps
// advance the descriptor's start address by DESC_STEP.
template <int NP>
__device__ __forceinline__ void ptc_issue(uint32_t d, uint32_t az, uint32_t azl, uint32_t a1, uint64_t bhi0, uint64_t blo0, uint32_t idesc, uint32_t bar);

template <>
__device__ __forceinline__ void ptc_issue<16>(uint32_t d, uint32_t az, uint32_t azl, uint32_t a1, uint64_t bhi0, uint64_t blo0, uint32_t idesc, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q, t, f;\n\t.reg .b32 az8, zl8;\n\t.reg .b64 h1, h2, l1;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "setp.eq.b32 t, 0, 0;\n\tsetp.ne.b32 f, 0, 0;\n\t"
                 "add.u32 az8, %1, 8;\n\tadd.u32 zl8, %2, 8;\n\t"
                 "add.u64 h1, %4, 32;\n\tadd.u64 h2, %4, 64;\n\tadd.u64 l1, %5, 32;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %4, %6, f;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [az8], h1, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %4, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [zl8], h1, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %5, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [az8], l1, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%3], h2, %6, t;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}"
                 ::"r"(d), "r"(az), "r"(azl), "r"(a1), "l"(bhi0), "l"(blo0), "r"(idesc), "r"(bar) : "memory");
}

template <>
__device__ __forceinline__ void ptc_issue<32>(uint32_t d, uint32_t az, uint32_t azl, uint32_t a1, uint64_t bhi0, uint64_t blo0, uint32_t idesc, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q, t, f;\n\t.reg .b32 z1, z2, z3, y1, y2, y3;\n\t.reg .b64 h1, h2, h3, h4, l1, l2, l3;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "setp.eq.b32 t, 0, 0;\n\tsetp.ne.b32 f, 0, 0;\n\t"
                 "add.u32 z1, %1, 8;\n\tadd.u32 z2, %1, 16;\n\tadd.u32 z3, %1, 24;\n\t"
                 "add.u32 y1, %2, 8;\n\tadd.u32 y2, %2, 16;\n\tadd.u32 y3, %2, 24;\n\t"
                 "add.u64 h1, %4, 64;\n\tadd.u64 h2, %4, 128;\n\tadd.u64 h3, %4, 192;\n\tadd.u64 h4, %4, 256;\n\t"
                 "add.u64 l1, %5, 64;\n\tadd.u64 l2, %5, 128;\n\tadd.u64 l3, %5, 192;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %4, %6, f;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [z1], h1, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [z2], h2, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [z3], h3, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %4, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [y1], h1, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [y2], h2, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [y3], h3, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %5, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [z1], l1, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [z2], l2, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [z3], l3, %6, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%3], h4, %6, t;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}"
                 ::"r"(d), "r"(az), "r"(azl), "r"(a1), "l"(bhi0), "l"(blo0), "r"(idesc), "r"(bar) : "memory");
}

// Issuing one tcgen05.mma costs the issuing warp tens of cycles whatever the MMA's shape (tools/tc_probe/tc_chain.cu: the operands
// travel through uniform registers): with per-MMA issue code ONE warp capped an SM at one 7-MMA tile-step per ~540 cycles
// (6.7e10 path-steps/s, below the SIMT kernel).  Hence: the whole tile-step as one asm block under one elect, and `nmma` issue warps
// on different schedulers (warp index mod 4; blockDim.x = TILES x 128 + 32 nmma), warp j serving the tiles t = j (mod nmma).
template <int NP, int WG, int PPT, int STAGES, int ROUNDS>
__global__ void __launch_bounds__(WG* PTC_ROWS + 32 * PTC_MAX_MMA_WARPS, 1) path_kernel_tc(const __grid_constant__ PtcArgs<NP> a) {
    // WG warpgroups of 128 generator threads; a thread carries PPT paths: the same TMEM lane of the PPT consecutive row tiles its
    // warpgroup owns.  PPT = 2 doubles the independent instruction streams inside a warp (two Philox / Box-Muller chains interleave),
    // which is what this issue-latency-bound generator responds to; the price is registers (two V vectors, two A stages in flight).
    using Cfg = PtcCfg<NP, STAGES>;
    constexpr int TILES = WG * PPT;
    static_assert(TILES <= Cfg::MAX_TILES, "tiles x columns exceed tensor memory");
    constexpr int GEN_WARPS = 4 * WG;
    constexpr int NBINS = 1 << MCP_SEL_BITS;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sHi = smem;
    unsigned char* sLo = smem + Cfg::HI_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sLo + Cfg::LO_BYTES);
    uint64_t* full = bars;                               // [TILES][2] 128 arrivals: the stage's A operand is in tensor memory
    uint64_t* done = bars + 2 * PTC_MAX_TILES;           // [TILES][2] tcgen05.commit: the stage's MMAs are complete
    uint64_t* table_bar = bars + 4 * PTC_MAX_TILES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * PTC_MAX_TILES + 1);
    unsigned int* sHist = reinterpret_cast<unsigned int*>(bars + 4 * PTC_MAX_TILES + 2);   // [NBINS] when a.hist0

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = (int)blockDim.x, nmma = (nthreads - WG * PTC_ROWS) >> 5;
    constexpr uint32_t TMEM_COLS = TILES * Cfg::TILE_COLS <= 128 ? 128u : TILES * Cfg::TILE_COLS <= 256 ? 256u : 512u;
    if (tid == 0) {
        for (int t = 0; t < 2 * TILES; ++t) { mbar_init(&full[t], PTC_ROWS); mbar_init(&done[t], 1); }
        mbar_init(table_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = Cfg::HI_BYTES + Cfg::LO_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(table_bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem)), "l"(a.table), "r"(bytes), "r"(smem_u32(table_bar)) : "memory");
    }
    if (warp == GEN_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (a.hist0 != nullptr)
        for (int i = tid; i < NBINS; i += nthreads) sHist[i] = 0u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(table_bar, 0u);
    const uint32_t tmem = *tmem_slot;

    const uint64_t n_tiles = (a.M + PTC_ROWS - 1) / PTC_ROWS;
    const uint64_t slots = (uint64_t)gridDim.x * TILES;
    const uint32_t S = (uint32_t)a.n_steps;

    if (warp < GEN_WARPS) {
        // ================= generators: thread = PPT paths.  Philox -> Box-Muller -> A stage; R row back -> compounding =================
        const int t0 = (warp >> 2) * PPT, row = tid & (PTC_ROWS - 1);
        const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const float kPi = 3.14159265358979323846f;
        const uint32_t one_bits = opaque_u32(0x3f800000u), two_bits = opaque_u32(0x40000000u);
#pragma unroll
        for (int k = 0; k < PPT; ++k) {   // the constant block [1 1 1 0 0 0 0 0] multiplies the three drift rows of B; written once
            const uint32_t ones[8] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(lane_base + Cfg::TILE_COLS * (uint32_t)(t0 + k) + Cfg::COL_ONE, ones);
        }
        uint32_t ph = 0;                                  // steps this slot has published: step g uses stage g % STAGES for the (g / STAGES)-th time
        // the PPT tiles of a warpgroup advance together: tile index of path k = base + k (out-of-range tiles compute and drop)
        for (uint64_t tile = (uint64_t)blockIdx.x * TILES + t0; tile < n_tiles; tile += slots) {
            uint64_t m[PPT];
            uint32_t c0[PPT], c1[PPT];
            float2 V[PPT][NP / 2];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                m[k] = (tile + (uint64_t)k) * PTC_ROWS + (uint64_t)row;
                const uint64_t gidx = a.first + m[k];
                c0[k] = (uint32_t)gidx;
                c1[k] = (uint32_t)(gidx >> 32);
#pragma unroll
                for (int i = 0; i < NP / 2; ++i) V[k][i] = make_float2(1.f, 1.f);
            }
            auto compound = [&](int k, uint32_t g) {
                // R row of published step g of path k: wait for its MMAs, read the accumulator, V *= 1 + r (asset pairs per FFMA2;
                // the order in which the steps' factors are applied does not matter)
                const uint32_t st = g % STAGES, use = g / STAGES;
                mbar_wait(&done[2 * (t0 + k) + st], use & 1u);
                tc_fence_after();
                uint32_t d[NP];
                const uint32_t addr = lane_base + Cfg::TILE_COLS * (uint32_t)(t0 + k) + st * Cfg::STAGE_COLS + Cfg::COL_D;
                if constexpr (NP == 16) tmem_ld16(addr, d);
                else tmem_ld32(addr, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < NP / 2; ++i)
                    V[k][i] = fma2(V[k][i], make_float2(__uint_as_float(d[2 * i]), __uint_as_float(d[2 * i + 1])), V[k][i]);
            };
            for (uint32_t s = 0; s < S; ++s) {
                // ---- normals of step s: 24-bit Philox fields, Box-Muller on pair j = (field 2j: radius, field 2j+1: angle) ----
                uint32_t z[PPT][NP], zlo[PPT][NP];
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    uint32_t f[NP];
                    philox_fields<NP, ROUNDS>(c0[k], c1[k], s, STREAM_NORMALS, a.rk, f);
#pragma unroll
                    for (int q = 0; q < NP / 4; ++q) {           // normal pairs j = 2q, 2q + 1 in the two halves of a packed operation
                        const float2 f1 = make_float2(__uint_as_float(mant_or(f[4 * q], one_bits)), __uint_as_float(mant_or(f[4 * q + 2], one_bits)));
                        const float2 u1 = fma2(f1, bcast2(-1.0f), bcast2(2.0f));                          // (0, 1]
                        // radius / sqrt(2 ln 2) = sqrt(-lg2 U1): the constant rides in B (host), the negation in the MUFU operand
                        const float2 r = make_float2(Math<float>::sqrt(-Math<float>::lg2(u1.x)), Math<float>::sqrt(-Math<float>::lg2(u1.y)));
                        const float2 f2 = make_float2(__uint_as_float(mant_or(f[4 * q + 1], two_bits)), __uint_as_float(mant_or(f[4 * q + 3], two_bits)));
                        const float2 th = fma2(f2, bcast2(kPi), bcast2(-3.0f * kPi));                     // pi (2f - 1) in [-pi, pi)
                        const float2 cs = make_float2(Math<float>::cosf_(th.x), Math<float>::cosf_(th.y));
                        const float2 sn = make_float2(Math<float>::sinf_(th.x), Math<float>::sinf_(th.y));
                        const float2 zc = fma2(r, cs, bcast2(0.0f));                                      // z[4q], z[4q + 2]
                        const float2 zs = fma2(r, sn, bcast2(0.0f));                                      // z[4q + 1], z[4q + 3]
                        // zlo = z - trunc_tf32(z), exact: what the tensor core drops when it reads z as a TF32 operand
                        const float2 hc = make_float2(__uint_as_float(__float_as_uint(zc.x) & 0xffffe000u), __uint_as_float(__float_as_uint(zc.y) & 0xffffe000u));
                        const float2 hs = make_float2(__uint_as_float(__float_as_uint(zs.x) & 0xffffe000u), __uint_as_float(__float_as_uint(zs.y) & 0xffffe000u));
                        const float2 lc = fma2(hc, bcast2(-1.0f), zc), ls = fma2(hs, bcast2(-1.0f), zs);
                        z[k][4 * q] = __float_as_uint(zc.x); z[k][4 * q + 1] = __float_as_uint(zc.y);
                        z[k][4 * q + 2] = __float_as_uint(zs.x); z[k][4 * q + 3] = __float_as_uint(zs.y);
                        zlo[k][4 * q] = __float_as_uint(lc.x); zlo[k][4 * q + 1] = __float_as_uint(lc.y);
                        zlo[k][4 * q + 2] = __float_as_uint(ls.x); zlo[k][4 * q + 3] = __float_as_uint(ls.y);
                    }
                }
                // ---- the stage this step goes into was last used STAGES steps ago: compound that step (its MMAs are long done),
                //      which also frees the stage's A columns and accumulator ----
                if (s >= (uint32_t)STAGES) {
#pragma unroll
                    for (int k = 0; k < PPT; ++k) compound(k, ph - STAGES);
                }
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const uint32_t stage = lane_base + Cfg::TILE_COLS * (uint32_t)(t0 + k) + (ph % STAGES) * Cfg::STAGE_COLS;
                    if constexpr (NP == 16) {
                        tmem_st16(stage + Cfg::COL_Z, z[k]);
                        tmem_st16(stage + Cfg::COL_ZLO, zlo[k]);
                    } else {
                        tmem_st32(stage + Cfg::COL_Z, z[k]);
                        tmem_st32(stage + Cfg::COL_ZLO, zlo[k]);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
#pragma unroll
                for (int k = 0; k < PPT; ++k) mbar_arrive(&full[2 * (t0 + k) + (ph % STAGES)]);
                ++ph;
            }
            // drain: the last min(S, STAGES) steps are still to be compounded (oldest first)
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                if (STAGES == 2 && S >= 2u) compound(k, ph - 2u);
                compound(k, ph - 1u);
                float2 x2 = make_float2(-1.0f, 0.0f);
#pragma unroll
                for (int i = 0; i < NP / 2; ++i) x2 = fma2(make_float2(a.w[2 * i], a.w[2 * i + 1]), V[k][i], x2);
                const float x = x2.x + x2.y;
                if (m[k] < a.M) {
                    a.terminal[m[k]] = x;
                    if (a.hist0 != nullptr) {                 // first radix digit of the order-preserving key (mcp_quantile.cu, pass 0)
                        const unsigned digit = f32_to_key(__float_as_uint(x)) >> (32 - MCP_SEL_BITS);
                        const unsigned act = __activemask();
                        const unsigned peers = __match_any_sync(act, digit);
                        if (lane == __ffs(peers) - 1) atomicAdd(&sHist[digit], (unsigned)__popc(peers));
                    }
                }
            }
        }
    } else {
        // ================= MMA issue warp j: one elected thread, warp-convergent; its tiles t = j (mod nmma), step by step =================
        // (a warpgroup's PPT tiles advance together, so every tile of the CTA runs the same number of rounds: tiles past the end of
        //  the range are computed and dropped by their generators)
        const int mj = warp - GEN_WARPS;
        const uint32_t idesc = tc_idesc(2u, (uint32_t)NP);           // TF32 operands, FP32 accumulate, M = 128, N = NP
        const uint64_t bhi0 = tc_sdesc(smem_u32(sHi), Cfg::LBO, Cfg::SBO), blo0 = tc_sdesc(smem_u32(sLo), Cfg::LBO, Cfg::SBO);
        const uint64_t first0 = (uint64_t)blockIdx.x * TILES;
        uint32_t ph = 0;
        for (uint64_t it = 0; first0 + it * slots < n_tiles; ++it) {
            for (uint32_t s = 0; s < S; ++s, ++ph) {
                const uint32_t st = ph % STAGES, use = ph / STAGES;
                for (int t = mj; t < TILES; t += nmma) {
                    if (first0 + (uint64_t)(t / PPT * PPT) + it * slots >= n_tiles) continue;        // this warpgroup has run out of tiles
                    const uint32_t base = tmem + Cfg::TILE_COLS * (uint32_t)t, sb = base + st * Cfg::STAGE_COLS;
                    mbar_wait_idle(&full[2 * t + st], use & 1u);
                    tc_fence_after();
                    ptc_issue<NP>(sb + Cfg::COL_D, sb + Cfg::COL_Z, sb + Cfg::COL_ZLO, base + Cfg::COL_ONE, bhi0, blo0, idesc, smem_u32(&done[2 * t + st]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == GEN_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    if (a.hist0 != nullptr) {
        for (int i = tid; i < NBINS; i += nthreads) {
            const unsigned c = sHist[i];
            if (c) atomicAdd(&a.hist0[i], (unsigned long long)c);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// 32 < N <= 128: the same scheme with the step's normals produced and stored 16 at a time.  A thread still owns one path (its
// compounded values V are N registers), but a whole step's normals and their remainders would not fit next to them, so chunk c
// (normals 16c .. 16c+15 = Philox blocks 3c .. 3c+2 of the step's stream, as in every other path kernel) goes from the registers
// straight into its 16 + 16 tensor-memory columns before the next chunk is drawn.  One A / D stage per tile (3 N + 8 columns:
// two tiles at N = 64, one at N = 128); the first chunk of step s + 1 is drawn before the thread waits for the MMAs of step s.
// The MMA warp issues the 3 N / 8 + 1 tcgen05.mma of a tile-step in a loop (49 at N = 128): at these widths the generator's
// work per step, which grows with N like the MMA count, keeps the issue off the critical path.
// ---------------------------------------------------------------------------------------------------------------------------
template <int NP, int WG, int ROUNDS>
__global__ void __launch_bounds__(WG* PTC_ROWS + 32 * WG, 1) path_kernel_tc_wide(const __grid_constant__ PtcArgs<NP> a) {
    using Cfg = PtcCfg<NP, 1>;
    constexpr int TILES = WG, GEN_WARPS = 4 * WG, NCH = NP / 16;
    constexpr int NBINS = 1 << MCP_SEL_BITS;
    static_assert(TILES * Cfg::TILE_COLS <= 512, "tiles x columns exceed tensor memory");
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr uint32_t IMG_BYTES = Cfg::HI_BYTES + Cfg::LO_BYTES;       // one K round: Bhi image | Blo image
    const int KR = a.k_rounds;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)KR * IMG_BYTES);
    uint64_t* full = bars;                               // [TILES] 128 arrivals: the round's A operand is in tensor memory
    uint64_t* done = bars + PTC_MAX_TILES;               // [TILES] tcgen05.commit: the round's MMAs are complete
    uint64_t* table_bar = bars + 2 * PTC_MAX_TILES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PTC_MAX_TILES + 1);
    unsigned int* sHist = reinterpret_cast<unsigned int*>(bars + 2 * PTC_MAX_TILES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = (int)blockDim.x;
    if (tid == 0) {
        for (int t = 0; t < TILES; ++t) { mbar_init(&full[t], PTC_ROWS); mbar_init(&done[t], 1); }
        mbar_init(table_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)KR * IMG_BYTES;
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

    const uint64_t n_tiles = (a.M + PTC_ROWS - 1) / PTC_ROWS;
    const uint64_t slots = (uint64_t)gridDim.x * TILES;
    const uint32_t S = (uint32_t)a.n_steps;

    if (warp < GEN_WARPS) {
        const int t0 = warp >> 2, row = tid & (PTC_ROWS - 1);
        const uint32_t tile_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + Cfg::TILE_COLS * (uint32_t)t0;
        const float kPi = 3.14159265358979323846f;
        const uint32_t one_bits = opaque_u32(0x3f800000u), two_bits = opaque_u32(0x40000000u);
        {
            const uint32_t ones[8] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(tile_base + Cfg::COL_ONE, ones);
        }
        uint32_t ph = 0;
        for (uint64_t tile = (uint64_t)blockIdx.x * TILES + t0; tile < n_tiles; tile += slots) {
            const uint64_t m = tile * PTC_ROWS + (uint64_t)row;
            const uint64_t gidx = a.first + m;
            const uint32_t c0 = (uint32_t)gidx, c1 = (uint32_t)(gidx >> 32);
            float2 V[NP / 2];
#pragma unroll
            for (int i = 0; i < NP / 2; ++i) V[i] = make_float2(1.f, 1.f);
            // normals 16c .. 16c+15 of step s and their TF32 remainders (register order z[4q], z[4q+2], z[4q+1], z[4q+3], see ptc_normal_of_col)
            auto draw = [&](uint32_t s, int c, uint32_t (&z)[16], uint32_t (&zlo)[16]) {
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
                    const float2 hc = make_float2(__uint_as_float(__float_as_uint(zc.x) & 0xffffe000u), __uint_as_float(__float_as_uint(zc.y) & 0xffffe000u));
                    const float2 hs = make_float2(__uint_as_float(__float_as_uint(zs.x) & 0xffffe000u), __uint_as_float(__float_as_uint(zs.y) & 0xffffe000u));
                    const float2 lc = fma2(hc, bcast2(-1.0f), zc), ls = fma2(hs, bcast2(-1.0f), zs);
                    z[4 * q] = __float_as_uint(zc.x); z[4 * q + 1] = __float_as_uint(zc.y);
                    z[4 * q + 2] = __float_as_uint(zs.x); z[4 * q + 3] = __float_as_uint(zs.y);
                    zlo[4 * q] = __float_as_uint(lc.x); zlo[4 * q + 1] = __float_as_uint(lc.y);
                    zlo[4 * q + 2] = __float_as_uint(ls.x); zlo[4 * q + 3] = __float_as_uint(ls.y);
                }
            };
            auto compound = [&]() {                      // R row of the finished step: V *= 1 + r
#pragma unroll
                for (int b = 0; b < NP / 32; ++b) {
                    uint32_t d[32];
                    tmem_ld32(tile_base + Cfg::COL_D + 32u * (uint32_t)b, d);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        V[16 * b + i] = fma2(V[16 * b + i], make_float2(__uint_as_float(d[2 * i]), __uint_as_float(d[2 * i + 1])), V[16 * b + i]);
                }
            };
            for (uint32_t s = 0; s < S; ++s) {
                for (int r = 0; r < KR; ++r) {           // K rounds: normals r NP .. r NP + NP - 1 against round r's images of B
                    uint32_t z[16], zlo[16];
                    draw(s, r * NCH, z, zlo);            // overlaps the MMAs of the previous round
                    if (s > 0 || r > 0) {                // ... which must be complete before the tile's one A stage is overwritten
                        mbar_wait(&done[t0], (ph - 1u) & 1u);
                        tc_fence_after();
                        if (r == 0) compound();          // a new step: the finished one's R row is read first (frees the accumulator)
                    }
                    tmem_st16(tile_base + Cfg::COL_Z, z);
                    tmem_st16(tile_base + Cfg::COL_ZLO, zlo);
#pragma unroll
                    for (int c = 1; c < NCH; ++c) {
                        draw(s, r * NCH + c, z, zlo);
                        tmem_st16(tile_base + Cfg::COL_Z + 16u * (uint32_t)c, z);
                        tmem_st16(tile_base + Cfg::COL_ZLO + 16u * (uint32_t)c, zlo);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    mbar_arrive(&full[t0]);
                    ++ph;
                }
            }
            mbar_wait(&done[t0], (ph - 1u) & 1u);
            tc_fence_after();
            compound();
            float2 x2 = make_float2(a.x0, 0.0f);
#pragma unroll
            for (int i = 0; i < NP / 2; ++i) x2 = fma2(make_float2(a.w[2 * i], a.w[2 * i + 1]), V[i], x2);
            const float x = x2.x + x2.y;
            if (m < a.M) {
                a.terminal[m] = x;
                if (a.hist0 != nullptr) {
                    const unsigned digit = f32_to_key(__float_as_uint(x)) >> (32 - MCP_SEL_BITS);
                    const unsigned act = __activemask();
                    const unsigned peers = __match_any_sync(act, digit);
                    if (lane == __ffs(peers) - 1) atomicAdd(&sHist[digit], (unsigned)__popc(peers));
                }
            }
        }
    } else {
        // MMA issue warp of tile t = warp - GEN_WARPS: R = z Lhi (overwrites) + zlo Lhi + z Llo + [1 1 1 0 ..] drift, K = 8 per tcgen05.mma
        const int t = warp - GEN_WARPS;
        const uint32_t idesc = tc_idesc(2u, (uint32_t)NP);
        const uint32_t base = tmem + Cfg::TILE_COLS * (uint32_t)t;
        uint32_t ph = 0;
        for (uint64_t tile = (uint64_t)blockIdx.x * TILES + (uint64_t)t; tile < n_tiles; tile += slots) {
            for (uint32_t s = 0; s < S; ++s) {
                for (int r = 0; r < KR; ++r, ++ph) {
                    const uint32_t img = smem_u32(smem) + (uint32_t)r * IMG_BYTES;
                    const uint64_t bhi0 = tc_sdesc(img, Cfg::LBO, Cfg::SBO), blo0 = tc_sdesc(img + Cfg::HI_BYTES, Cfg::LBO, Cfg::SBO);
                    mbar_wait_idle(&full[t], ph & 1u);
                    tc_fence_after();
#pragma unroll 1
                    for (uint32_t k = 0; k < (uint32_t)(NP / 8); ++k)      // round 0 overwrites the accumulator, later rounds add to it
                        mma_tf32_ts_elect(base + Cfg::COL_D, base + Cfg::COL_Z + 8u * k, bhi0 + (uint64_t)(k * Cfg::DESC_STEP), idesc, (k > 0 || r > 0) ? 1u : 0u);
#pragma unroll 1
                    for (uint32_t k = 0; k < (uint32_t)(NP / 8); ++k)
                        mma_tf32_ts_elect(base + Cfg::COL_D, base + Cfg::COL_ZLO + 8u * k, bhi0 + (uint64_t)(k * Cfg::DESC_STEP), idesc, 1u);
#pragma unroll 1
                    for (uint32_t k = 0; k < (uint32_t)(NP / 8); ++k)
                        mma_tf32_ts_elect(base + Cfg::COL_D, base + Cfg::COL_Z + 8u * k, blo0 + (uint64_t)(k * Cfg::DESC_STEP), idesc, 1u);
                    if (r == 0)                                            // the drift rows ride in round 0's image
                        mma_tf32_ts_elect(base + Cfg::COL_D, base + Cfg::COL_ONE, bhi0 + (uint64_t)((NP / 8) * Cfg::DESC_STEP), idesc, 1u);
                    tc_commit_elect(&done[t]);
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

static inline float ptc_tf32_round(float x) {          // round-to-nearest-even onto 10 explicit mantissa bits
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x00000fffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

bool path_tc_eligible(const PathJob& job) {
    const char* v = getenv("MCP_PATHS_TC");               // "0" forces the SIMT kernels (A/B tests, benchmarks)
    if (v && v[0] == '0') return false;
    return job.dtype == MCP_F32 && job.z_in == nullptr && job.n <= 256;
}

static int ptc_env(const char* name, int lo, int hi) {   // tuning knobs: 0 = default
    const char* v = getenv(name);
    const int t = v ? atoi(v) : 0;
    return t >= lo && t <= hi ? t : 0;
}

template <int NP, int WG, int PPT, int STAGES, int ROUNDS>
static int ptc_launch_t(mcp_context* h, PathJob& job, const PtcArgs<NP>& a, int nmma) {
    using Cfg = PtcCfg<NP, STAGES>;
    constexpr int TILES = WG * PPT;
    auto kern = path_kernel_tc<NP, WG, PPT, STAGES, ROUNDS>;
    const size_t smem = Cfg::HI_BYTES + Cfg::LO_BYTES + (4 * PTC_MAX_TILES + 2) * sizeof(uint64_t) + (job.hist0 ? sizeof(unsigned int) << MCP_SEL_BITS : 0) + 128;
    MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t n_tiles = (job.M + PTC_ROWS - 1) / PTC_ROWS;
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount, (n_tiles + TILES - 1) / TILES);
    if (grid < 1) grid = 1;
    nmma = std::max(1, std::min(nmma, std::min(TILES, PTC_MAX_MMA_WARPS)));
    kern<<<(unsigned)grid, WG * PTC_ROWS + 32 * nmma, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <int NP, int WG, int ROUNDS>
static int ptc_launch_wide_t(mcp_context* h, PathJob& job, const PtcArgs<NP>& a) {
    using Cfg = PtcCfg<NP, 1>;
    auto kern = path_kernel_tc_wide<NP, WG, ROUNDS>;
    const size_t smem = (size_t)a.k_rounds * (Cfg::HI_BYTES + Cfg::LO_BYTES) + (2 * PTC_MAX_TILES + 2) * sizeof(uint64_t) +
                        (a.hist0 ? sizeof(unsigned int) << MCP_SEL_BITS : 0) + 128;
    if (smem > h->prop.sharedMemPerBlockOptin)
        return mcp_fail(h, MCP_ERR_INVALID, "path_kernel_tc_wide: %d K rounds need %zu B of shared memory (max %zu)", a.k_rounds, smem, (size_t)h->prop.sharedMemPerBlockOptin);
    MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t n_tiles = (job.M + PTC_ROWS - 1) / PTC_ROWS;
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount, (n_tiles + WG - 1) / WG);
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, WG * PTC_ROWS + 32 * WG, smem, job.stream>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <int NP>
static int ptc_launch(mcp_context* h, PathJob& job) {
    using Cfg = PtcCfg<NP, 1>;
    const int n = job.n;
    const std::vector<double>& L = *job.L;
    const double sdt = std::sqrt(job.dt), c = std::sqrt(2.0 * std::log(2.0));     // the kernel's normals come out divided by sqrt(2 ln 2)
    // B[i][kc] (output asset i, A column kc), canonical K-major no-swizzle core-matrix layout: [K core of 4][N group of 8][8 rows x 16 B]
    std::vector<unsigned char> host(Cfg::HI_BYTES + Cfg::LO_BYTES, 0);
    auto at = [&](size_t image, int i, int kc) -> float* {
        return reinterpret_cast<float*>(host.data() + image + ((size_t)(kc / 4) * (NP / 8) + i / 8) * 128 + (i % 8) * 16 + (kc % 4) * 4);
    };
    for (int i = 0; i < n; ++i) {
        for (int kc = 0; kc < NP; ++kc) {
            const int j = ptc_normal_of_col(kc);
            if (j > i || j >= n) continue;                                       // L is lower triangular
            const double v = L[(size_t)i * n + j] * sdt * c;
            const float hi = ptc_tf32_round((float)v);
            *at(0, i, kc) = hi;
            *at(Cfg::HI_BYTES, i, kc) = ptc_tf32_round((float)(v - (double)hi));
        }
        // drift mu_i dt as three TF32 pieces against the constant A block [1 1 1 0 ...]
        const double dr = job.mu[i] * job.dt;
        const float d0 = ptc_tf32_round((float)dr), d1 = ptc_tf32_round((float)(dr - (double)d0));
        const float d2 = ptc_tf32_round((float)(dr - (double)d0 - (double)d1));
        *at(0, i, NP) = d0;
        *at(0, i, NP + 1) = d1;
        *at(0, i, NP + 2) = d2;
    }
    unsigned char* dev = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 6, host.size(), (void**)&dev));
    ++h->const_epoch;
    MCP_CUDA(h, cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, job.stream));
    MCP_CUDA(h, cudaStreamSynchronize(job.stream));        // `host` is pageable and dies at scope exit

    PtcArgs<NP> a;
    a.table = dev;
    for (int i = 0; i < NP; ++i) a.w[i] = i < n ? (float)job.w[i] : 0.f;
    a.terminal = (float*)job.terminal;
    a.hist0 = job.hist0;
    a.first = job.first;
    a.M = job.M;
    a.n_steps = job.n_steps;
    a.k_rounds = 1;
    a.x0 = -1.0f;
    philox_keys_fill(a.rk, job.seed);
    job.hist0_filled = job.hist0 != nullptr;
    static const int wg_env = ptc_env("MCP_PATHS_TC_WG", 1, PTC_MAX_TILES), ppt_env = ptc_env("MCP_PATHS_TC_PPT", 1, 2),
                     stages_env = ptc_env("MCP_PATHS_TC_STAGES", 1, 2), nmma_env = ptc_env("MCP_PATHS_TC_MMA", 1, PTC_MAX_MMA_WARPS);
    const int nmma = nmma_env ? nmma_env : 4;
    if constexpr (NP > 32) {
        // one stage per tile, 3 NP + 8 columns: two tiles at NP = 64, one at NP = 128
        constexpr int WGW = NP == 64 ? 2 : 1;
        (void)wg_env; (void)ppt_env; (void)stages_env; (void)nmma;
        return job.rounds == 7 ? ptc_launch_wide_t<NP, WGW, 7>(h, job, a) : ptc_launch_wide_t<NP, WGW, 10>(h, job, a);
    } else {
#define MCP_PTC(W, P, ST)                                                                                  \
    return job.rounds == 7 ? ptc_launch_t<NP, W, P, ST, 7>(h, job, a, nmma) : ptc_launch_t<NP, W, P, ST, 10>(h, job, a, nmma);
    if constexpr (NP == 16) {
        // measured on B200 (C4: 1e7 paths x 252 steps; SIMT kernel 7.25e10 path-steps/s): WG 4 / PPT 1 / 1 stage / 4 issue warps 7.82e10;
        // 2 stages 7.53e10; PPT 2 (WG 3) 7.27e10; 7 tiles 7.55e10; one issue warp 6.75e10 -- the generator (Philox + Box-Muller: ALU and
        // XU pipes) bounds every variant, the contraction is off its critical path
        const int ppt = ppt_env ? ppt_env : 1, stages = stages_env ? stages_env : 1;
        if (ppt == 2) {
            if (stages == 2) {                   // 104 columns per tile: 4 tiles
                switch (wg_env) {
                    case 1: MCP_PTC(1, 2, 2)
                    default: MCP_PTC(2, 2, 2)
                }
            }
            switch (wg_env) {                    // 56 columns per tile: up to 7 tiles
                case 1: MCP_PTC(1, 2, 1)
                case 2: MCP_PTC(2, 2, 1)
                default: MCP_PTC(3, 2, 1)
            }
        }
        if (stages == 2) {
            switch (wg_env) {
                case 2: MCP_PTC(2, 1, 2)
                case 3: MCP_PTC(3, 1, 2)
                default: MCP_PTC(4, 1, 2)
            }
        }
        switch (wg_env) {
            case 2: MCP_PTC(2, 1, 1)
            case 3: MCP_PTC(3, 1, 1)
            case 5: MCP_PTC(5, 1, 1)
            case 6: MCP_PTC(6, 1, 1)
            case 7: MCP_PTC(7, 1, 1)
            default: MCP_PTC(4, 1, 1)
        }
    } else {
        switch (wg_env) {
            case 2: MCP_PTC(2, 1, 1)
            case 4: MCP_PTC(4, 1, 1)
            default: MCP_PTC(3, 1, 1)
        }
    }
#undef MCP_PTC
    }
}

// terminal[m] = sum_b partial[b][m] - 1, blocks added in ascending order (deterministic)
__global__ void __launch_bounds__(256) ptc_combine_blocks(const float* __restrict__ partial, int n_blocks, uint64_t M, float* __restrict__ terminal) {
    for (uint64_t m = (uint64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (uint64_t)gridDim.x * 256) {
        float x = -1.0f;
        for (int b = 0; b < n_blocks; ++b) x += partial[(size_t)b * M + m];
        terminal[m] = x;
    }
}

// 128 < N <= 256: the universe is cut into blocks of 64 assets.  L is lower triangular, so block b (assets 64 b .. 64 b + 63) only
// reaches the normals 0 .. 64 (b + 1) - 1: one launch of the N = 64 kernel per block, with b + 1 K rounds of 64 normals each
// accumulating in the tile's tensor-memory accumulator (b + 1 image pairs of B in shared memory: 139 KB for the last block of a
// 256-asset universe).  A launch compounds its 64 assets and writes the block's part of w . V_T; a small kernel adds the parts.
// The normals of the earlier blocks are drawn again by the later ones (the Philox counter makes them the same numbers): 2.5x the
// draws of a single pass at N = 256 -- still an order of magnitude above the SIMT warp-per-path kernel, whose N^2 FMAs per step it
// moves to the tensor cores.  (A thread cannot own more assets: V is one register per asset.)
static int ptc_launch_blocks(mcp_context* h, PathJob& job) {
    using Cfg = PtcCfg<64, 1>;
    constexpr size_t IMG = Cfg::HI_BYTES + Cfg::LO_BYTES;
    const int n = job.n, B = (n + 63) / 64;
    const std::vector<double>& L = *job.L;
    const double sdt = std::sqrt(job.dt), c = std::sqrt(2.0 * std::log(2.0));
    std::vector<size_t> off(B + 1, 0);
    for (int b = 0; b < B; ++b) off[b + 1] = off[b] + (size_t)(b + 1) * IMG;
    std::vector<unsigned char> host(off[B], 0);
    auto at = [&](size_t image, int i, int kc) -> float* {
        return reinterpret_cast<float*>(host.data() + image + ((size_t)(kc / 4) * (64 / 8) + i / 8) * 128 + (i % 8) * 16 + (kc % 4) * 4);
    };
    for (int b = 0; b < B; ++b) {
        for (int r = 0; r <= b; ++r) {
            const size_t img = off[b] + (size_t)r * IMG;
            for (int i = 0; i < 64; ++i) {
                const int gi = 64 * b + i;
                if (gi >= n) continue;
                for (int kc = 0; kc < 64; ++kc) {
                    const int j = 64 * r + ptc_normal_of_col(kc);
                    if (j > gi || j >= n) continue;
                    const double v = L[(size_t)gi * n + j] * sdt * c;
                    const float hi = ptc_tf32_round((float)v);
                    *at(img, i, kc) = hi;
                    *at(img + Cfg::HI_BYTES, i, kc) = ptc_tf32_round((float)(v - (double)hi));
                }
                if (r == 0) {
                    const double dr = job.mu[gi] * job.dt;
                    const float d0 = ptc_tf32_round((float)dr), d1 = ptc_tf32_round((float)(dr - (double)d0));
                    const float d2 = ptc_tf32_round((float)(dr - (double)d0 - (double)d1));
                    *at(img, i, 64) = d0;
                    *at(img, i, 65) = d1;
                    *at(img, i, 66) = d2;
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
    MCP_CHECK(mcp_dev_reserve(h, 23, (size_t)B * job.M * sizeof(float), (void**)&partial));
    for (int b = 0; b < B; ++b) {
        PtcArgs<64> a;
        a.table = dev + off[b];
        for (int i = 0; i < 64; ++i) a.w[i] = 64 * b + i < n ? (float)job.w[64 * b + i] : 0.f;
        a.terminal = partial + (size_t)b * job.M;
        a.hist0 = nullptr;
        a.first = job.first;
        a.M = job.M;
        a.n_steps = job.n_steps;
        a.k_rounds = b + 1;
        a.x0 = 0.0f;
        philox_keys_fill(a.rk, job.seed);
        MCP_CHECK(job.rounds == 7 ? (ptc_launch_wide_t<64, 2, 7>(h, job, a)) : (ptc_launch_wide_t<64, 2, 10>(h, job, a)));
    }
    const unsigned blocks = (unsigned)std::min<uint64_t>((job.M + 255) / 256, (uint64_t)h->prop.multiProcessorCount * 8);
    ptc_combine_blocks<<<blocks, 256, 0, job.stream>>>(partial, B, job.M, (float*)job.terminal);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    job.hist0_filled = false;              // the terminal values exist only after the combination: the select starts at pass 0
    return MCP_OK;
}

int path_launch_tc(mcp_context* h, PathJob& job) {
    if (job.n <= 16) return ptc_launch<16>(h, job);
    if (job.n <= 32) return ptc_launch<32>(h, job);
    if (path_tc16_enabled()) return path_launch_tc16(h, job);
    if (job.n <= 64) return ptc_launch<64>(h, job);
    return job.n <= 128 ? ptc_launch<128>(h, job) : ptc_launch_blocks(h, job);
}

}  // namespace mcp
