// mcp_historical_var: per-portfolio HISTORICAL VaR / CVaR and the 'VaR' / 'CVaR' selections.
//
// Reference: app.py:710 `port_series = returns_df @ ws`, 712-713 `var(port_series, .95)` /
// `cvar(port_series, .95)` (definitions 258-263), 717 metric = -var / -cvar, 673-674 + 747
// `np.argmin(metric)`  ==  first index of the LARGEST var (cvar).  This is where ~70 % of the
// reference's loop time goes.
//
// One warp per portfolio.  The returns matrix R (T x N) sits in shared memory, transposed and
// padded so that lanes reading consecutive periods hit consecutive banks; the portfolio's
// weights are a shared-memory broadcast.  Each lane owns periods lane, lane+32, ... of the
// series in registers.  The two order statistics np.percentile needs are found exactly on
// order-preserving integer keys: for lower-tail ranks (k < 40, the reference's 5 % tail) by
// sorting each lane's values and popping the warp-wide minimum k+1 times (shuffle min + ballot),
// otherwise by an MSB-first bitwise radix select (32 or 64 REDUX rounds).  The tail mean is a
// warp reduction; the selections reuse the (key, first index) argmax machinery of the sweep.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mcp_device.cuh"
#include "mcp_portfolio.h"

namespace mcp {

constexpr int HV_BLOCK = 256;
constexpr int HV_WARPS = HV_BLOCK / 32;
constexpr int HV_EXTRACT_MAX = 40;        // ranks below this are found by repeated warp-min extraction
constexpr int HV_SHRINK = 68;             // set sizes 0 .. 64 (two positions of 32 lanes)

template <typename K> __device__ __forceinline__ K shfl_xor_key(K v, int m);
template <> __device__ __forceinline__ uint32_t shfl_xor_key<uint32_t>(uint32_t v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ uint64_t shfl_xor_key<uint64_t>(uint64_t v, int m) { return __shfl_xor_sync(0xffffffffu, (unsigned long long)v, m); }

template <typename T> struct HKey;
template <> struct HKey<float> {
    using K = uint32_t;
    static constexpr int BITS = 32;
    static __device__ __forceinline__ K to_key(float v) { return f32_to_key(__float_as_uint(v)); }
    static __device__ __forceinline__ float from_key(K k) { return __uint_as_float(key_to_f32(k)); }
};
template <> struct HKey<double> {
    using K = uint64_t;
    static constexpr int BITS = 64;
    static __device__ __forceinline__ K to_key(double v) { return f64_to_key((uint64_t)__double_as_longlong(v)); }
    static __device__ __forceinline__ double from_key(K k) { return __longlong_as_double((long long)key_to_f64(k)); }
};

template <typename T>
struct HistArgs {
    const T* w_in;          // [P, n]
    const T* r_t;           // [n][t_pad] transposed returns (global; staged into smem)
    T* var_out;
    T* cvar_out;
    PfCand* cands;
    uint64_t first, P;
    int n, T_, t_pad;
    int k_lo, k_hi;         // 0-based order statistics
    T gamma;                // interpolation weight
    T out_sign;             // +1, or -1 when the arrays hold the app's metric -var / -cvar (app.py:717); picks use the unsigned values
    float shrink[HV_SHRINK]; // hist_var_fast: where between its two thresholds the refined one goes, by the size of the first set
};

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor<T>(v, m);
    return v;
}

template <typename T, int VPL>
__global__ void __launch_bounds__(HV_BLOCK) hist_var_kernel(const HistArgs<T> a) {
    using K = typename HKey<T>::K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sR = reinterpret_cast<T*>(smem_raw);                 // [n][t_pad]
    T* sW = sR + (size_t)a.n * a.t_pad;                     // [HV_WARPS][n]
    for (int i = threadIdx.x; i < a.n * a.t_pad; i += HV_BLOCK) sR[i] = a.r_t[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* myW = sW + warp * a.n;

    T best_v = -Math<T>::inf(), best_c = -Math<T>::inf();
    uint64_t idx_v = MCP_NO_INDEX, idx_c = MCP_NO_INDEX;
    const uint64_t warps_total = (uint64_t)gridDim.x * HV_WARPS;
    for (uint64_t p = (uint64_t)blockIdx.x * HV_WARPS + warp; p < a.P; p += warps_total) {
        __syncwarp();
        for (int i = lane; i < a.n; i += 32) myW[i] = a.w_in[p * a.n + i];
        __syncwarp();
        // ---- series[t] = sum_i R[t, i] w_i for this lane's periods ----
        T x[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) x[v] = (T)0;
        for (int i = 0; i < a.n; ++i) {
            const T wi = myW[i];
            const T* row = sR + (size_t)i * a.t_pad + lane;
#pragma unroll
            for (int v = 0; v < VPL; ++v) x[v] = Math<T>::fma(row[32 * v], wi, x[v]);   // padded columns are 0
        }
        K key[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const bool valid = lane + 32 * v < a.T_;
            key[v] = valid ? HKey<T>::to_key(x[v]) : ~(K)0;      // padding sorts last
        }
        T var, cvar;
        if (a.k_hi < HV_EXTRACT_MAX) {
            // ---- lower-tail ranks (the reference's alpha = 0.95: k = 18 of T = 365): sort each lane's
            // values once, then pop the warp-wide minimum k_lo + 1 times (5 shuffles + a register shift each).
            // The tail mean comes out of the same pops: the set {x <= VaR} is exactly what has been popped when the
            // next minimum exceeds VaR (the loop below pops on while it does not: ties at the rank, gamma = 0).  Each
            // lane adds ITS popped values in ascending order, the lanes' sums meet in the butterfly -- hist_var_fast
            // reaches the same set by another route and adds in the same order (bit-identical results).
            K y[VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) y[v] = key[v];
#pragma unroll
            for (int round = 0; round < VPL; ++round) {          // odd-even transposition sort (ascending)
#pragma unroll
                for (int v = round & 1; v + 1 < VPL; v += 2) {
                    const K lo = y[v] < y[v + 1] ? y[v] : y[v + 1], hi = y[v] < y[v + 1] ? y[v + 1] : y[v];
                    y[v] = lo; y[v + 1] = hi;
                }
            }
            T acc = (T)0;
            auto warp_min_key = [&]() {
                K m = y[0];
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) { const K o = shfl_xor_key<K>(m, d); m = o < m ? o : m; }
                return m;
            };
            auto pop = [&](K m) {                                // the lowest lane holding m gives it up
                const unsigned owners = __ballot_sync(0xffffffffu, y[0] == m);
                if (lane == __ffs(owners) - 1) {
                    acc += HKey<T>::from_key(m);
#pragma unroll
                    for (int v = 0; v + 1 < VPL; ++v) y[v] = y[v + 1];
                    y[VPL - 1] = ~(K)0;
                }
            };
            K k_at_lo = 0;
#pragma unroll 1
            for (int r = 0; r <= a.k_lo; ++r) { k_at_lo = warp_min_key(); pop(k_at_lo); }
            K peek = warp_min_key();
            const T v_lo = HKey<T>::from_key(k_at_lo), v_hi = a.k_hi > a.k_lo ? HKey<T>::from_key(peek) : v_lo;
            const T diff = v_hi - v_lo;                          // numpy _lerp
            var = v_lo + diff * a.gamma;
            if (a.gamma >= (T)0.5) var = v_hi - diff * ((T)1 - a.gamma);
            int total = a.k_lo + 1;
#pragma unroll 1
            while (total < a.T_ && HKey<T>::from_key(peek) <= var) { pop(peek); ++total; peek = warp_min_key(); }
            cvar = warp_sum<T>(acc) / (T)total;
        } else {
            // ---- any rank: MSB-first radix select, count keys with the candidate prefix and bit 0 ----
            K prefix = 0;
            int rank = a.k_lo;
#pragma unroll 1
            for (int b = HKey<T>::BITS - 1; b >= 0; --b) {
                const K himask = b == HKey<T>::BITS - 1 ? (K)0 : (K)(~(K)0 << (b + 1));
                int c = 0;
#pragma unroll
                for (int v = 0; v < VPL; ++v) c += ((key[v] & himask) == prefix && !((key[v] >> b) & 1)) ? 1 : 0;
                c = __reduce_add_sync(0xffffffffu, c);
                if (rank >= c) { rank -= c; prefix |= (K)1 << b; }
            }
            const T v_lo = HKey<T>::from_key(prefix);
            // (k_lo+1)-th: v_lo again if it is repeated far enough, else the smallest value above it
            int le = 0;
            T above = Math<T>::inf();
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const bool valid = lane + 32 * v < a.T_;
                le += (valid && key[v] <= prefix) ? 1 : 0;
                if (valid && key[v] > prefix && x[v] < above) above = x[v];
            }
            le = __reduce_add_sync(0xffffffffu, le);
            above = warp_min<T>(above);
            const T v_hi = (a.k_hi == a.k_lo || le >= a.k_hi + 1) ? v_lo : above;
            const T diff = v_hi - v_lo;                          // numpy _lerp
            var = v_lo + diff * a.gamma;
            if (a.gamma >= (T)0.5) var = v_hi - diff * ((T)1 - a.gamma);
            // ---- CVaR = mean(x[x <= VaR]) (VaR if empty), app.py:261-263 ----
            T s = (T)0;
            int cnt = 0;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const bool in = (lane + 32 * v < a.T_) && x[v] <= var;
                s += in ? x[v] : (T)0;
                cnt += in ? 1 : 0;
            }
            s = warp_sum<T>(s);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            cvar = cnt > 0 ? s / (T)cnt : var;
        }
        if (lane == 0) {
            if (a.var_out) a.var_out[p] = a.out_sign * var;
            if (a.cvar_out) a.cvar_out[p] = a.out_sign * cvar;
        }
        const uint64_t g = a.first + p;
        if (var > best_v) { best_v = var; idx_v = g; }          // p ascends per warp: first occurrence kept
        if (cvar > best_c) { best_c = cvar; idx_c = g; }
    }
    __shared__ PfCand wc[HV_WARPS];
    if (lane == 0) wc[warp] = PfCand{(double)best_v, idx_v, (double)best_c, idx_c, 0.0, 0.0};
    __syncthreads();
    if (threadIdx.x == 0) {
        PfCand b = wc[0];
        for (int w = 1; w < HV_WARPS; ++w) {
            const PfCand o = wc[w];
            if (cand_better<double>(o.key_s, o.idx_s, b.key_s, b.idx_s)) { b.key_s = o.key_s; b.idx_s = o.idx_s; }
            if (cand_better<double>(o.key_d, o.idx_d, b.key_d, b.idx_d)) { b.key_d = o.key_d; b.idx_d = o.idx_d; }
        }
        a.cands[blockIdx.x] = b;
    }
}

// ---------------------------------------------------------------------------------------------
// FP32 fast path for lower-tail ranks (the reference's alpha = 0.95 on T <= 512 periods): a warp works on
// FOUR consecutive portfolios at once.
//   * series: lane l owns the periods l, l + 32, ... as in the plain kernel.  R sits in shared memory as PAIRS of a lane's
//     consecutive periods (one LDS.64 = two periods), the weights as (w, w) pairs (one broadcast LDS.128 = two portfolios),
//     so a packed FFMA2 advances two periods of one portfolio: per asset 6 + 2 loads feed 24 FFMA2 (VPL = 12), no
//     register moves, no zeroing (the first asset is an FMUL2); the FMA order per (period, portfolio) is the plain kernel's.
//   * selection, on the FP32 values themselves (FMNMX, CREDUX.MIN/MAX.F32 -- no integer keys): each lane sorts its VPL
//     values with a sorting network and parks them in shared memory.  With tau = the warp-wide minimum of the lanes'
//     (ROW+1)-th smallest values, everything <= tau in the lanes' first ROW positions is a set S of the |S| smallest values
//     of the series (nothing outside S is below tau), found with ONE reduction; for ROW = 2 and 32 lanes |S| is 21 +- 7
//     (measured), around the 19 values the reference's alpha = 0.95 needs at T = 365.  Any threshold below tau is as
//     valid, so a second one -- interpolated between the minimum of the lanes' ROW-th values and tau -- is counted too and
//     the set nearer to the rank is kept (2-3 values off on average).  From there single steps finish: warp-min pops from
//     below when the set is too small, warp-max removals when it is too big (one CREDUX + ballot each; the owning lane moves
//     to its neighbouring parked value).  ROW is chosen from k_lo on the host (0: pops only).  The next portfolio's sort
//     shares a basic block with this one's reductions (two parking areas).
//   * tail mean: the set {x <= VaR} is what has been taken when the next minimum exceeds VaR (else the pops go on: ties);
//     each lane adds ITS taken values in ascending order and the lanes' sums meet in the butterfly, as in the plain kernel.
// VaR / CVaR are bit-identical to hist_var_kernel<float, VPL> up to the sign of a zero (tests compare the two).
// Measured (B200, T = 365, N = 16, 4e6 portfolios): 1.7e9 pf/s at alpha = 0.95 (round 1: 1.0e9 with k + 1 pops on integer
// keys), 1.8e9 at alpha = 0.99; the plain kernel 3.8e8 / 7.6e8.
// ---------------------------------------------------------------------------------------------
constexpr int HF_PPW = 4;
constexpr int HF_PAD = 2;                 // sentinels on either side of a lane's parked values

__device__ __forceinline__ float redux_min_f32(float v) {
    float m;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
__device__ __forceinline__ float redux_max_f32(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}

template <int VPL> __device__ __forceinline__ void lane_sort(float (&y)[VPL]) {
#define MCP_CE(i, j) { const float lo_ = fminf(y[i], y[j]), hi_ = fmaxf(y[i], y[j]); y[i] = lo_; y[j] = hi_; }
    if constexpr (VPL == 12) {          // 39-comparator network (verified with the 0/1 principle)
        MCP_CE(0, 1) MCP_CE(2, 3) MCP_CE(4, 5) MCP_CE(6, 7) MCP_CE(8, 9) MCP_CE(10, 11)
        MCP_CE(1, 3) MCP_CE(5, 7) MCP_CE(9, 11) MCP_CE(0, 2) MCP_CE(4, 6) MCP_CE(8, 10)
        MCP_CE(1, 2) MCP_CE(5, 6) MCP_CE(9, 10) MCP_CE(0, 4) MCP_CE(7, 11)
        MCP_CE(1, 5) MCP_CE(6, 10) MCP_CE(3, 7) MCP_CE(4, 8)
        MCP_CE(5, 9) MCP_CE(2, 6) MCP_CE(0, 4) MCP_CE(7, 11) MCP_CE(3, 8)
        MCP_CE(1, 5) MCP_CE(6, 10) MCP_CE(2, 3) MCP_CE(8, 9)
        MCP_CE(1, 4) MCP_CE(7, 10) MCP_CE(3, 5) MCP_CE(6, 8)
        MCP_CE(2, 4) MCP_CE(7, 9) MCP_CE(5, 6)
        MCP_CE(3, 4) MCP_CE(7, 8)
    } else {                            // odd-even transposition
#pragma unroll
        for (int round = 0; round < VPL; ++round) {
#pragma unroll
            for (int v = round & 1; v + 1 < VPL; v += 2) MCP_CE(v, v + 1)
        }
    }
#undef MCP_CE
}

template <int VPL, int ROW, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) hist_var_fast(const HistArgs<float> a) {
    static_assert(VPL % 2 == 0 && ROW < VPL, "period pairs");
    constexpr int VP2 = VPL / 2, KSTRIDE = VPL + 2 * HF_PAD, WARPS = BLOCK / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sR = reinterpret_cast<float*>(smem_raw);                                 // [n][VP2][32][2], t_pad = 32 VPL
    float* sW = sR + (size_t)a.n * a.t_pad;                                         // [WARPS][n][4][2]
    float* sK = sW + (size_t)WARPS * a.n * HF_PPW * 2;                              // [WARPS][2][KSTRIDE][32]
    for (int d = threadIdx.x; d < a.n * a.t_pad; d += BLOCK) {
        const int i = d / a.t_pad, rem = d - i * a.t_pad, j = rem >> 6, l = (rem & 63) >> 1, half = rem & 1;
        sR[d] = a.r_t[i * a.t_pad + l + 64 * j + 32 * half];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* myW = sW + (size_t)warp * a.n * HF_PPW * 2;
    float* myK = sK + (size_t)warp * 2 * KSTRIDE * 32 + lane;                       // entry e of this lane: myK[e * 32]; two areas
    unsigned lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    const float inf = Math<float>::inf();
#pragma unroll
    for (int e = 0; e < HF_PAD; ++e)
#pragma unroll
        for (int ar = 0; ar < 2; ++ar) { myK[(ar * KSTRIDE + e) * 32] = -inf; myK[(ar * KSTRIDE + HF_PAD + VPL + e) * 32] = inf; }
    __syncthreads();

    float best_v = -inf, best_c = -inf;                     // lane pp < 4 follows portfolio pp of every group
    uint32_t it_v = 0xffffffffu, it_c = 0xffffffffu, it = 0;   // the warp's sweep number of the best group (a 64-bit index costs two more registers each)
    const uint64_t groups = (a.P + HF_PPW - 1) / HF_PPW;
    const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
    const bool last_only = a.T_ > 32 * (VPL - 1);           // only a lane's last period can be padding
    const bool last_valid = lane + 32 * (VPL - 1) < a.T_;
    // the four weight rows of a group are 4 n contiguous floats; with 4 n <= 64 (n <= 16) a lane holds the NEXT group's
    // values in two registers while the current group is being processed (the global-load latency sat in front of the STS)
    const bool prefetch = HF_PPW * a.n <= 64;
    const int j0 = lane, j1 = lane + 32, gsz = HF_PPW * a.n;
    const int o0 = ((j0 % a.n) * HF_PPW + j0 / a.n) * 2, o1 = ((j1 % a.n) * HF_PPW + j1 / a.n) * 2;
    const uint64_t g_first = (uint64_t)blockIdx.x * WARPS + warp;
    const uint64_t w_total = a.P * (uint64_t)a.n, w_step = warps_total * (uint64_t)gsz;      // elements of w_in, per sweep of the grid
    uint64_t e0 = g_first * (uint64_t)gsz + (uint64_t)j0;                                      // this lane's element of the NEXT group
    float nxt0 = 0.f, nxt1 = 0.f;
    if (prefetch) {
        nxt0 = (j0 < gsz && e0 < w_total) ? __ldg(a.w_in + e0) : 0.f;
        nxt1 = (j1 < gsz && e0 + 32 < w_total) ? __ldg(a.w_in + e0 + 32) : 0.f;
    }
    for (uint64_t gq = g_first; gq < groups; gq += warps_total, ++it) {
        const uint64_t p0 = gq * HF_PPW;
        __syncwarp();
        if (prefetch) {
            if (j0 < gsz) *reinterpret_cast<float2*>(myW + o0) = make_float2(nxt0, nxt0);
            if (j1 < gsz) *reinterpret_cast<float2*>(myW + o1) = make_float2(nxt1, nxt1);
            e0 += w_step;                                   // rows past P read as 0 (a partial last group, the tail of the grid)
            nxt0 = (j0 < gsz && e0 < w_total) ? __ldg(a.w_in + e0) : 0.f;
            nxt1 = (j1 < gsz && e0 + 32 < w_total) ? __ldg(a.w_in + e0 + 32) : 0.f;
        } else {
            for (int j = lane; j < gsz; j += 32) {                        // rows p0 .. p0+3 are contiguous: coalesced
                const int pp = j / a.n, i = j - pp * a.n;
                const float w = (p0 + (uint64_t)pp < a.P) ? a.w_in[p0 * (uint64_t)a.n + (uint64_t)j] : 0.f;
                *reinterpret_cast<float2*>(myW + (i * HF_PPW + pp) * 2) = make_float2(w, w);
            }
        }
        __syncwarp();
        // ---- series[t] = sum_i R[t, i] w_i: this lane's periods in pairs (l + 64 j, l + 64 j + 32), four portfolios ----
        float2 x[HF_PPW][VP2];
        {
            const float2* row = reinterpret_cast<const float2*>(sR) + lane;
            const float4* wq = reinterpret_cast<const float4*>(myW);
            const int row_step = a.t_pad >> 1;
            {   // asset 0: fma(r, w, 0) = r w
                const float4 wa = wq[0], wb = wq[1];                    // (w0, w0, w1, w1), (w2, w2, w3, w3)
                const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w), w2 = make_float2(wb.x, wb.y), w3 = make_float2(wb.z, wb.w);
#pragma unroll
                for (int j = 0; j < VP2; ++j) {
                    const float2 r2 = row[32 * j];                      // padded periods are 0
                    x[0][j] = mul2(r2, w0);
                    x[1][j] = mul2(r2, w1);
                    x[2][j] = mul2(r2, w2);
                    x[3][j] = mul2(r2, w3);
                }
            }
#pragma unroll 1
            for (int i = 1; i < a.n; ++i) {                            // (unrolled, ptxas hoists the next assets' loads and spills)
                row += row_step;
                wq += 2;
                const float4 wa = wq[0], wb = wq[1];
                const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w), w2 = make_float2(wb.x, wb.y), w3 = make_float2(wb.z, wb.w);
#pragma unroll
                for (int j = 0; j < VP2; ++j) {
                    const float2 r2 = row[32 * j];
                    x[0][j] = fma2(r2, w0, x[0][j]);
                    x[1][j] = fma2(r2, w1, x[1][j]);
                    x[2][j] = fma2(r2, w2, x[2][j]);
                    x[3][j] = fma2(r2, w3, x[3][j]);
                }
            }
        }
        float var4[HF_PPW], acc4[HF_PPW];
        int total4[HF_PPW];
        // padded periods become +inf (they sort last): once per group, so that the per-portfolio code below has no branch in it
        if (last_only) {
#pragma unroll
            for (int pp = 0; pp < HF_PPW; ++pp) x[pp][VP2 - 1].y = last_valid ? x[pp][VP2 - 1].y : inf;
        } else {
#pragma unroll
            for (int j = 0; j < VP2; ++j) {
                const bool v0 = lane + 64 * j < a.T_, v1 = lane + 64 * j + 32 < a.T_;
#pragma unroll
                for (int pp = 0; pp < HF_PPW; ++pp) { x[pp][j].x = v0 ? x[pp][j].x : inf; x[pp][j].y = v1 ? x[pp][j].y : inf; }
            }
        }
        // ---- sort a portfolio's values in this lane and park them between the sentinels ----
        float ys[HF_PPW][VPL];
        auto prepare = [&](int pp) {                        // pp is a literal at every call (unrolled)
            float (&y)[VPL] = ys[pp];
#pragma unroll
            for (int j = 0; j < VP2; ++j) { y[2 * j] = x[pp][j].x; y[2 * j + 1] = x[pp][j].y; }
            lane_sort<VPL>(y);
            float* park = myK + (pp & 1) * KSTRIDE * 32;   // two areas: the next portfolio is parked while this one's steps run
#pragma unroll
            for (int v = 0; v < VPL; ++v) park[(HF_PAD + v) * 32] = y[v];
        };
        prepare(0);
#pragma unroll
        for (int pp = 0; pp < HF_PPW; ++pp) {
            float (&y)[VPL] = ys[pp];
            float* const parked = myK + (pp & 1) * KSTRIDE * 32;
            // ---- S = everything <= tau in the first ROW positions: the |S| smallest values of the series ----
            int cnt = 0;                                    // how many of this lane's values have been taken
            int c = 0;
            if constexpr (ROW > 0) {
                // The size of S has a wide spread (tau is an extreme value), so a second threshold, interpolated between the minimum
                // of the lanes' ROW-th values and tau by the sizes the two sets typically have, is counted as well: it lands within
                // 2-3 values of the rank.  Any theta <= tau gives a valid set; if S is too big, the nearer of the two sets is kept.
                // Branch-free, and in one block with the NEXT portfolio's sort: the reductions' latencies hide behind its FMNMXs.
                const float tau = redux_min_f32(y[ROW]);
                const float tau_l = redux_min_f32(y[ROW - 1]);
#pragma unroll
                for (int v = 0; v < ROW; ++v) cnt += y[v] <= tau ? 1 : 0;       // a prefix: y ascends
                c = __reduce_add_sync(0xffffffffu, cnt);
                const float theta = fminf(fmaf(tau - tau_l, a.shrink[c], tau_l), tau);
                int cnt2 = 0;
#pragma unroll
                for (int v = 0; v < ROW; ++v) cnt2 += y[v] <= theta ? 1 : 0;
                const int c2 = __reduce_add_sync(0xffffffffu, cnt2);
                const int d2 = c2 > a.k_lo + 1 ? c2 - (a.k_lo + 1) : a.k_lo + 1 - c2;
                const bool use2 = c > a.k_lo + 1 && d2 < c - (a.k_lo + 1);
                cnt = use2 ? cnt2 : cnt;
                c = use2 ? c2 : c;
            }
            if (pp + 1 < HF_PPW) prepare(pp + 1);
            float acc = 0.f, v_lo, h;
            const float* nx;                                // the parked value after h
            if (c <= a.k_lo) {
                // pops from below: the lowest lane holding the minimum gives it up
#pragma unroll
                for (int v = 0; v < ROW; ++v) acc = v < cnt ? acc + y[v] : acc;
                h = y[0];
#pragma unroll
                for (int v = 1; v <= ROW; ++v) h = cnt >= v ? y[v] : h;
                nx = parked + (HF_PAD + 1 + cnt) * 32;
                float m = 0.f;
#pragma unroll 1
                for (int need = a.k_lo + 1 - c; need > 0; --need) {
                    m = redux_min_f32(h);
                    const unsigned owners = __ballot_sync(0xffffffffu, h == m);
                    if (h == m && (owners & lt_mask) == 0u) { acc += h; h = *nx; nx += 32; }
                }
                v_lo = m;
            } else {
                // removals from above: S loses its largest values until k_lo + 1 are left
                const float* pv = parked + (HF_PAD - 1 + cnt) * 32;      // top of this lane's part of S (-inf: none)
                float top = *pv;
#pragma unroll 1
                for (int rem = c - a.k_lo - 1; rem > 0; --rem) {
                    const float M = redux_max_f32(top);
                    const unsigned owners = __ballot_sync(0xffffffffu, top == M);
                    if (top == M && (owners & lt_mask) == 0u) { --cnt; pv -= 32; top = *pv; }
                }
                v_lo = redux_max_f32(top);
#pragma unroll
                for (int v = 0; v < ROW; ++v) acc = v < cnt ? acc + y[v] : acc;
                nx = pv + 32;
                h = *nx;
                nx += 32;
            }
            // ---- numpy _lerp, then CVaR = mean(x[x <= VaR]), app.py:261-263 ----
            float peek = redux_min_f32(h);
            const float v_hi = a.k_hi > a.k_lo ? peek : v_lo;
            const float diff = v_hi - v_lo;
            float var = v_lo + diff * a.gamma;
            if (a.gamma >= 0.5f) var = v_hi - diff * (1.f - a.gamma);
            int total = a.k_lo + 1;
#pragma unroll 1
            while (total < a.T_ && peek <= var) {           // ties at the rank (or gamma = 0 / 1 with equal neighbours)
                const unsigned owners = __ballot_sync(0xffffffffu, h == peek);
                if (h == peek && (owners & lt_mask) == 0u) { acc += h; h = *nx; nx += 32; }
                ++total;
                peek = redux_min_f32(h);
            }
            var4[pp] = var;
            acc4[pp] = acc;
            total4[pp] = total;
        }
#pragma unroll
        for (int pp = 0; pp < HF_PPW; ++pp) acc4[pp] = warp_sum<float>(acc4[pp]);        // four independent butterflies
        // ---- lane pp writes and follows portfolio pp (one division per lane) ----
        const float my_var = lane == 0 ? var4[0] : lane == 1 ? var4[1] : lane == 2 ? var4[2] : var4[3];
        const float my_sum = lane == 0 ? acc4[0] : lane == 1 ? acc4[1] : lane == 2 ? acc4[2] : acc4[3];
        const int my_total = lane == 0 ? total4[0] : lane == 1 ? total4[1] : lane == 2 ? total4[2] : total4[3];
        const float my_cvar = my_sum / (float)my_total;
        const uint64_t p = p0 + (uint64_t)lane;
        if (lane < HF_PPW && p < a.P) {
            if (a.var_out) a.var_out[p] = a.out_sign * my_var;
            if (a.cvar_out) a.cvar_out[p] = a.out_sign * my_cvar;
            if (my_var > best_v) { best_v = my_var; it_v = it; }          // p ascends per lane: first occurrence kept
            if (my_cvar > best_c) { best_c = my_cvar; it_c = it; }
        }
    }
    uint64_t idx_v = it_v == 0xffffffffu ? MCP_NO_INDEX : a.first + (g_first + (uint64_t)it_v * warps_total) * HF_PPW + (uint64_t)lane;
    uint64_t idx_c = it_c == 0xffffffffu ? MCP_NO_INDEX : a.first + (g_first + (uint64_t)it_c * warps_total) * HF_PPW + (uint64_t)lane;
    warp_argmax<float>(best_v, idx_v);                      // larger value, then lower index
    warp_argmax<float>(best_c, idx_c);
    __shared__ PfCand wc[WARPS];
    if (lane == 0) wc[warp] = PfCand{(double)best_v, idx_v, (double)best_c, idx_c, 0.0, 0.0};
    __syncthreads();
    if (threadIdx.x == 0) {
        PfCand b = wc[0];
        for (int w = 1; w < WARPS; ++w) {
            const PfCand o = wc[w];
            if (cand_better<double>(o.key_s, o.idx_s, b.key_s, b.idx_s)) { b.key_s = o.key_s; b.idx_s = o.idx_s; }
            if (cand_better<double>(o.key_d, o.idx_d, b.key_d, b.idx_d)) { b.key_d = o.key_d; b.idx_d = o.idx_d; }
        }
        a.cands[blockIdx.x] = b;
    }
}

constexpr size_t hist_fast_smem(int n, int t_pad, int vpl, int warps) {
    return ((size_t)n * t_pad + (size_t)warps * n * HF_PPW * 2 + (size_t)warps * 2 * (vpl + 2 * HF_PAD) * 32) * sizeof(float);
}

template <int VPL, int ROW, int BLOCK, int MINB>
static int hist_launch_fast_occ(mcp_context* h, const HistArgs<float>& a, int max_blocks, int* blocks, cudaStream_t st, bool* done) {
    constexpr int WARPS = BLOCK / 32;
    const size_t smem = hist_fast_smem(a.n, a.t_pad, VPL, WARPS);
    if (smem > h->prop.sharedMemPerBlockOptin) return MCP_OK;                                   // the plain kernel takes it
    auto kern = hist_var_fast<VPL, ROW, BLOCK, MINB>;
    if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MCP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem));
    if (per_sm < 1) return MCP_OK;
    const uint64_t groups = (a.P + HF_PPW - 1) / HF_PPW;
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount * per_sm, (groups + WARPS - 1) / WARPS);
    grid = std::max<uint64_t>(1, std::min<uint64_t>(grid, max_blocks));
    *blocks = (int)grid;
    kern<<<(unsigned)grid, BLOCK, smem, st>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    *done = true;
    return MCP_OK;
}

template <int VPL, int ROW>
static int hist_launch_fast_row(mcp_context* h, const HistArgs<float>& a, int max_blocks, int* blocks, cudaStream_t st, bool* done) {
    const char* env = getenv("MCP_HIST_OCC");                     // A/B tests: block size x resident blocks the kernel is compiled for
    if (env && env[0] == '2') return hist_launch_fast_occ<VPL, ROW, 256, 2>(h, a, max_blocks, blocks, st, done);     // 16 warps, <= 128 registers
    if (env && env[0] == '3') return hist_launch_fast_occ<VPL, ROW, 256, 3>(h, a, max_blocks, blocks, st, done);     // 24 warps, 80 registers (spills)
    return hist_launch_fast_occ<VPL, ROW, 320, 2>(h, a, max_blocks, blocks, st, done);                               // 20 warps, 96 registers
}

template <int VPL>
static int hist_launch_fast(mcp_context* h, const HistArgs<float>& a, int max_blocks, int* blocks, cudaStream_t st, bool* done) {
    *done = false;
    if (a.t_pad != 32 * VPL) return MCP_OK;                                                     // the plain kernel takes it
    // first positions of the 32 lanes holding (in expectation) 7 / 16 of the smallest values: start next to k_lo + 1
    int row = a.k_lo + 1 >= 12 ? 2 : a.k_lo + 1 >= 5 ? 1 : 0;
    if (const char* env = getenv("MCP_HIST_ROW")) row = std::max(0, std::min(2, atoi(env)));     // A/B tests
    // refined threshold = tau_l + (tau - tau_l) shrink[|S|]: linear in the rank between the sizes the sets below tau_l (typically
    // 1 / 7 values for row 1 / 2) and below tau (|S|) have; a heuristic only -- every threshold <= tau is exact
    HistArgs<float> b = a;
    const float need = (float)(a.k_lo + 1), c_l = row == 2 ? 7.f : 1.f;
    for (int c = 0; c < HV_SHRINK; ++c) b.shrink[c] = (float)c > c_l ? std::min(1.f, std::max(0.f, (need - c_l) / ((float)c - c_l))) : 1.f;
    if (getenv("MCP_HIST_REFINE") && getenv("MCP_HIST_REFINE")[0] == '0') for (float& f : b.shrink) f = 1.f;      // A/B tests: theta = tau
    if (row == 2) return hist_launch_fast_row<VPL, 2>(h, b, max_blocks, blocks, st, done);
    if (row == 1) return hist_launch_fast_row<VPL, 1>(h, b, max_blocks, blocks, st, done);
    return hist_launch_fast_row<VPL, 0>(h, b, max_blocks, blocks, st, done);
}

template <typename T, int VPL>
static int hist_launch_t(mcp_context* h, const HistArgs<T>& a, size_t smem, int max_blocks, int* blocks, cudaStream_t st) {
    auto kern = hist_var_kernel<T, VPL>;
    if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MCP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, HV_BLOCK, smem));
    if (per_sm < 1) return mcp_fail(h, MCP_ERR_INVALID, "mcp_historical_var: returns matrix does not fit in shared memory (%zu B)", smem);
    uint64_t grid = std::min<uint64_t>((uint64_t)h->prop.multiProcessorCount * per_sm, (a.P + HV_WARPS - 1) / HV_WARPS);
    grid = std::max<uint64_t>(1, std::min<uint64_t>(grid, max_blocks));
    *blocks = (int)grid;
    kern<<<(unsigned)grid, HV_BLOCK, smem, st>>>(a);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

template <typename T>
static int hist_dispatch(mcp_context* h, const HistArgs<T>& a, size_t smem, int max_blocks, int* blocks, cudaStream_t st) {
    const int vpl = (a.T_ + 31) / 32;
    if constexpr (sizeof(T) == 4) {
        const char* env = getenv("MCP_HIST_FAST");                 // "0" forces the plain kernel (A/B tests)
        if (a.k_hi < HV_EXTRACT_MAX && !(env && env[0] == '0')) {
            bool done = false;
            if (vpl <= 4 && a.t_pad == 128) MCP_CHECK(hist_launch_fast<4>(h, a, max_blocks, blocks, st, &done));
            else if (vpl <= 8 && a.t_pad == 256) MCP_CHECK(hist_launch_fast<8>(h, a, max_blocks, blocks, st, &done));
            else if (vpl <= 12 && a.t_pad == 384) MCP_CHECK(hist_launch_fast<12>(h, a, max_blocks, blocks, st, &done));
            else if (vpl <= 16 && a.t_pad == 512) MCP_CHECK(hist_launch_fast<16>(h, a, max_blocks, blocks, st, &done));
            if (done) return MCP_OK;
        }
    }
#define MCP_HV(V) if (vpl <= V) return hist_launch_t<T, V>(h, a, smem, max_blocks, blocks, st);
    MCP_HV(1) MCP_HV(2) MCP_HV(4) MCP_HV(8) MCP_HV(12) MCP_HV(16) MCP_HV(24) MCP_HV(32) MCP_HV(64)
#undef MCP_HV
    return mcp_fail(h, MCP_ERR_INVALID, "mcp_historical_var: n_periods=%d exceeds the supported maximum of 2048", a.T_);
}

// ---- FP32 near-tie recheck of the 'VaR' / 'CVaR' picks (mcp_hist_params.recheck) --------------------------------------------
// np.argmin(-var) on the reference's FP64 values (app.py:673-674, 747) can land on another row than the FP32 argmax when two
// portfolios' VaR differ by less than FP32 rounding.  The FP32 pass screens: every row whose value is within a rounding
// tolerance of the best is recorded, those few rows are re-evaluated in FP64 (series, order statistics, tail mean: the plain
// kernel in double) and the pick is the FP64 maximum with the lowest index.
__global__ void __launch_bounds__(256) hv_collect(const float* __restrict__ var, const float* __restrict__ cvar, uint64_t n, uint64_t base,
                                                  const PfCand* __restrict__ fin, float sign, float tol, RcLists* lists) {
    const PfCand f = *fin;
    if (f.idx_s == MCP_NO_INDEX) return;
    const float thr_v = (float)f.key_s - tol, thr_c = (float)f.key_d - tol;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const float v = sign * var[i], c = sign * cvar[i];
        if (v >= thr_v) {
            const unsigned slot = atomicAdd(&lists->count[0], 1u);
            if (slot < RC_CAP) { lists->idx[0][slot] = base + i; lists->key[0][slot] = v; }
        }
        if (c >= thr_c) {
            const unsigned slot = atomicAdd(&lists->count[1], 1u);
            if (slot < RC_CAP) { lists->idx[1][slot] = base + i; lists->key[1][slot] = c; }
        }
    }
}

// rows[k][:] = (double) w[idx[k] - base][:]
__global__ void __launch_bounds__(256) hv_gather_rows(const float* __restrict__ w, const unsigned long long* __restrict__ idx, uint64_t base,
                                                      int count, int n, double* __restrict__ rows) {
    for (int e = blockIdx.x * 256 + threadIdx.x; e < count * n; e += gridDim.x * 256) {
        const int k = e / n, i = e - k * n;
        rows[e] = (double)w[(idx[k] - base) * (uint64_t)n + (uint64_t)i];
    }
}

// transposed, padded copy of R in the kernel's arithmetic type at `dst` (device); returns the padded period count
template <typename T>
static int hist_upload_r(mcp_context* h, const double* R, int n, int Tn, T* dst, int tp, cudaStream_t st) {
    std::vector<T> rt((size_t)n * tp, (T)0);
    for (int t = 0; t < Tn; ++t)
        for (int i = 0; i < n; ++i) rt[(size_t)i * tp + t] = (T)R[(size_t)t * n + i];
    MCP_CUDA(h, cudaMemcpyAsync(dst, rt.data(), rt.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));               // `rt` is pageable and dies at scope exit
    return MCP_OK;
}

static int hist_bucket(int Tn) {
    const int vpl = (Tn + 31) / 32;
    for (int v : {1, 2, 4, 8, 12, 16, 24, 32, 64}) if (vpl <= v) return v;
    return 64;
}

// one pass of the historical kernel over device rows; leaves ev[0] / ev[1] around the kernel and the merged candidate in *fin
template <typename T>
static int hist_exec(mcp_context* h, const T* d_w, T* d_var, T* d_cvar, uint64_t P, uint64_t first, int n, int Tn, int tp, double alpha,
                     T out_sign, const T* d_rt, PfCand* cands, PfCand* fin, int max_blocks, cudaStream_t st) {
    // order statistics of np.percentile(x, (1-alpha)*100), 'linear': h = (T-1) q  (app.py:259)
    const double percent = (1 - alpha) * 100, q = percent / 100.0, hidx = (double)(Tn - 1) * q;
    int k_lo, k_hi;
    if (hidx >= (double)(Tn - 1)) k_lo = k_hi = Tn - 1;
    else if (hidx < 0) k_lo = k_hi = 0;
    else { k_lo = (int)std::floor(hidx); k_hi = k_lo + 1; }
    HistArgs<T> a;
    a.w_in = d_w; a.r_t = d_rt; a.var_out = d_var; a.cvar_out = d_cvar; a.cands = cands;
    a.first = first; a.P = P; a.n = n; a.T_ = Tn; a.t_pad = tp;
    a.k_lo = k_lo; a.k_hi = k_hi; a.gamma = (T)(hidx - std::floor(hidx)); a.out_sign = out_sign;
    const size_t smem = ((size_t)n * tp + (size_t)HV_WARPS * n) * sizeof(T);
    MCP_REQUIRE(h, smem <= h->prop.sharedMemPerBlockOptin, "mcp_historical_var: T=%d x N=%d needs %zu B of shared memory (max %zu)",
                Tn, n, smem, (size_t)h->prop.sharedMemPerBlockOptin);
    int blocks = 0;
    MCP_CUDA(h, cudaEventRecord(h->ev[0], st));
    MCP_CHECK(hist_dispatch<T>(h, a, smem, max_blocks, &blocks, st));
    MCP_CUDA(h, cudaEventRecord(h->ev[1], st));
    return pf_reduce_launch(h, cands, blocks, fin, 0, st);
}

template <typename T>
static int hist_run(mcp_context* h, const mcp_hist_params* p, const double* R, mcp_hist_out* out) {
    const int n = p->n_assets, Tn = p->n_periods;
    const uint64_t P = p->n_portfolios;
    cudaStream_t st = h->stream;
    MCP_REQUIRE(h, (Tn + 31) / 32 <= 64, "mcp_historical_var: n_periods=%d exceeds the supported maximum of 2048", Tn);
    const int tp = 32 * hist_bucket(Tn);        // padded lanes read columns up to 32 * VPL_template: pad to the dispatch bucket
    const bool recheck = sizeof(T) == 4 && p->recheck != 0;

    const int max_blocks = h->prop.multiProcessorCount * 16;
    unsigned char* base = nullptr;
    const size_t off_rt = sizeof(PfCand) * (size_t)(max_blocks + 1);
    MCP_CHECK(mcp_dev_reserve(h, 0, off_rt + (size_t)n * tp * sizeof(T) + 256, (void**)&base));
    PfCand* cands = (PfCand*)base;
    PfCand* fin = cands + max_blocks;
    T* d_rt = (T*)(base + (off_rt + 255) / 256 * 256);
    MCP_CHECK(hist_upload_r<T>(h, R, n, Tn, d_rt, tp, st));

    const T* d_w = (const T*)p->weights_in;
    T* d_var = (T*)out->var;
    T* d_cvar = (T*)out->cvar;
    const bool stage = p->space == MCP_HOST;
    if (stage) {
        void* q = nullptr;
        MCP_CHECK(mcp_dev_reserve(h, 1, P * n * sizeof(T), &q));
        MCP_CUDA(h, cudaMemcpyAsync(q, p->weights_in, P * n * sizeof(T), cudaMemcpyHostToDevice, st));
        d_w = (const T*)q;
    }
    if (stage || (recheck && !(out->var && out->cvar))) {
        // device copies of the arrays: HOST space always, DEVICE space when the recheck needs values the caller did not ask for
        void* q = nullptr;
        MCP_CHECK(mcp_dev_reserve(h, 3, 2 * P * sizeof(T), &q));
        if (stage) {
            d_var = (out->var || recheck) ? (T*)q : nullptr;
            d_cvar = (out->cvar || recheck) ? (T*)q + P : nullptr;
        } else {
            if (!d_var) d_var = (T*)q;
            if (!d_cvar) d_cvar = (T*)q + P;
        }
    }
    const T sign = p->negate ? (T)-1 : (T)1;
    MCP_CHECK(hist_exec<T>(h, d_w, d_var, d_cvar, P, p->first_index, n, Tn, tp, p->alpha, sign, d_rt, cands, fin, max_blocks, st));
    if (stage) {
        if (out->var) MCP_CUDA(h, cudaMemcpyAsync(out->var, d_var, P * sizeof(T), cudaMemcpyDeviceToHost, st));
        if (out->cvar) MCP_CUDA(h, cudaMemcpyAsync(out->cvar, d_cvar, P * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
    PfCand f;
    MCP_CUDA(h, cudaMemcpyAsync(&f, fin, sizeof f, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));
    float ms = 0;
    MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    out->best_var_index = f.idx_s;
    out->best_var = f.key_s;
    out->best_cvar_index = f.idx_d;
    out->best_cvar = f.key_d;
    out->kernel_ms = ms;
    h->last_ms = ms;

    if constexpr (sizeof(T) == 4) {
        if (recheck && f.idx_s != MCP_NO_INDEX) {
            double rmax = 0;
            for (size_t i = 0; i < (size_t)Tn * n; ++i) rmax = std::max(rmax, std::fabs(R[i]));
            // FP32 error of a series value: eps * sum_i |R_ti w_i| <= eps * max|R| * sum|w|; 64 ulps of headroom (sum|w| = 1 for
            // Dirichlet weights, and the tail mean adds less than one ulp per term)
            const float tol = (float)(5.96e-8 * 64 * (rmax + std::max(std::fabs(f.key_s), std::fabs(f.key_d))));
            unsigned char* rb = nullptr;
            const size_t rows_cap = (size_t)RC_CAP * n * sizeof(double);
            const size_t off_rows = (sizeof(RcLists) + 255) / 256 * 256, off_vals = off_rows + (rows_cap + 255) / 256 * 256;
            const size_t off_rt64 = off_vals + (size_t)2 * RC_CAP * sizeof(double) + 256;
            const size_t off_c64 = off_rt64 + ((size_t)n * tp * sizeof(double) + 255) / 256 * 256;
            MCP_CHECK(mcp_dev_reserve(h, 19, off_c64 + sizeof(PfCand) * (size_t)(max_blocks + 1) + 256, (void**)&rb));
            RcLists* d_lists = (RcLists*)rb;
            double* d_rows = (double*)(rb + off_rows);
            double* d_vals = (double*)(rb + off_vals);
            double* d_rt64 = (double*)(rb + off_rt64);
            PfCand* c64 = (PfCand*)(rb + off_c64);
            MCP_CUDA(h, cudaMemsetAsync(d_lists, 0, 16, st));
            const uint64_t g = std::max<uint64_t>(1, std::min<uint64_t>((P + 255) / 256, (uint64_t)h->prop.multiProcessorCount * 8));
            hv_collect<<<(unsigned)g, 256, 0, st>>>((const float*)d_var, (const float*)d_cvar, P, p->first_index, fin, (float)sign, tol, d_lists);
            MCP_CUDA(h, cudaGetLastError());
            h->launches++;
            std::vector<unsigned char> raw(sizeof(RcLists));
            MCP_CUDA(h, cudaMemcpyAsync(raw.data(), d_lists, sizeof(RcLists), cudaMemcpyDeviceToHost, st));
            MCP_CUDA(h, cudaStreamSynchronize(st));
            const RcLists* Lh = reinterpret_cast<const RcLists*>(raw.data());
            bool r_up = false;
            for (int c = 0; c < 2; ++c) {
                if (Lh->count[c] <= 1) continue;                         // no near-tie: the FP32 pick stands
                MCP_REQUIRE(h, Lh->count[c] <= RC_CAP, "mcp_historical_var: more than %u portfolios tie with the best %s within FP32 rounding; "
                            "run with dtype = MCP_F64 for an exact pick", RC_CAP, c == 0 ? "VaR" : "CVaR");
                std::vector<unsigned long long> idx(Lh->idx[c], Lh->idx[c] + Lh->count[c]);
                std::sort(idx.begin(), idx.end());
                const int cnt = (int)idx.size();
                // the sorted indices go back into the list slot (device) for the gather
                MCP_CUDA(h, cudaMemcpyAsync(d_lists->idx[c], idx.data(), sizeof(unsigned long long) * cnt, cudaMemcpyHostToDevice, st));
                hv_gather_rows<<<std::max(1, std::min(1024, (cnt * n + 255) / 256)), 256, 0, st>>>((const float*)d_w, d_lists->idx[c], p->first_index, cnt, n, d_rows);
                MCP_CUDA(h, cudaGetLastError());
                h->launches++;
                if (!r_up) { MCP_CHECK(hist_upload_r<double>(h, R, n, Tn, d_rt64, tp, st)); r_up = true; }
                MCP_CHECK(hist_exec<double>(h, d_rows, d_vals, d_vals + RC_CAP, (uint64_t)cnt, 0, n, Tn, tp, p->alpha, 1.0, d_rt64, c64, c64 + max_blocks, max_blocks, st));
                std::vector<double> vals((size_t)cnt);
                MCP_CUDA(h, cudaMemcpyAsync(vals.data(), c == 0 ? d_vals : d_vals + RC_CAP, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
                MCP_CUDA(h, cudaStreamSynchronize(st));
                int best = 0;
                for (int k = 1; k < cnt; ++k) if (vals[k] > vals[best]) best = k;      // ascending index: first occurrence wins ties
                if (c == 0) { out->best_var_index = idx[best]; out->best_var = vals[best]; }
                else { out->best_cvar_index = idx[best]; out->best_cvar = vals[best]; }
            }
        }
    }
    return MCP_OK;
}

}  // namespace mcp

using namespace mcp;

static int historical_var_impl(mcp_handle h, const mcp_hist_params* p, const double* R, mcp_hist_out* out) {
    MCP_REQUIRE(h, p && R && out, "mcp_historical_var: NULL argument");
    MCP_REQUIRE(h, p->n_assets >= 1 && p->n_assets <= 4096, "mcp_historical_var: bad n_assets %d", p->n_assets);
    MCP_REQUIRE(h, p->n_periods >= 1, "mcp_historical_var: n_periods must be >= 1 (np.percentile of an empty series is an error)");
    MCP_REQUIRE(h, p->dtype == MCP_F32 || p->dtype == MCP_F64, "mcp_historical_var: bad dtype %d", p->dtype);
    MCP_REQUIRE(h, p->space == MCP_HOST || p->space == MCP_DEVICE, "mcp_historical_var: bad space %d", p->space);
    MCP_REQUIRE(h, p->alpha >= 0 && p->alpha <= 1, "mcp_historical_var: alpha=%g outside [0, 1]", p->alpha);
    MCP_REQUIRE(h, p->weights_in || p->n_portfolios == 0, "mcp_historical_var: weights_in is NULL");
    out->best_var_index = out->best_cvar_index = MCP_NO_INDEX;
    out->best_var = out->best_cvar = NAN;
    out->kernel_ms = 0;
    if (p->n_portfolios == 0) return MCP_OK;
    mcp_device_guard guard(h->device);
    return p->dtype == MCP_F64 ? hist_run<double>(h, p, R, out) : hist_run<float>(h, p, R, out);
}

extern "C" int mcp_historical_var(mcp_handle h, const mcp_hist_params* p, const double* R, mcp_hist_out* out) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_historical_var", [&] { return historical_var_impl(h, p, R, out); });
}
