// mcp_quantiles: exact VaR / CVaR over a (possibly sharded) vector of simulated returns.
//
// Conventions of the reference's var / cvar (app.py:258-263): VaR = np.percentile(x, (1-a)*100)
// with numpy's default 'linear' method -- h = (n-1) q, lerp between the floor(h)-th and next
// order statistics -- and CVaR = mean(x[x <= VaR]) (VaR if empty).  The order statistics are
// found exactly by an MSB-first radix select on order-preserving integer keys: each pass
// histograms 11 key bits of the values whose higher bits match the target's prefix
// (warp-aggregated shared-memory atomics via __match_any_sync), the per-rank histograms are
// summed (NCCL all-reduce supplied by the host as a callback when the vector is sharded), and a
// scan picks the digit -- on the device (select_advance_kernel: the passes, their all-reduces and the digit selection run
// back to back on one stream) or, with a synchronous all-reduce callback, on the host.  3 passes for FP32, 6 for FP64; the vector (40 MB at C4) is
// L2-resident after the first pass.  A last pass accumulates the tail sums in FP64.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

#include "mcp_context.h"
#include "mcp_device.cuh"

namespace mcp {

constexpr int SEL_BITS = MCP_SEL_BITS;
constexpr int SEL_BLOCK = 256;

struct SelSlots {
    uint64_t prefix[MCP_MAX_TARGETS];
};

template <typename T> struct KeyOf;
template <> struct KeyOf<float> {
    using K = uint32_t;
    static constexpr int BITS = 32;
    static __device__ __forceinline__ K key(float v) { return f32_to_key(__float_as_uint(v)); }
};
template <> struct KeyOf<double> {
    using K = uint64_t;
    static constexpr int BITS = 64;
    static __device__ __forceinline__ K key(double v) { return f64_to_key((uint64_t)__double_as_longlong(v)); }
};

// hist[slot][digit] += #{ i : key_i >> (shift + bits) == prefix[slot], digit(key_i) = digit }
template <typename T>
__global__ void __launch_bounds__(SEL_BLOCK) select_hist_kernel(const T* __restrict__ v, uint64_t n, int n_slots,
                                                                const __grid_constant__ SelSlots slots, int shift, int bits,
                                                                unsigned long long* __restrict__ hist,
                                                                const unsigned long long* __restrict__ dev_prefix) {
    using K = typename KeyOf<T>::K;
    extern __shared__ unsigned int sh[];
    __shared__ K s_prefix[MCP_MAX_TARGETS];           // per-slot prefix: kernel parameter, or the device-side select state
    if (threadIdx.x < MCP_MAX_TARGETS) s_prefix[threadIdx.x] = (K)(dev_prefix ? dev_prefix[threadIdx.x] : slots.prefix[threadIdx.x]);
    const int nb = 1 << bits;
    for (int i = threadIdx.x; i < n_slots * nb; i += SEL_BLOCK) sh[i] = 0;
    __syncthreads();
    const int hi_shift = shift + bits;
    const bool all_match = hi_shift >= KeyOf<T>::BITS;
    const K mask = (K)(nb - 1);
    const int lane = threadIdx.x & 31;
    for (uint64_t i = (uint64_t)blockIdx.x * SEL_BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * SEL_BLOCK) {
        const K k = KeyOf<T>::key(v[i]);
        const unsigned digit = (unsigned)((k >> shift) & mask);
        const K hi = all_match ? (K)0 : (K)(k >> (all_match ? 0 : hi_shift));
        for (int s = 0; s < n_slots; ++s) {
            if (all_match || hi == s_prefix[s]) {
                const unsigned act = __activemask();
                const unsigned peers = __match_any_sync(act, digit);
                if (lane == __ffs(peers) - 1) atomicAdd(&sh[s * nb + digit], (unsigned)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_slots * nb; i += SEL_BLOCK) {
        const unsigned c = sh[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
}

// sums[t] += sum of v_i <= thr[t] (FP64), counts[t] += their number; thr lives in device memory (written by
// select_finish_kernel, or uploaded by the host state machine)
template <typename T>
__global__ void __launch_bounds__(SEL_BLOCK) tail_sum_kernel(const T* __restrict__ v, uint64_t n, int nt,
                                                             const double* __restrict__ thr, double* __restrict__ sums,
                                                             double* __restrict__ counts) {
    double s[MCP_MAX_ALPHAS], c[MCP_MAX_ALPHAS];
    struct { double thr[MCP_MAX_ALPHAS]; } a;
#pragma unroll
    for (int t = 0; t < MCP_MAX_ALPHAS; ++t) { s[t] = c[t] = 0.0; a.thr[t] = t < nt ? thr[t] : 0.0; }
    for (uint64_t i = (uint64_t)blockIdx.x * SEL_BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * SEL_BLOCK) {
        const double x = (double)v[i];
#pragma unroll
        for (int t = 0; t < MCP_MAX_ALPHAS; ++t)
            if (t < nt && x <= a.thr[t]) { s[t] += x; c[t] += 1.0; }
    }
    __shared__ double ws[SEL_BLOCK / 32][2 * MCP_MAX_ALPHAS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < MCP_MAX_ALPHAS; ++t) {
        double a0 = s[t], a1 = c[t];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, m);
            a1 += __shfl_xor_sync(0xffffffffu, a1, m);
        }
        if (lane == 0) { ws[warp][2 * t] = a0; ws[warp][2 * t + 1] = a1; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * nt) {
        double acc = 0;
        for (int w = 0; w < SEL_BLOCK / 32; ++w) acc += ws[w][threadIdx.x];
        if (acc != 0.0) atomicAdd((threadIdx.x & 1) ? &counts[threadIdx.x >> 1] : &sums[threadIdx.x >> 1], acc);
    }
}

// Device-side twin of mcp_select_advance: one CTA per target (slot t = target t, no prefix sharing).  Finds the
// digit d with cum(d) <= rank < cum(d) + hist[d], then rank -= cum(d), prefix = prefix << bits | d.
struct SelDevState {
    unsigned long long prefix[MCP_MAX_TARGETS];
    unsigned long long rank[MCP_MAX_TARGETS];
    int error;                                         // a rank fell outside its prefix's population
};

__global__ void __launch_bounds__(256) select_advance_kernel(const unsigned long long* __restrict__ hist, int bits, SelDevState* st, int shared_row) {
    const int t = blockIdx.x, nb = 1 << bits, per = (nb + 255) / 256;        // bins per thread (<= 8)
    const unsigned long long* row = hist + (shared_row ? 0 : (size_t)t * nb);   // first pass: every target has the empty prefix
    unsigned long long mine = 0;
    for (int j = 0; j < per; ++j) {
        const int d = threadIdx.x * per + j;
        if (d < nb) mine += row[d];
    }
    // exclusive scan of the per-thread sums over the CTA
    __shared__ unsigned long long wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = mine;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, m);
        if (lane >= m) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned long long base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    const unsigned long long excl = base + inc - mine, rank = st->rank[t];
    __shared__ int found;
    if (threadIdx.x == 0) found = 0;
    __syncthreads();
    if (rank >= excl && rank < excl + mine) {          // exactly one thread owns the rank
        unsigned long long cum = excl;
        for (int j = 0; j < per; ++j) {
            const int d = threadIdx.x * per + j;
            const unsigned long long c = d < nb ? row[d] : 0ull;
            if (rank < cum + c) {
                st->rank[t] = rank - cum;
                st->prefix[t] = (st->prefix[t] << bits) | (unsigned long long)d;
                found = 1;
                break;
            }
            cum += c;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && !found) st->error = 1;
}

// VaR from the final prefixes, on the device (so that a sharded call needs no host round trip between the select passes and
// the tail pass): numpy's _lerp on the two exact order statistics, with explicitly rounded operations -- the host code this
// replaces is compiled without FMA contraction, and the results are compared bit for bit with np.percentile.
struct FinishArgs {
    int n_alphas, dtype;
    int lo_t[MCP_MAX_ALPHAS], hi_t[MCP_MAX_ALPHAS];
    double gamma[MCP_MAX_ALPHAS];
};
__device__ __forceinline__ double key_value_dev(unsigned long long key, int dtype) {
    if (dtype == MCP_F64) return __longlong_as_double((long long)key_to_f64(key));
    return (double)__uint_as_float(key_to_f32((uint32_t)key));
}
__global__ void select_finish_kernel(const SelDevState* st, const __grid_constant__ FinishArgs f, double* var_thr) {
    const int a = threadIdx.x;
    if (a >= f.n_alphas) return;
    const double lo = key_value_dev(st->prefix[f.lo_t[a]], f.dtype), hi = key_value_dev(st->prefix[f.hi_t[a]], f.dtype);
    const double t = f.gamma[a], diff = __dsub_rn(hi, lo);
    double r = __dadd_rn(lo, __dmul_rn(diff, t));
    if (t >= 0.5) r = __dsub_rn(hi, __dmul_rn(diff, __dsub_rn(1.0, t)));
    var_thr[a] = r;
}

static void assign_slots(mcp_select_state* s) {
    s->n_slots = 0;
    for (int t = 0; t < s->n_targets; ++t) {
        int found = -1;
        for (int k = 0; k < s->n_slots; ++k)
            if (s->slot_prefix[k] == s->prefix[t]) { found = k; break; }
        if (found < 0) {
            found = s->n_slots++;
            s->slot_prefix[found] = s->prefix[t];
        }
        s->slot_of[t] = found;
    }
}

static int grid_for(mcp_context* h, uint64_t n, int per_sm) {
    uint64_t g = (n + SEL_BLOCK - 1) / SEL_BLOCK;
    const uint64_t cap = (uint64_t)h->prop.multiProcessorCount * per_sm;
    return (int)std::max<uint64_t>(1, std::min(g, cap));
}

}  // namespace mcp

using namespace mcp;

extern "C" {

int mcp_select_init(mcp_select_state* s, int key_bits, const uint64_t* ranks, int n_targets) {
    if (!s || !ranks || (key_bits != 32 && key_bits != 64) || n_targets < 1 || n_targets > MCP_MAX_TARGETS) return MCP_ERR_INVALID;
    memset(s, 0, sizeof *s);
    s->key_bits = key_bits;
    s->n_targets = n_targets;
    for (int t = 0; t < n_targets; ++t) s->rank[t] = ranks[t];
    assign_slots(s);
    return MCP_OK;
}

int mcp_select_pass_bits(const mcp_select_state* s) {
    if (!s) return 0;
    const int left = s->key_bits - s->bits_done;
    return left < SEL_BITS ? left : SEL_BITS;
}

int mcp_select_advance(mcp_select_state* s, const uint64_t* hist) {
    if (!s || !hist) return MCP_ERR_INVALID;
    const int bits = mcp_select_pass_bits(s);
    if (bits <= 0) return MCP_ERR_INVALID;
    const int nb = 1 << bits;
    for (int t = 0; t < s->n_targets; ++t) {
        const uint64_t* hrow = hist + (size_t)s->slot_of[t] * nb;
        uint64_t cum = 0;
        int d = 0;
        for (; d < nb; ++d) {
            if (cum + hrow[d] > s->rank[t]) break;
            cum += hrow[d];
        }
        if (d == nb) return MCP_ERR_INVALID;      // rank beyond the population of this prefix
        s->rank[t] -= cum;
        s->prefix[t] = (s->prefix[t] << bits) | (uint64_t)d;
    }
    s->bits_done += bits;
    assign_slots(s);
    return MCP_OK;
}

double mcp_key_to_value(uint64_t key, int dtype) {
    if (dtype == MCP_F64) {
        const uint64_t b = key_to_f64(key);
        double d;
        memcpy(&d, &b, 8);
        return d;
    }
    const uint32_t b = key_to_f32((uint32_t)key);
    float f;
    memcpy(&f, &b, 4);
    return (double)f;
}

int mcp_select_hist(mcp_handle h, const void* values_dev, int dtype, uint64_t n, const mcp_select_state* s,
                    uint64_t* hist_dev) {
    if (!h) return MCP_ERR_INVALID;
    MCP_REQUIRE(h, s && hist_dev && (values_dev || n == 0), "mcp_select_hist: NULL argument");
    MCP_REQUIRE(h, (dtype == MCP_F32 && s->key_bits == 32) || (dtype == MCP_F64 && s->key_bits == 64),
                "mcp_select_hist: dtype %d does not match key_bits %d", dtype, s->key_bits);
    mcp_device_guard guard(h->device);
    const int bits = mcp_select_pass_bits(s);
    MCP_REQUIRE(h, bits > 0, "mcp_select_hist: selection already finished");
    const int nb = 1 << bits;
    const int shift = s->key_bits - s->bits_done - bits;
    cudaStream_t st = h->stream;
    MCP_CUDA(h, cudaMemsetAsync(hist_dev, 0, sizeof(uint64_t) * s->n_slots * nb, st));
    if (n == 0) return MCP_OK;
    SelSlots slots;
    memset(&slots, 0, sizeof slots);
    for (int k = 0; k < s->n_slots; ++k) slots.prefix[k] = s->slot_prefix[k];
    const size_t smem = sizeof(unsigned int) * s->n_slots * nb;
    if (dtype == MCP_F64) {
        if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(select_hist_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        select_hist_kernel<double><<<grid_for(h, n, 4), SEL_BLOCK, smem, st>>>((const double*)values_dev, n, s->n_slots, slots, shift, bits,
                                                                                (unsigned long long*)hist_dev, nullptr);
    } else {
        if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(select_hist_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        select_hist_kernel<float><<<grid_for(h, n, 4), SEL_BLOCK, smem, st>>>((const float*)values_dev, n, s->n_slots, slots, shift, bits,
                                                                               (unsigned long long*)hist_dev, nullptr);
    }
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

}  // extern "C" (reopened below)

// The select + tail passes on device-resident values.  `allreduce` may be NULL (single shard), a caller callback, or
// MCP_ALLREDUCE_COMM (the handle's NCCL communicator, always stream-ordered).  `hist0_dev` (optional): this shard's
// histogram of the FIRST radix digit, 1 << SEL_BITS uint64 counts, already filled on the stream by the producer of the values
// (the path kernels do it in their epilogue): pass 0 then skips its sweep over the data.
int mcp_quantiles_device(mcp_context* h, const void* v, int dtype, uint64_t n, uint64_t n_total, const double* alphas, int n_alphas,
                         double* var_out, double* cvar_out, mcp_allreduce_fn allreduce, void* user,
                         const unsigned long long* hist0_dev, double* ms_out) {
    const bool use_comm = allreduce == MCP_ALLREDUCE_COMM;
    if (use_comm) MCP_REQUIRE(h, h->comm != nullptr, "mcp_quantiles: MCP_ALLREDUCE_COMM needs a communicator on this handle (mcp_comm_init)");
    if (!allreduce) n_total = n;
    MCP_REQUIRE(h, n_total >= 1 && n_total >= n, "mcp_quantiles: empty input (np.percentile of an empty array is an error)");
    cudaStream_t st = h->stream;
    // reduction over ranks of a device buffer, on this stream
    auto reduce = [&](void* buf, size_t count, int kind) -> int {
        if (!allreduce) return MCP_OK;
        if (use_comm) return mcp_comm_allreduce_dev(h, buf, count, kind == 0 ? MCP_REDUCE_U64_SUM : MCP_REDUCE_F64_SUM, st);
        if (allreduce(buf, count, kind, user) != 0) return mcp_fail(h, MCP_ERR_INVALID, "mcp_quantiles: allreduce callback failed");
        return MCP_OK;
    };
    // ---- order-statistic ranks: numpy 'linear' (virtual index (n-1) q, q = percent / 100) ----
    uint64_t ranks[MCP_MAX_TARGETS];
    FinishArgs fin;
    memset(&fin, 0, sizeof fin);
    fin.n_alphas = n_alphas;
    fin.dtype = dtype;
    int nt = 0;
    auto add_rank = [&](uint64_t r) {
        for (int k = 0; k < nt; ++k) if (ranks[k] == r) return k;
        ranks[nt] = r;
        return nt++;
    };
    for (int a = 0; a < n_alphas; ++a) {
        MCP_REQUIRE(h, alphas[a] >= 0.0 && alphas[a] <= 1.0, "mcp_quantiles: alpha[%d]=%g outside [0, 1]", a, alphas[a]);
        const double percent = (1 - alphas[a]) * 100;          // app.py:259, same FP64 expression
        const double q = percent / 100.0;
        const double hidx = (double)(n_total - 1) * q;
        uint64_t lo, hi;
        if (hidx >= (double)(n_total - 1)) lo = hi = n_total - 1;
        else if (hidx < 0) lo = hi = 0;
        else { lo = (uint64_t)std::floor(hidx); hi = lo + 1; }
        fin.gamma[a] = hidx - std::floor(hidx);
        fin.lo_t[a] = add_rank(lo);
        fin.hi_t[a] = add_rank(hi);
    }
    // device block: histograms | results (state, thresholds = VaR, tail sums, tail counts): ONE copy back
    const size_t hist_elems = (size_t)MCP_MAX_TARGETS << SEL_BITS;
    struct Results {
        SelDevState state;
        double var[MCP_MAX_ALPHAS];
        double sums[MCP_MAX_ALPHAS];
        double counts[MCP_MAX_ALPHAS];
    };
    unsigned long long* d_hist = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 5, hist_elems * 8 + sizeof(Results) + 64, (void**)&d_hist));
    Results* d_res = (Results*)(d_hist + hist_elems);
    Results* h_res = nullptr;
    MCP_CHECK(mcp_pinned_reserve(h, 4, sizeof(Results), (void**)&h_res));
    const bool device_resident = !allreduce || use_comm || h->allreduce_stream_ordered;
    MCP_CUDA(h, cudaEventRecord(h->ev[0], st));
    memset(h_res, 0, sizeof *h_res);
    auto launch_hist = [&](int slots_now, int shift, int bits, const unsigned long long* dev_prefix, const SelSlots& slots) -> int {
        const size_t cnt = (size_t)slots_now << bits;
        const size_t smem = sizeof(unsigned int) * cnt;
        if (dtype == MCP_F64) {
            if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(select_hist_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            select_hist_kernel<double><<<grid_for(h, n, 4), SEL_BLOCK, smem, st>>>((const double*)v, n, slots_now, slots, shift, bits, d_hist, dev_prefix);
        } else {
            if (smem > 40 * 1024) MCP_CUDA(h, cudaFuncSetAttribute(select_hist_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            select_hist_kernel<float><<<grid_for(h, n, 4), SEL_BLOCK, smem, st>>>((const float*)v, n, slots_now, slots, shift, bits, d_hist, dev_prefix);
        }
        MCP_CUDA(h, cudaGetLastError());
        h->launches++;
        return MCP_OK;
    };
    if (device_resident) {
        // ---- device-resident refinement: histogram -> (all-reduce on this stream) -> digit selection, pass after pass; then the
        // interpolation and the tail sums, still on the stream; one host wait at the very end.  One slot per target.
        const int key_bits = dtype == MCP_F64 ? 64 : 32;
        for (int t = 0; t < nt; ++t) h_res->state.rank[t] = ranks[t];
        MCP_CUDA(h, cudaMemcpyAsync(d_res, h_res, sizeof(Results), cudaMemcpyHostToDevice, st));
        SelSlots unused;
        memset(&unused, 0, sizeof unused);
        for (int done = 0; done < key_bits;) {
            const int bits = std::min(SEL_BITS, key_bits - done), shift = key_bits - done - bits;
            const int slots_now = done == 0 ? 1 : nt;                 // pass 0: one histogram serves all targets
            const size_t cnt = (size_t)slots_now << bits;
            if (done == 0 && hist0_dev) {
                MCP_CUDA(h, cudaMemcpyAsync(d_hist, hist0_dev, cnt * 8, cudaMemcpyDeviceToDevice, st));
            } else {
                MCP_CUDA(h, cudaMemsetAsync(d_hist, 0, cnt * 8, st));
                if (n) MCP_CHECK(launch_hist(slots_now, shift, bits, d_res->state.prefix, unused));
            }
            MCP_CHECK(reduce(d_hist, cnt, 0));
            select_advance_kernel<<<nt, 256, 0, st>>>(d_hist, bits, &d_res->state, done == 0 ? 1 : 0);
            MCP_CUDA(h, cudaGetLastError());
            h->launches++;
            done += bits;
        }
        select_finish_kernel<<<1, 32, 0, st>>>(&d_res->state, fin, d_res->var);
        MCP_CUDA(h, cudaGetLastError());
        h->launches++;
    } else {
        // ---- synchronous callback: the host state machine (mcp_select_*) scans the all-reduced histograms ----
        mcp_select_state sel;
        MCP_CHECK(mcp_select_init(&sel, dtype == MCP_F64 ? 64 : 32, ranks, nt) == MCP_OK ? MCP_OK
                  : mcp_fail(h, MCP_ERR_INVALID, "mcp_quantiles: select init failed"));
        std::vector<uint64_t> h_hist(hist_elems);
        while (mcp_select_pass_bits(&sel) > 0) {
            const int bits = mcp_select_pass_bits(&sel);
            const size_t cnt = (size_t)sel.n_slots << bits;
            MCP_CHECK(mcp_select_hist(h, v, dtype, n, &sel, (uint64_t*)d_hist));
            MCP_CUDA(h, cudaStreamSynchronize(st));
            MCP_CHECK(reduce(d_hist, cnt, 0));
            MCP_CUDA(h, cudaMemcpyAsync(h_hist.data(), d_hist, cnt * 8, cudaMemcpyDeviceToHost, st));
            MCP_CUDA(h, cudaStreamSynchronize(st));
            if (mcp_select_advance(&sel, h_hist.data()) != MCP_OK)
                return mcp_fail(h, MCP_ERR_INVALID, "mcp_quantiles: rank outside the population (n_total=%llu inconsistent with the data?)",
                                (unsigned long long)n_total);
        }
        // numpy's _lerp on the two exact order statistics
        for (int a = 0; a < n_alphas; ++a) {
            const double lo = mcp_key_to_value(sel.prefix[fin.lo_t[a]], dtype);
            const double hi = mcp_key_to_value(sel.prefix[fin.hi_t[a]], dtype);
            const double t = fin.gamma[a], diff = hi - lo;
            double r = lo + diff * t;
            if (t >= 0.5) r = hi - diff * (1 - t);
            h_res->var[a] = r;
        }
        MCP_CUDA(h, cudaMemcpyAsync(d_res, h_res, sizeof(Results), cudaMemcpyHostToDevice, st));
    }
    // ---- CVaR: FP64 tail sums (thresholds = the VaR values, read from device memory) ----
    if (n) {
        if (dtype == MCP_F64) tail_sum_kernel<double><<<grid_for(h, n, 8), SEL_BLOCK, 0, st>>>((const double*)v, n, n_alphas, d_res->var, d_res->sums, d_res->counts);
        else tail_sum_kernel<float><<<grid_for(h, n, 8), SEL_BLOCK, 0, st>>>((const float*)v, n, n_alphas, d_res->var, d_res->sums, d_res->counts);
        MCP_CUDA(h, cudaGetLastError());
        h->launches++;
    }
    if (allreduce) {
        if (!device_resident) MCP_CUDA(h, cudaStreamSynchronize(st));
        static_assert(offsetof(Results, counts) == offsetof(Results, sums) + sizeof(double) * MCP_MAX_ALPHAS, "sums and counts are reduced as one buffer");
        MCP_CHECK(reduce(d_res->sums, 2 * MCP_MAX_ALPHAS, 1));
    }
    MCP_CUDA(h, cudaEventRecord(h->ev[1], st));
    MCP_CUDA(h, cudaMemcpyAsync(h_res, d_res, sizeof(Results), cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));
    if (use_comm) MCP_CHECK(mcp_comm_check(h));
    if (device_resident && h_res->state.error)
        return mcp_fail(h, MCP_ERR_INVALID, "mcp_quantiles: rank outside the population (n_total=%llu inconsistent with the data?)",
                        (unsigned long long)n_total);
    for (int a = 0; a < n_alphas; ++a) {
        var_out[a] = h_res->var[a];
        const double cnt = h_res->counts[a];
        cvar_out[a] = cnt > 0 ? h_res->sums[a] / cnt : var_out[a];
    }
    float ms = 0;
    MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->last_ms = ms;
    if (ms_out) *ms_out = ms;
    return MCP_OK;
}

static int quantiles_impl(mcp_handle h, const void* values, int space, int dtype, uint64_t n, uint64_t n_total,
                          const double* alphas, int n_alphas, double* var_out, double* cvar_out,
                          mcp_allreduce_fn allreduce, void* user) {
    MCP_REQUIRE(h, alphas && var_out && cvar_out, "mcp_quantiles: NULL argument");
    MCP_REQUIRE(h, values || n == 0, "mcp_quantiles: values is NULL");
    MCP_REQUIRE(h, n_alphas >= 1 && n_alphas <= MCP_MAX_ALPHAS, "mcp_quantiles: n_alphas=%d out of range [1, %d]", n_alphas, MCP_MAX_ALPHAS);
    MCP_REQUIRE(h, dtype == MCP_F32 || dtype == MCP_F64, "mcp_quantiles: bad dtype %d", dtype);
    MCP_REQUIRE(h, space == MCP_HOST || space == MCP_DEVICE, "mcp_quantiles: bad space %d", space);
    mcp_device_guard guard(h->device);
    const size_t es = dtype == MCP_F64 ? 8 : 4;
    const void* v = values;
    if (space == MCP_HOST && n) {
        void* d = nullptr;
        MCP_CHECK(mcp_dev_reserve(h, 3, n * es, &d));
        MCP_CUDA(h, cudaMemcpyAsync(d, values, n * es, cudaMemcpyHostToDevice, h->stream));
        v = d;
    }
    return mcp_quantiles_device(h, v, dtype, n, n_total, alphas, n_alphas, var_out, cvar_out, allreduce, user, nullptr, nullptr);
}

extern "C" {

int mcp_quantiles(mcp_handle h, const void* values, int space, int dtype, uint64_t n, uint64_t n_total,
                  const double* alphas, int n_alphas, double* var_out, double* cvar_out,
                  mcp_allreduce_fn allreduce, void* user) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_quantiles",
                       [&] { return quantiles_impl(h, values, space, dtype, n, n_total, alphas, n_alphas, var_out, cvar_out, allreduce, user); });
}

int mcp_set_allreduce_stream_ordered(mcp_handle h, int on) {
    if (!h) return MCP_ERR_INVALID;
    h->allreduce_stream_ordered = on != 0;
    return MCP_OK;
}

}  // extern "C"
