// mcp_portfolios: host orchestration of the fused portfolio sweep (C ABI entry point).
//
// DEVICE space: one persistent launch over the whole index range on the handle's stream.
// HOST space:   the range is cut into chunks that flow through two slots (stream + device
//               buffers each): H2D of chunk c+1 overlaps the sweep of chunk c and the D2H of
//               chunk c-1, so the end-to-end rate is set by the slower of PCIe and the kernel.
// Selection: per-CTA candidates -> pf_reduce_cands (one CTA) -> replay kernel that re-evaluates
// the winners with the sweep's own arithmetic -> one small D2H of the two records.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "mcp_device.cuh"
#include "mcp_portfolio.h"

namespace mcp {

__device__ __forceinline__ void cand_merge(PfCand& b, const PfCand& o) {
    if (cand_better<double>(o.key_s, o.idx_s, b.key_s, b.idx_s)) { b.key_s = o.key_s; b.idx_s = o.idx_s; }
    if (cand_better<double>(o.key_d, o.idx_d, b.key_d, b.idx_d)) { b.key_d = o.key_d; b.idx_d = o.idx_d; }
    b.rmin = o.rmin < b.rmin ? o.rmin : b.rmin;
    b.rmax = o.rmax > b.rmax ? o.rmax : b.rmax;
}

// Merges n per-CTA candidates (and, if accumulate, the running record) into *acc.
__global__ void __launch_bounds__(256) pf_reduce_cands(const PfCand* __restrict__ cands, int n, PfCand* acc, int accumulate) {
    __shared__ PfCand sm[256];
    const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
    PfCand b{ninf, MCP_NO_INDEX, ninf, MCP_NO_INDEX, -ninf, ninf};
    for (int i = threadIdx.x; i < n; i += blockDim.x) cand_merge(b, cands[i]);
    sm[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) cand_merge(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        PfCand r = sm[0];
        if (accumulate) cand_merge(r, *acc);
        *acc = r;
    }
}

int pf_reduce_launch(mcp_context* h, const PfCand* cands, int n, PfCand* acc, int accumulate, cudaStream_t st) {
    pf_reduce_cands<<<1, 256, 0, st>>>(cands, n, acc, accumulate);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

static const int kSmallNP[] = {4, 8, 16, 24, 32};

template <typename T>
static int small_dispatch(mcp_context* h, PfJob& job, const PfReplay* rp) {
    int np = 0;
    for (int c : kSmallNP) if (job.n <= c) { np = c; break; }
    switch (np) {
#define MCP_CASE(NP) case NP: return rp ? pf_small_replay_t<T, NP>(h, job, *rp) : pf_small_launch_t<T, NP>(h, job);
        MCP_CASE(4) MCP_CASE(8) MCP_CASE(16) MCP_CASE(24) MCP_CASE(32)
#undef MCP_CASE
    }
    return mcp_fail(h, MCP_ERR_INVALID, "no register kernel for n_assets=%d", job.n);
}

int pf_small_launch(mcp_context* h, PfJob& job) {
    return job.dtype == MCP_F64 ? small_dispatch<double>(h, job, nullptr) : small_dispatch<float>(h, job, nullptr);
}
int pf_small_replay(mcp_context* h, const PfJob& job, const PfReplay& rp) {
    PfJob j = job;
    return job.dtype == MCP_F64 ? small_dispatch<double>(h, j, &rp) : small_dispatch<float>(h, j, &rp);
}

static int pf_launch(mcp_context* h, PfJob& job) {
    return job.n <= PF_SMALL_MAX_N ? pf_small_launch(h, job) : pf_large_launch(h, job);
}
static int pf_replay(mcp_context* h, const PfJob& job, const PfReplay& rp) {
    return job.n <= PF_SMALL_MAX_N ? pf_small_replay(h, job, rp) : pf_large_replay(h, job, rp);
}

}  // namespace mcp

using namespace mcp;

// Is this host pointer page-locked (cudaHostAlloc / cudaHostRegister / managed)?  Pageable memory is staged.
static bool host_is_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Staging copies between pageable caller memory and the pinned slot buffers: one core moves ~9 GB/s, the PCIe DMA next to it
// 50 GB/s, so copies of a few MB and more are split over up to four short-lived threads.
static void host_copy(void* dst, const void* src, size_t bytes) {
    constexpr size_t kMin = (size_t)2 << 20;
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t parts = std::min<size_t>(std::min<size_t>(4, hw ? hw : 1), bytes / kMin);
    if (parts < 2) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t step = (bytes / parts + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    th.reserve(parts - 1);
    for (size_t k = 1; k < parts; ++k) {
        const size_t o = k * step;
        if (o >= bytes) break;
        th.emplace_back([=] { memcpy((unsigned char*)dst + o, (const unsigned char*)src + o, std::min(step, bytes - o)); });
    }
    memcpy(dst, src, std::min(step, bytes));
    for (auto& t : th) t.join();
}

// larger key wins, ties go to the lower global index (numpy's first occurrence, app.py:672); NaN keys never win
static bool record_better(double ka, uint64_t ia, double kb, uint64_t ib) {
    if (ia == MCP_NO_INDEX || std::isnan(ka)) return false;
    if (ib == MCP_NO_INDEX || std::isnan(kb)) return true;
    return ka > kb || (ka == kb && ia < ib);
}

static void fill_selection(mcp_selection& sel, const double* rec, int n) {
    uint64_t idx;
    memcpy(&idx, &rec[0], 8);
    sel.index = idx;
    sel.key = rec[1];
    sel.ret = rec[2];
    sel.risk = rec[3];
    sel.sharpe = rec[4];
    if (sel.weights) memcpy(sel.weights, rec + PF_REC_HEADER, sizeof(double) * n);
}

static int portfolios_impl(mcp_handle h, const mcp_portfolio_params* p, const double* mu, const double* sigma, mcp_portfolio_out* out) {
    MCP_REQUIRE(h, p && mu && sigma && out, "mcp_portfolios: NULL argument");
    MCP_REQUIRE(h, p->n_assets >= 1 && p->n_assets <= 4096, "mcp_portfolios: n_assets=%d out of range [1, 4096]", p->n_assets);
    MCP_REQUIRE(h, p->dtype == MCP_F32 || p->dtype == MCP_F64, "mcp_portfolios: bad dtype %d", p->dtype);
    MCP_REQUIRE(h, p->space == MCP_HOST || p->space == MCP_DEVICE, "mcp_portfolios: bad space %d", p->space);
    MCP_REQUIRE(h, p->n_portfolios <= (1ull << 39), "mcp_portfolios: n_portfolios=%llu exceeds 2^39 per call; shard the range",
                (unsigned long long)p->n_portfolios);
    MCP_REQUIRE(h, p->max_tries >= 1, "mcp_portfolios: max_tries must be >= 1");
    MCP_REQUIRE(h, p->philox_rounds == 0 || p->philox_rounds == 10 || p->philox_rounds == 7,
                "mcp_portfolios: philox_rounds=%d (supported: 10 = default, 7)", p->philox_rounds);
    MCP_REQUIRE(h, !p->comm_merge || h->comm != nullptr, "mcp_portfolios: comm_merge needs a communicator on this handle (mcp_comm_init)");
    const int N = p->n_assets;
    for (int i = 0; i < N; ++i) {
        MCP_REQUIRE(h, std::isfinite(mu[i]), "mcp_portfolios: mean_returns[%d] is not finite", i);
        for (int j = 0; j < N; ++j) MCP_REQUIRE(h, std::isfinite(sigma[(size_t)i * N + j]), "mcp_portfolios: cov_matrix[%d,%d] is not finite", i, j);
    }
    mcp_device_guard guard(h->device);
    const size_t es = p->dtype == MCP_F64 ? 8 : 4;
    const uint64_t P = p->n_portfolios;
    const bool supplied = p->weights_in != nullptr;
    const bool empty = P == 0;             // only reachable with comm_merge: the rank still joins the collective

    out->n_accepted = 0;
    out->n_accepted_global = 0;
    out->recheck_overflow = 0;
    out->reserved = 0;
    out->kernel_ms = 0;
    out->risk_min = out->risk_max = NAN;
    for (mcp_selection* s : {&out->max_sharpe, &out->target_risk}) {
        s->index = MCP_NO_INDEX;
        s->key = s->ret = s->risk = s->sharpe = NAN;
    }
    const int K = p->n_bins;
    MCP_REQUIRE(h, K >= 0 && K <= ENV_MAX_BINS, "mcp_portfolios: n_bins=%d out of range [0, %d]", K, ENV_MAX_BINS);
    if (K > 0) {
        MCP_REQUIRE(h, out->bin_best_return && out->bin_best_index, "mcp_portfolios: envelope requested but bin outputs are NULL");
        MCP_REQUIRE(h, std::isfinite(p->risk_lo) && std::isfinite(p->risk_hi) && p->risk_hi > p->risk_lo,
                    "mcp_portfolios: envelope needs finite risk_lo < risk_hi");
        for (int b = 0; b < K; ++b) { out->bin_best_return[b] = -INFINITY; out->bin_best_index[b] = MCP_NO_INDEX; }
    }
    if (P == 0 && !p->comm_merge) return MCP_OK;          // empty arrays, no selection (app.py:719-722 on empty lists)

    // ---- scratch layout (device slot 0) ----
    const int max_blocks = h->prop.multiProcessorCount * 16;
    const size_t rec_doubles = (size_t)2 * (PF_REC_HEADER + N);
    size_t off_cands = 0;
    size_t off_run = off_cands + sizeof(PfCand) * (size_t)max_blocks * 2;      // two slots of CTA candidates
    size_t off_acc = off_run + sizeof(PfCand) * 4;                              // running[0..1], final
    size_t off_rec = off_acc + 64;
    size_t off_rows = off_rec + rec_doubles * sizeof(double);
    size_t off_lists = (off_rows + (size_t)2 * N * 8 + 255) / 256 * 256;
    const bool recheck = p->dtype == MCP_F32 && supplied && p->weights_recheck != nullptr;
    size_t total = off_lists + (recheck ? sizeof(RcLists) : 0) + 64;
    unsigned char* base = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 0, total, (void**)&base));
    PfCand* cands = (PfCand*)(base + off_cands);
    PfCand* running = (PfCand*)(base + off_run);
    unsigned long long* d_acc = (unsigned long long*)(base + off_acc);
    double* d_rec = (double*)(base + off_rec);
    void* d_rows = base + off_rows;
    RcLists* d_lists = (RcLists*)(base + off_lists);
    double rf_mu = std::fabs(p->risk_free);
    {
        double m = 0;
        for (int i = 0; i < N; ++i) m = std::max(m, std::fabs(mu[i]));
        rf_mu += m;
    }
    // envelope bins: running (f*) and per-chunk (c*) sets for each pipeline slot
    unsigned long long* env = nullptr;
    if (K > 0) MCP_CHECK(mcp_dev_reserve(h, 8, sizeof(unsigned long long) * 8 * (size_t)K, (void**)&env));
    auto fmax = [&](int s) { return env + (size_t)(0 + s) * K; };
    auto fidx = [&](int s) { return env + (size_t)(2 + s) * K; };
    auto cmax = [&](int s) { return env + (size_t)(4 + s) * K; };
    auto cidx = [&](int s) { return env + (size_t)(6 + s) * K; };
    // bins one finished chunk (its risks / returns are on the device) into slot s's running bins
    auto envelope_chunk = [&](int s, const PfJob& j, cudaStream_t ss) -> int {
        MCP_CHECK(env_reset(h, K, cmax(s), cidx(s), ss));
        MCP_CHECK(env_chunk(h, p->dtype, j.risk_out, j.ret_out, j.P, j.first, p->risk_lo, p->risk_hi, K, cmax(s), cidx(s), ss));
        return env_fold(h, K, cmax(s), cidx(s), fmax(s), fidx(s), ss);
    };

    PfJob job;
    job.n = N;
    job.dtype = p->dtype;
    job.max_tries = p->max_tries;
    job.keep_last = p->keep_last;
    // weights always lie in [0, 1]: the app's default bounds 0 / 1 (app.py:453-454) can never reject,
    // so only bounds that bite switch the rejection machinery on
    job.bounds = supplied && (p->min_weights != nullptr || p->max_weights != nullptr);   // supplied rows may be anything
    for (int i = 0; i < N; ++i) {
        if (p->min_weights) {
            MCP_REQUIRE(h, !std::isnan(p->min_weights[i]), "mcp_portfolios: min_weights[%d] is NaN", i);
            job.bounds = job.bounds || p->min_weights[i] > 0.0;
        }
        if (p->max_weights) {
            MCP_REQUIRE(h, !std::isnan(p->max_weights[i]), "mcp_portfolios: max_weights[%d] is NaN", i);
            job.bounds = job.bounds || p->max_weights[i] < 1.0;
        }
    }
    job.seed = p->seed;
    job.rounds = p->philox_rounds == 7 ? 7 : 10;
    job.rf = p->risk_free;
    job.target = p->risk_target;
    job.sigma = sigma;
    job.mu = mu;
    job.lo = p->min_weights;
    job.hi = p->max_weights;
    job.max_blocks = max_blocks;
    job.n_accepted = d_acc;

    cudaStream_t st = h->stream;
    MCP_CUDA(h, cudaMemsetAsync(d_acc, 0, 8, st));
    if (recheck) MCP_CUDA(h, cudaMemsetAsync(d_lists, 0, 16, st));
    if (K > 0) {
        MCP_CHECK(env_reset(h, K, fmax(0), fidx(0), st));
        MCP_CHECK(env_reset(h, K, fmax(1), fidx(1), st));
    }
    PfCand* final_cand = running + 2;
    double kernel_ms = 0;
    bool env_slot1 = false;

    if (empty) {
        // nothing to sweep on this rank: it still takes part in the merge below
    } else if (p->space == MCP_DEVICE) {
        // one launch, unless the envelope needs (risk, return) scratch: then chunks of 2^26
        const bool s_ret = K > 0 && !out->returns, s_risk = (K > 0 || recheck) && !out->risks, s_sharpe = recheck && !out->sharpes;
        const bool any_scratch = s_ret || s_risk || s_sharpe;
        const uint64_t chunk = any_scratch ? std::min<uint64_t>(P, 1ull << 26) : P;
        unsigned char* scratch = nullptr;
        if (any_scratch) MCP_CHECK(mcp_dev_reserve(h, 9, 3 * chunk * es, (void**)&scratch));
        job.cands = cands;
        job.stream = st;
        MCP_CUDA(h, cudaEventRecord(h->ev[0], st));
        for (uint64_t r0 = 0; r0 < P; r0 += chunk) {
            const uint64_t rows = std::min<uint64_t>(chunk, P - r0);
            auto at = [&](void* q, size_t row_bytes) -> void* { return q ? (unsigned char*)q + r0 * row_bytes : nullptr; };
            job.first = p->first_index + r0;
            job.P = rows;
            job.w_in = p->weights_in ? (const unsigned char*)p->weights_in + r0 * N * es : nullptr;
            job.w_out = at(out->weights, (size_t)N * es);
            job.ret_out = s_ret ? (void*)scratch : at(out->returns, es);
            job.risk_out = s_risk ? (void*)(scratch + chunk * es) : at(out->risks, es);
            job.sharpe_out = s_sharpe ? (void*)(scratch + 2 * chunk * es) : at(out->sharpes, es);
            job.acc_out = (uint8_t*)at(out->accepted, 1);
            MCP_CHECK(pf_launch(h, job));
            MCP_CHECK(pf_reduce_launch(h, cands, job.blocks_used, running, r0 > 0 ? 1 : 0, st));
            if (K > 0) MCP_CHECK(envelope_chunk(0, job, st));
            if (recheck) MCP_CHECK(rc_collect_launch(h, job.sharpe_out, job.risk_out, rows, job.first, running, rf_mu, p->risk_target, d_lists, st));
        }
        MCP_CUDA(h, cudaEventRecord(h->ev[1], st));
        MCP_CHECK(pf_reduce_launch(h, running, 1, final_cand, 0, st));
    } else {
        // ---- HOST space: two-slot chunk pipeline ----
        MCP_CUDA(h, cudaStreamSynchronize(st));           // the memset above must precede the side streams
        const bool want_w = out->weights != nullptr;
        const size_t in_row = supplied ? (size_t)N * es : 0;
        const bool want_ret = out->returns != nullptr || K > 0, want_risk = out->risks != nullptr || K > 0 || recheck;
        const bool want_sharpe = out->sharpes != nullptr || recheck;
        const size_t out_row = (want_w ? (size_t)N * es : 0) + (want_ret ? es : 0) + (want_risk ? es : 0) +
                               (want_sharpe ? es : 0) + (out->accepted ? 1 : 0);
        // Pageable caller memory is staged through the handle's pinned buffers (one per pipeline slot): the DMA of chunk c runs
        // at full PCIe rate while the host memcpy of chunk c-1 into (or chunk c+1 out of) the caller's pages overlaps it.  Smaller
        // chunks then, so that there is something to overlap with.  Page-locked buffers (mcp_host_alloc) are copied directly.
        const bool stage_in = in_row && !host_is_pinned(p->weights_in);
        // per output array: only the pageable ones are staged (a small pageable array -- the accepted flags, say -- must not pull the
        // page-locked 64 MB of weights through the staging buffers with it)
        const bool stg[5] = {out->weights && !host_is_pinned(out->weights), out->returns && !host_is_pinned(out->returns),
                             out->risks && !host_is_pinned(out->risks), out->sharpes && !host_is_pinned(out->sharpes),
                             out->accepted && !host_is_pinned(out->accepted)};
        const size_t staged_row = (stg[0] ? (size_t)N * es : 0) + (stg[1] ? es : 0) + (stg[2] ? es : 0) + (stg[3] ? es : 0) + (stg[4] ? 1 : 0);
        const bool stage_out = staged_row != 0;
        uint64_t chunk = P;
        // small chunks only when a substantial part of the traffic is staged (there must be something for the host memcpy to overlap)
        const size_t budget = (stage_in || staged_row * 8 > out_row) ? (size_t)16 << 20 : (size_t)96 << 20;   // bytes per slot and direction
        if (in_row) chunk = std::min<uint64_t>(chunk, std::max<size_t>(budget / in_row, 1024));
        if (out_row) chunk = std::min<uint64_t>(chunk, std::max<size_t>(budget / out_row, 1024));
        chunk = (chunk + PF_BLOCK - 1) / PF_BLOCK * PF_BLOCK;
        if (!in_row && !out_row) chunk = P;
        const uint64_t n_chunks = (P + chunk - 1) / chunk;
        bool used[2] = {false, false};
        bool timed[2] = {false, false};
        // staged output of the chunk a slot carried last: copied to the caller's arrays once the slot's stream has drained
        struct Pending { bool on = false; uint64_t r0 = 0, rows = 0; unsigned char *w = nullptr, *r = nullptr, *k = nullptr, *s = nullptr, *a = nullptr; } pend[2];
        auto flush = [&](int s) {
            Pending& q = pend[s];
            if (!q.on) return;
            if (q.w) host_copy((unsigned char*)out->weights + q.r0 * N * es, q.w, q.rows * N * es);
            if (q.r) memcpy((unsigned char*)out->returns + q.r0 * es, q.r, q.rows * es);
            if (q.k) memcpy((unsigned char*)out->risks + q.r0 * es, q.k, q.rows * es);
            if (q.s) memcpy((unsigned char*)out->sharpes + q.r0 * es, q.s, q.rows * es);
            if (q.a) memcpy(out->accepted + q.r0, q.a, q.rows);
            q.on = false;
        };
        for (uint64_t c = 0; c < n_chunks; ++c) {
            const int s = (int)(c & 1);
            cudaStream_t ss = n_chunks == 1 ? st : h->side_stream[s];
            const uint64_t r0 = c * chunk, rows = std::min<uint64_t>(chunk, P - r0);
            if (used[s]) {
                MCP_CUDA(h, cudaStreamSynchronize(ss));
                if (timed[s]) {
                    float ms = 0;
                    MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[2 * s], h->ev[2 * s + 1]));
                    kernel_ms += ms;
                }
                flush(s);
            }
            unsigned char* d_in = nullptr;
            unsigned char* d_out = nullptr;
            if (in_row) MCP_CHECK(mcp_dev_reserve(h, 1 + s, (size_t)chunk * in_row, (void**)&d_in));
            if (out_row) MCP_CHECK(mcp_dev_reserve(h, 3 + s, (size_t)chunk * (out_row + 16), (void**)&d_out));
            if (in_row) {
                const unsigned char* src = (const unsigned char*)p->weights_in + r0 * in_row;
                if (stage_in) {
                    unsigned char* pin = nullptr;
                    MCP_CHECK(mcp_pinned_reserve(h, 2 + s, (size_t)chunk * in_row, (void**)&pin));     // slots 2, 3: inputs
                    host_copy(pin, src, rows * in_row);
                    src = pin;
                }
                MCP_CUDA(h, cudaMemcpyAsync(d_in, src, rows * in_row, cudaMemcpyHostToDevice, ss));
            }
            // carve the output slot (device) and, when staging, the same layout in the slot's pinned buffer
            unsigned char* p_out = nullptr;
            if (stage_out) MCP_CHECK(mcp_pinned_reserve(h, s, (size_t)chunk * (out_row + 16), (void**)&p_out));   // slots 0, 1: outputs
            size_t o = 0;
            size_t offs[5] = {0, 0, 0, 0, 0};
            int n_carved = 0;
            auto carve = [&](bool want, size_t bytes_per_row) -> unsigned char* {
                offs[n_carved++] = o;
                if (!want) return nullptr;
                unsigned char* q = d_out + o;
                o += (chunk * bytes_per_row + 255) / 256 * 256;
                return q;
            };
            unsigned char* dw = carve(want_w, (size_t)N * es);
            unsigned char* dr = carve(want_ret, es);
            unsigned char* dk = carve(want_risk, es);
            unsigned char* ds = carve(want_sharpe, es);
            unsigned char* da = carve(out->accepted != nullptr, 1);
            job.first = p->first_index + r0;
            job.P = rows;
            job.w_in = d_in;
            job.w_out = dw;
            job.ret_out = dr;
            job.risk_out = dk;
            job.sharpe_out = ds;
            job.acc_out = da;
            job.cands = cands + (size_t)s * max_blocks;
            job.stream = ss;
            MCP_CUDA(h, cudaEventRecord(h->ev[2 * s], ss));
            MCP_CHECK(pf_launch(h, job));
            MCP_CUDA(h, cudaEventRecord(h->ev[2 * s + 1], ss));
            timed[s] = true;
            MCP_CHECK(pf_reduce_launch(h, job.cands, job.blocks_used, running + s, used[s] ? 1 : 0, ss));
            if (K > 0) {
                MCP_CHECK(envelope_chunk(s, job, ss));
                env_slot1 = env_slot1 || (s == 1 && n_chunks > 1);
            }
            if (recheck) MCP_CHECK(rc_collect_launch(h, ds, dk, rows, job.first, running + s, rf_mu, p->risk_target, d_lists, ss));
            Pending& q = pend[s];
            auto back = [&](unsigned char* dsrc, void* user, size_t row_bytes, int which, unsigned char*& staged) {
                staged = nullptr;
                if (!dsrc || !user) return cudaSuccess;
                if (stg[which]) {
                    staged = p_out + offs[which];
                    return cudaMemcpyAsync(staged, dsrc, rows * row_bytes, cudaMemcpyDeviceToHost, ss);
                }
                return cudaMemcpyAsync((unsigned char*)user + r0 * row_bytes, dsrc, rows * row_bytes, cudaMemcpyDeviceToHost, ss);
            };
            MCP_CUDA(h, back(dw, out->weights, (size_t)N * es, 0, q.w));
            MCP_CUDA(h, back(dr, out->returns, es, 1, q.r));
            MCP_CUDA(h, back(dk, out->risks, es, 2, q.k));
            MCP_CUDA(h, back(ds, out->sharpes, es, 3, q.s));
            MCP_CUDA(h, back(da, out->accepted, 1, 4, q.a));
            q.on = stage_out;
            q.r0 = r0;
            q.rows = rows;
            used[s] = true;
        }
        for (int s = 0; s < 2; ++s) {
            if (!used[s]) continue;
            cudaStream_t ss = n_chunks == 1 ? st : h->side_stream[s];
            MCP_CUDA(h, cudaStreamSynchronize(ss));
            if (timed[s]) {
                float ms = 0;
                MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[2 * s], h->ev[2 * s + 1]));
                kernel_ms += ms;
            }
            flush(s);
        }
        const int n_run = used[1] ? 2 : 1;
        MCP_CHECK(pf_reduce_launch(h, running, n_run, final_cand, 0, st));
    }


    // ---- winners: indices to the host, rows (supplied mode) to scratch, replay, records back ----
    PfCand fin;
    fin.key_s = fin.key_d = -INFINITY;
    fin.idx_s = fin.idx_d = MCP_NO_INDEX;
    fin.rmin = INFINITY;
    fin.rmax = -INFINITY;
    unsigned long long n_acc = 0;
    if (!empty) {
        MCP_CUDA(h, cudaMemcpyAsync(&fin, final_cand, sizeof fin, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(h, cudaMemcpyAsync(&n_acc, d_acc, 8, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(h, cudaStreamSynchronize(st));
        if (p->space == MCP_DEVICE) {
            float ms = 0;
            MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
            kernel_ms = ms;
        }
    }
    out->n_accepted = n_acc;
    out->n_accepted_global = n_acc;
    out->kernel_ms = kernel_ms;
    h->last_ms = kernel_ms;
    std::vector<unsigned long long> hb(2 * (size_t)std::max(K, 1), 0ull);       // local envelope bins: keys, then indices
    if (K > 0 && !empty) {
        if (env_slot1) MCP_CHECK(env_fold(h, K, fmax(1), fidx(1), fmax(0), fidx(0), st));
        MCP_CUDA(h, cudaMemcpyAsync(hb.data(), fmax(0), sizeof(unsigned long long) * K, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(h, cudaMemcpyAsync(hb.data() + K, fidx(0), sizeof(unsigned long long) * K, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(h, cudaStreamSynchronize(st));
    }
    const bool have_local = n_acc != 0 && fin.idx_s != MCP_NO_INDEX;
    std::vector<double> rec(rec_doubles, NAN);
    {
        const uint64_t none = MCP_NO_INDEX;
        memcpy(&rec[0], &none, 8);
        memcpy(&rec[PF_REC_HEADER + N], &none, 8);
    }
    if (have_local) {
        PfReplay rp;
        rp.n_sel = 2;
        rp.idx[0] = fin.idx_s;
        rp.idx[1] = fin.idx_d;
        rp.rec = d_rec;
        job.first = p->first_index;
        job.P = P;
        job.stream = st;
        const void* row_src = p->weights_in;
        size_t row_es = es;
        if (recheck) {
            // FP32 screen -> FP64 decision among the near-ties; the records are then FP64 too
            int overflow = 0;
            MCP_CHECK(rc_decide(h, p, job, fin, rf_mu, d_lists, cands, max_blocks, d_acc, rp.idx, &overflow));
            if (overflow) {
                // more near-ties than the lists hold (many duplicated / equal rows): decide on a full FP64 pass instead
                out->recheck_overflow = 1;
                MCP_CHECK(rc_full_fp64(h, p, job, cands, max_blocks, d_acc, rp.idx));
            }
            job.dtype = MCP_F64;
            row_src = p->weights_recheck;
            row_es = 8;
        }
        if (supplied) {
            for (int k = 0; k < 2; ++k) {
                const size_t row = (size_t)(rp.idx[k] - p->first_index) * N * row_es;
                MCP_CUDA(h, cudaMemcpyAsync((unsigned char*)d_rows + (size_t)k * N * row_es, (const unsigned char*)row_src + row,
                                            (size_t)N * row_es, p->space == MCP_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
            }
            rp.rows = d_rows;
        }
        MCP_CHECK(pf_replay(h, job, rp));
        MCP_CUDA(h, cudaMemcpyAsync(rec.data(), d_rec, rec_doubles * sizeof(double), cudaMemcpyDeviceToHost, st));
        MCP_CUDA(h, cudaStreamSynchronize(st));
    }
    double rmin = have_local ? fin.rmin : INFINITY, rmax = have_local ? fin.rmax : -INFINITY;

    if (p->comm_merge) {
        // ---- cross-rank merge inside the library: ONE all-gather of this rank's block
        //      [2 records | n_accepted | risk range | K bin keys | K bin indices], the same deterministic pick on every rank ----
        const size_t blk = rec_doubles + 3 + 2 * (size_t)K;                   // 8-byte words per rank
        const int world = h->comm_size;
        std::vector<double> mine(blk), all(blk * (size_t)world);
        memcpy(mine.data(), rec.data(), rec_doubles * 8);
        memcpy(&mine[rec_doubles], &n_acc, 8);
        mine[rec_doubles + 1] = rmin;
        mine[rec_doubles + 2] = rmax;
        if (K > 0) memcpy(&mine[rec_doubles + 3], hb.data(), 2 * (size_t)K * 8);
        MCP_CHECK(mcp_comm_allgather(h, mine.data(), blk * 8, all.data()));
        const int stride = PF_REC_HEADER + N;
        uint64_t best_i[2] = {MCP_NO_INDEX, MCP_NO_INDEX};
        double best_k[2] = {NAN, NAN};
        int best_r[2] = {-1, -1};
        unsigned long long total_acc = 0;
        rmin = INFINITY;
        rmax = -INFINITY;
        std::fill(hb.begin(), hb.end(), 0ull);
        for (int r = 0; r < world; ++r) {
            const double* b = all.data() + (size_t)r * blk;
            for (int c = 0; c < 2; ++c) {
                uint64_t idx;
                memcpy(&idx, &b[(size_t)c * stride], 8);
                // record key: Sharpe (larger wins) / |risk - target| (smaller wins)
                const double key = c == 0 ? b[(size_t)c * stride + 1] : -b[(size_t)c * stride + 1];
                if (record_better(key, idx, best_k[c], best_i[c])) { best_k[c] = key; best_i[c] = idx; best_r[c] = r; }
            }
            unsigned long long a8;
            memcpy(&a8, &b[rec_doubles], 8);
            total_acc += a8;
            if (a8) { rmin = std::min(rmin, b[rec_doubles + 1]); rmax = std::max(rmax, b[rec_doubles + 2]); }
            for (int k = 0; k < K; ++k) {
                unsigned long long key, idx;
                memcpy(&key, &b[rec_doubles + 3 + k], 8);
                memcpy(&idx, &b[rec_doubles + 3 + K + k], 8);
                if (key == 0ull) continue;                                     // empty bin on that rank
                if (key > hb[k] || (key == hb[k] && idx < hb[K + k])) { hb[k] = key; hb[K + k] = idx; }
            }
        }
        for (int c = 0; c < 2; ++c)
            if (best_r[c] >= 0) memcpy(&rec[(size_t)c * stride], all.data() + (size_t)best_r[c] * blk + (size_t)c * stride, (size_t)stride * 8);
        out->n_accepted_global = total_acc;
        n_acc = total_acc;
    }
    if (K > 0) {
        for (int b = 0; b < K; ++b) {
            if (hb[b] == 0ull) continue;                 // empty bin: -inf / MCP_NO_INDEX
            out->bin_best_return[b] = mcp_key_to_value(hb[b], p->dtype);
            out->bin_best_index[b] = hb[K + b];
        }
    }
    if (n_acc == 0) return MCP_OK;
    out->risk_min = rmin;
    out->risk_max = rmax;
    fill_selection(out->max_sharpe, rec.data(), N);
    fill_selection(out->target_risk, rec.data() + PF_REC_HEADER + N, N);
    return MCP_OK;
}


static int envelope_arrays_impl(mcp_handle h, int dtype, const void* risks, const void* returns, uint64_t n, uint64_t first_index,
                                double risk_lo, double risk_hi, int K, double* bin_best_return, uint64_t* bin_best_index) {
    MCP_REQUIRE(h, dtype == MCP_F32 || dtype == MCP_F64, "mcp_envelope_arrays: bad dtype %d", dtype);
    MCP_REQUIRE(h, K >= 1 && K <= ENV_MAX_BINS, "mcp_envelope_arrays: n_bins=%d out of range [1, %d]", K, ENV_MAX_BINS);
    MCP_REQUIRE(h, bin_best_return && bin_best_index, "mcp_envelope_arrays: bin outputs are NULL");
    MCP_REQUIRE(h, std::isfinite(risk_lo) && std::isfinite(risk_hi) && risk_hi > risk_lo, "mcp_envelope_arrays: needs finite risk_lo < risk_hi");
    MCP_REQUIRE(h, (risks && returns) || n == 0, "mcp_envelope_arrays: NULL arrays");
    for (int b = 0; b < K; ++b) { bin_best_return[b] = -INFINITY; bin_best_index[b] = MCP_NO_INDEX; }
    if (n == 0) return MCP_OK;
    mcp_device_guard guard(h->device);
    cudaStream_t st = h->stream;
    const size_t es = dtype == MCP_F64 ? 8 : 4;
    unsigned long long* env = nullptr;
    MCP_CHECK(mcp_dev_reserve(h, 8, sizeof(unsigned long long) * 8 * (size_t)K, (void**)&env));
    unsigned long long *fmax = env, *fidx = env + 2 * (size_t)K, *cmax = env + 4 * (size_t)K, *cidx = env + 6 * (size_t)K;
    MCP_CHECK(env_reset(h, K, fmax, fidx, st));
    MCP_CUDA(h, cudaEventRecord(h->ev[0], st));
    (void)es;
    MCP_CHECK(env_reset(h, K, cmax, cidx, st));
    MCP_CHECK(env_chunk(h, dtype, risks, returns, n, first_index, risk_lo, risk_hi, K, cmax, cidx, st));   // grid-stride over all n
    MCP_CHECK(env_fold(h, K, cmax, cidx, fmax, fidx, st));
    MCP_CUDA(h, cudaEventRecord(h->ev[1], st));
    std::vector<unsigned long long> hb(2 * (size_t)K);
    MCP_CUDA(h, cudaMemcpyAsync(hb.data(), fmax, sizeof(unsigned long long) * K, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaMemcpyAsync(hb.data() + K, fidx, sizeof(unsigned long long) * K, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));
    float ms = 0;
    MCP_CUDA(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->last_ms = ms;
    for (int b = 0; b < K; ++b) {
        if (hb[b] == 0ull) continue;                         // empty bin
        bin_best_return[b] = mcp_key_to_value(hb[b], dtype);
        bin_best_index[b] = hb[K + b];
    }
    return MCP_OK;
}

extern "C" int mcp_portfolios(mcp_handle h, const mcp_portfolio_params* p, const double* mu, const double* sigma,
                              mcp_portfolio_out* out) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_portfolios", [&] { return portfolios_impl(h, p, mu, sigma, out); });
}

extern "C" int mcp_envelope_arrays(mcp_handle h, int dtype, const void* risks, const void* returns, uint64_t n, uint64_t first_index,
                                   double risk_lo, double risk_hi, int K, double* bin_best_return, uint64_t* bin_best_index) {
    if (!h) return MCP_ERR_INVALID;
    return mcp_guarded(h, "mcp_envelope_arrays",
                       [&] { return envelope_arrays_impl(h, dtype, risks, returns, n, first_index, risk_lo, risk_hi, K, bin_best_return, bin_best_index); });
}
