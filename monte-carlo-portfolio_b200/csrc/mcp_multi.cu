// One call, n GPUs: the host side of the single-process multi-GPU mode in C++ (SURVEY.md 8(b) `devices=`, 8(e)).
//
// The reference is ONE Streamlit process (Procfile:1) whose script thread runs the loop app.py:682-722, so the drop-in cannot ask
// its caller for one process per GPU.  mcp_portfolios_multi / mcp_paths_stats_multi take n handles that share a communicator
// (mcp_comm_init_all), cut the job's global index range into n contiguous blocks -- block r = base + (r < rem) units, the same
// partition as mcportfolio.dist.shard_range -- and run the ordinary entry point on every handle from its own host thread with
// comm_merge set, so picks, counts, risk range, envelope bins, histograms and tail sums are merged over NVLink by the library
// itself.  The Philox counter is the global index: the result is the one-GPU result.  Host threads live in the library, not in
// the caller's interpreter: a Python caller pays one ctypes call (and one GIL release) per job instead of n.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mcp_context.h"

// One persistent host thread per handle (started by the first multi call that needs it, joined by mcp_destroy): posting a job
// to an idle worker costs a condition-variable wake-up, and all workers wake in parallel -- creating n - 1 threads per call cost
// ~0.15 ms at n = 8, which is 3 % of a 4.4 ms path step.
struct mcp_worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> task;
    bool has_task = false, done = true, stop = false;
    int rc = MCP_OK;

    void loop() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return has_task || stop; });
            if (stop) return;
            std::function<int()> t = std::move(task);
            has_task = false;
            lk.unlock();
            int r;
            try { r = t(); } catch (...) { r = MCP_ERR_INVALID; }
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    }
    void post(std::function<int()> t) {
        {
            std::lock_guard<std::mutex> lk(m);
            task = std::move(t);
            has_task = true;
            done = false;
        }
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};

void mcp_worker_release(mcp_context* h) {
    if (!h->worker) return;
    {
        std::lock_guard<std::mutex> lk(h->worker->m);
        h->worker->stop = true;
    }
    h->worker->cv.notify_all();
    if (h->worker->th.joinable()) h->worker->th.join();
    delete h->worker;
    h->worker = nullptr;
}

static mcp_worker* worker_of(mcp_context* h) {
    if (!h->worker) {
        h->worker = new mcp_worker();
        h->worker->th = std::thread([w = h->worker] { w->loop(); });
    }
    return h->worker;
}

namespace {

struct Shard { uint64_t first, count; };

Shard shard_of(uint64_t total, int rank, int world) {
    const uint64_t base = total / (uint64_t)world, rem = total % (uint64_t)world;
    const uint64_t r = (uint64_t)rank;
    return Shard{r * base + (r < rem ? r : rem), base + (r < rem ? 1 : 0)};
}

int check_group(mcp_handle* hs, int n, const char* what) {
    if (!hs || n < 1 || !hs[0]) return MCP_ERR_INVALID;
    for (int r = 0; r < n; ++r) {
        if (!hs[r]) return mcp_fail(hs[0], MCP_ERR_INVALID, "%s: handle %d is NULL", what, r);
        if (n > 1 && (hs[r]->comm == nullptr || hs[r]->comm_size != n || hs[r]->comm_rank != r))
            return mcp_fail(hs[0], MCP_ERR_COMM, "%s: handle %d is not rank %d of a %d-rank communicator (mcp_comm_init_all)", what, r, r, n);
    }
    return MCP_OK;
}

// run body(r) for r = 0 .. n-1, rank 0 on the calling thread; returns the first failure (its message copied to handle 0)
template <typename F>
int run_ranks(mcp_handle* hs, int n, F&& body) {
    std::vector<int> rc(n, MCP_OK);
    for (int r = 1; r < n; ++r) worker_of(hs[r])->post([&body, r] { return body(r); });
    rc[0] = body(0);
    for (int r = 1; r < n; ++r) rc[r] = hs[r]->worker->wait();
    for (int r = 0; r < n; ++r) {
        if (rc[r] != MCP_OK) {
            if (r != 0) {
                const std::string msg = hs[r]->err;
                mcp_fail(hs[0], rc[r], "device %d: %s", hs[r]->device, msg.c_str());
            }
            return rc[r];
        }
    }
    return MCP_OK;
}

}  // namespace

static int portfolios_multi_impl(mcp_handle* hs, int n, const mcp_portfolio_params* p, const double* mu, const double* sigma,
                                 mcp_portfolio_out* out, double* kernel_ms_per_device) {
    MCP_CHECK(check_group(hs, n, "mcp_portfolios_multi"));
    mcp_handle h0 = hs[0];
    MCP_REQUIRE(h0, p && mu && sigma && out, "mcp_portfolios_multi: NULL argument");
    if (n == 1) {
        const int rc = mcp_portfolios(h0, p, mu, sigma, out);
        if (kernel_ms_per_device) kernel_ms_per_device[0] = out->kernel_ms;
        return rc;
    }
    const bool any_array = p->weights_in || p->weights_recheck || out->weights || out->returns || out->risks || out->sharpes || out->accepted;
    MCP_REQUIRE(h0, p->space == MCP_HOST || !any_array, "mcp_portfolios_multi: arrays must be HOST space (each device fills its slice of the caller's buffers)");
    MCP_REQUIRE(h0, p->n_assets >= 1 && p->n_assets <= 4096, "mcp_portfolios_multi: n_assets=%d out of range [1, 4096]", p->n_assets);
    MCP_REQUIRE(h0, p->dtype == MCP_F32 || p->dtype == MCP_F64, "mcp_portfolios_multi: bad dtype %d", p->dtype);
    const size_t N = (size_t)p->n_assets, es = p->dtype == MCP_F64 ? 8 : 4;
    const int K = p->n_bins > 0 ? p->n_bins : 0;
    std::vector<mcp_portfolio_params> pp(n, *p);
    std::vector<mcp_portfolio_out> oo(n, *out);
    std::vector<std::vector<double>> wsel(n), bret(n);
    std::vector<std::vector<uint64_t>> bidx(n);
    auto at = [](const void* q, uint64_t rows, size_t row_bytes) -> void* { return q ? (unsigned char*)q + rows * row_bytes : nullptr; };
    for (int r = 0; r < n; ++r) {
        const Shard s = shard_of(p->n_portfolios, r, n);
        pp[r].n_portfolios = s.count;
        pp[r].first_index = p->first_index + s.first;
        pp[r].comm_merge = 1;
        pp[r].weights_in = at(p->weights_in, s.first, N * es);
        pp[r].weights_recheck = (const double*)at(p->weights_recheck, s.first, N * 8);
        oo[r].weights = at(out->weights, s.first, N * es);
        oo[r].returns = at(out->returns, s.first, es);
        oo[r].risks = at(out->risks, s.first, es);
        oo[r].sharpes = at(out->sharpes, s.first, es);
        oo[r].accepted = (uint8_t*)at(out->accepted, s.first, 1);
        if (r > 0) {                       // every rank returns the merged picks / bins: only rank 0 writes the caller's
            wsel[r].resize(2 * N);
            oo[r].max_sharpe.weights = wsel[r].data();
            oo[r].target_risk.weights = wsel[r].data() + N;
            if (K > 0) {
                bret[r].resize(K);
                bidx[r].resize(K);
                oo[r].bin_best_return = bret[r].data();
                oo[r].bin_best_index = bidx[r].data();
            }
        }
    }
    const int rc = run_ranks(hs, n, [&](int r) { return mcp_portfolios(hs[r], &pp[r], mu, sigma, &oo[r]); });
    if (rc != MCP_OK) return rc;
    uint64_t acc = 0;
    double ms = 0;
    for (int r = 0; r < n; ++r) {
        acc += oo[r].n_accepted;
        ms = oo[r].kernel_ms > ms ? oo[r].kernel_ms : ms;
        if (kernel_ms_per_device) kernel_ms_per_device[r] = oo[r].kernel_ms;
    }
    mcp_portfolio_out res = oo[0];
    res.weights = out->weights; res.returns = out->returns; res.risks = out->risks; res.sharpes = out->sharpes; res.accepted = out->accepted;
    res.n_accepted = acc;                  // the whole job's (= n_accepted_global)
    res.kernel_ms = ms;
    *out = res;
    return MCP_OK;
}

static int paths_stats_multi_impl(mcp_handle* hs, int n, const mcp_path_params* p, const double* mu, const double* sigma, const double* weights,
                                  mcp_path_stats* stats) {
    MCP_CHECK(check_group(hs, n, "mcp_paths_stats_multi"));
    mcp_handle h0 = hs[0];
    MCP_REQUIRE(h0, p && mu && sigma && weights && stats, "mcp_paths_stats_multi: NULL argument");
    MCP_REQUIRE(h0, p->normals_in == nullptr, "mcp_paths_stats_multi: supplied normals are a single-device (parity) mode");
    if (n == 1) return mcp_paths_stats(h0, p, mu, sigma, weights, nullptr, stats);
    std::vector<mcp_path_params> pp(n, *p);
    std::vector<mcp_path_stats> ss(n, *stats);
    for (int r = 0; r < n; ++r) {
        const Shard s = shard_of(p->n_paths, r, n);
        pp[r].n_paths = s.count;
        pp[r].first_index = p->first_index + s.first;
        pp[r].space = MCP_DEVICE;          // no arrays cross the boundary: the terminal values stay in library scratch
        ss[r].comm_merge = 1;
        ss[r].n_total = p->n_paths;
    }
    const int rc = run_ranks(hs, n, [&](int r) { return mcp_paths_stats(hs[r], &pp[r], mu, sigma, weights, nullptr, &ss[r]); });
    if (rc != MCP_OK) return rc;
    mcp_path_stats res = ss[0];
    for (int r = 1; r < n; ++r) {
        res.kernel_ms = ss[r].kernel_ms > res.kernel_ms ? ss[r].kernel_ms : res.kernel_ms;
        res.quantile_ms = ss[r].quantile_ms > res.quantile_ms ? ss[r].quantile_ms : res.quantile_ms;
    }
    res.comm_merge = stats->comm_merge;
    res.n_total = stats->n_total;
    *stats = res;
    return MCP_OK;
}

extern "C" int mcp_portfolios_multi(mcp_handle* handles, int n, const mcp_portfolio_params* params, const double* mu_host,
                                    const double* sigma_host, mcp_portfolio_out* out, double* kernel_ms_per_device) {
    if (!handles || n < 1 || !handles[0]) return MCP_ERR_INVALID;
    return mcp_guarded(handles[0], "mcp_portfolios_multi",
                       [&] { return portfolios_multi_impl(handles, n, params, mu_host, sigma_host, out, kernel_ms_per_device); });
}

extern "C" int mcp_paths_stats_multi(mcp_handle* handles, int n, const mcp_path_params* params, const double* mu_host,
                                     const double* sigma_host, const double* weights_host, mcp_path_stats* stats) {
    if (!handles || n < 1 || !handles[0]) return MCP_ERR_INVALID;
    return mcp_guarded(handles[0], "mcp_paths_stats_multi",
                       [&] { return paths_stats_multi_impl(handles, n, params, mu_host, sigma_host, weights_host, stats); });
}
