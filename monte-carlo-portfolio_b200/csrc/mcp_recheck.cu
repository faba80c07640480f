// FP32 near-tie recheck: "same selected index" as the FP64 reference in supplied-weights mode.
//
// In FP32 two portfolios whose Sharpe ratios differ by less than the rounding error can swap
// rank, so np.argmax on the reference's FP64 values may pick a different row than the FP32
// sweep.  When the caller also supplies the FP64 weights (mcp_portfolio_params.weights_recheck)
// the FP32 sweep is used as a SCREEN: every portfolio whose FP32 key is within a rounding
// tolerance of the (running) best is recorded by a post-pass over the chunk's Sharpe / risk
// arrays, and at the end those few rows are re-evaluated in FP64 (the same sweep kernels,
// dtype double, on the gathered rows sorted by index) to pick the winner.  All arithmetic stays
// on the device; the host only filters a list of <= RC_CAP (index, key) pairs.
#include <algorithm>
#include <cmath>
#include <vector>

#include "mcp_device.cuh"
#include "mcp_portfolio.h"

namespace mcp {

__host__ __device__ inline double rc_tol_sharpe(double best, double rmin, double rf_mu) {
    const double eps = 5.96e-8 * 64;
    const double r = rmin > 1e-30 ? rmin : 1e-30;
    return eps * (fabs(best) + rf_mu / r);
}
__host__ __device__ inline double rc_tol_dist(double best_neg_dist, double target) {
    const double eps = 5.96e-8 * 64;
    return eps * (fabs(target) + fabs(best_neg_dist));
}

__global__ void __launch_bounds__(256) rc_collect(const float* __restrict__ sharpe, const float* __restrict__ risk, uint64_t n,
                                                  uint64_t base, const PfCand* __restrict__ running, double rf_mu, double target,
                                                  RcLists* lists) {
    const PfCand run = *running;
    if (run.idx_s == MCP_NO_INDEX) return;
    // The running best can be off by its own rounding error (bounded with the running minimum risk, which is <= the risk of
    // the running best); a candidate by ITS error, which scales with 1 / its own risk -- so a later chunk that lowers the
    // minimum risk cannot invalidate what an earlier chunk decided.  The recorded key is the candidate's upper bound s + e.
    const float thr_s = (float)(run.key_s - rc_tol_sharpe(run.key_s, run.rmin, rf_mu));
    const float thr_d = (float)(run.key_d - rc_tol_dist(run.key_d, target));
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const float s = sharpe[i], k = risk[i];
        const float su = s + (float)rc_tol_sharpe((double)s, (double)k, rf_mu);
        if (su >= thr_s) {
            const unsigned slot = atomicAdd(&lists->count[0], 1u);
            if (slot < RC_CAP) { lists->idx[0][slot] = base + i; lists->key[0][slot] = su; }
        }
        const float d = -fabsf(k - (float)target);
        if (d >= thr_d) {
            const unsigned slot = atomicAdd(&lists->count[1], 1u);
            if (slot < RC_CAP) { lists->idx[1][slot] = base + i; lists->key[1][slot] = d; }
        }
    }
}

int rc_collect_launch(mcp_context* h, const void* sharpe, const void* risk, uint64_t n, uint64_t base, const PfCand* running,
                      double rf_mu, double target, RcLists* lists, cudaStream_t st) {
    uint64_t g = (n + 255) / 256;
    const uint64_t cap = (uint64_t)h->prop.multiProcessorCount * 8;
    rc_collect<<<(unsigned)std::max<uint64_t>(1, std::min(g, cap)), 256, 0, st>>>((const float*)sharpe, (const float*)risk, n, base,
                                                                                   running, rf_mu, target, lists);
    MCP_CUDA(h, cudaGetLastError());
    h->launches++;
    return MCP_OK;
}

// Final decision.  `fin` = FP32 winners; returns the FP64-rechecked global indices in out_idx.
int rc_decide(mcp_context* h, const mcp_portfolio_params* p, const PfJob& job32, const PfCand& fin, double rf_mu,
              RcLists* d_lists, PfCand* d_cand_scratch, int max_blocks, unsigned long long* d_acc_scratch, uint64_t out_idx[2],
              int* overflow) {
    cudaStream_t st = job32.stream;
    static_assert(sizeof(RcLists) < (1 << 20), "candidate lists are meant to be small");
    std::vector<unsigned char> raw(sizeof(RcLists));
    MCP_CUDA(h, cudaMemcpyAsync(raw.data(), d_lists, sizeof(RcLists), cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));
    const RcLists* L = reinterpret_cast<const RcLists*>(raw.data());
    const int N = p->n_assets;
    out_idx[0] = fin.idx_s;
    out_idx[1] = fin.idx_d;
    *overflow = 0;
    const double thr[2] = {fin.key_s - rc_tol_sharpe(fin.key_s, fin.rmin, rf_mu), fin.key_d - rc_tol_dist(fin.key_d, p->risk_target)};
    for (int c = 0; c < 2; ++c) {
        if (L->count[c] > RC_CAP) { *overflow = 1; continue; }           // too many near-ties: keep the FP32 pick
        std::vector<uint64_t> idx;
        for (unsigned k = 0; k < L->count[c]; ++k)
            if ((double)L->key[c][k] >= (float)thr[c]) idx.push_back(L->idx[c][k]);
        std::sort(idx.begin(), idx.end());
        idx.erase(std::unique(idx.begin(), idx.end()), idx.end());
        if (idx.size() <= 1) continue;                                    // no near-tie: the FP32 pick stands
        // gather the FP64 rows (ascending index: first occurrence wins ties, app.py:672)
        double* d_rows = nullptr;
        MCP_CHECK(mcp_dev_reserve(h, 11, idx.size() * (size_t)N * 8, (void**)&d_rows));
        for (size_t k = 0; k < idx.size(); ++k) {
            const double* src = p->weights_recheck + (idx[k] - p->first_index) * (uint64_t)N;
            MCP_CUDA(h, cudaMemcpyAsync(d_rows + k * (size_t)N, src, (size_t)N * 8,
                                        p->space == MCP_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        }
        PfJob j = job32;
        j.dtype = MCP_F64;
        j.first = 0;
        j.P = idx.size();
        j.w_in = d_rows;
        j.w_out = j.ret_out = j.risk_out = j.sharpe_out = nullptr;
        j.acc_out = nullptr;
        j.cands = d_cand_scratch;
        j.max_blocks = max_blocks;
        j.n_accepted = d_acc_scratch;
        j.bounds = false;                                   // candidates were accepted by the FP32 pass
        j.lo = j.hi = nullptr;
        MCP_CHECK(j.n <= PF_SMALL_MAX_N ? pf_small_launch(h, j) : pf_large_launch(h, j));
        MCP_CHECK(pf_reduce_launch(h, d_cand_scratch, j.blocks_used, d_cand_scratch + max_blocks, 0, st));
        PfCand best;
        MCP_CUDA(h, cudaMemcpyAsync(&best, d_cand_scratch + max_blocks, sizeof best, cudaMemcpyDeviceToHost, st));
        MCP_CUDA(h, cudaStreamSynchronize(st));
        const uint64_t pos = c == 0 ? best.idx_s : best.idx_d;
        if (pos != MCP_NO_INDEX && pos < idx.size()) out_idx[c] = idx[pos];
    }
    return MCP_OK;
}

// Exact fallback when the near-tie lists overflowed: every row of weights_recheck evaluated in FP64 by the same sweep
// kernels (device space: one launch; host space: chunks through scratch slot 11), picks = the FP64 sweep's own.
int rc_full_fp64(mcp_context* h, const mcp_portfolio_params* p, const PfJob& job32, PfCand* d_cand_scratch, int max_blocks,
                 unsigned long long* d_acc_scratch, uint64_t out_idx[2]) {
    cudaStream_t st = job32.stream;
    const int N = p->n_assets;
    const uint64_t P = p->n_portfolios;
    PfJob j = job32;
    j.dtype = MCP_F64;
    j.w_out = j.ret_out = j.risk_out = j.sharpe_out = nullptr;
    j.acc_out = nullptr;
    j.cands = d_cand_scratch;
    j.max_blocks = max_blocks;
    j.n_accepted = d_acc_scratch;
    j.n_bins = 0;
    PfCand* running = d_cand_scratch + max_blocks;
    const uint64_t chunk = p->space == MCP_DEVICE ? P : std::max<uint64_t>(1024, ((uint64_t)32 << 20) / ((uint64_t)N * 8));
    int first = 1;
    for (uint64_t r0 = 0; r0 < P; r0 += chunk) {
        const uint64_t rows = std::min<uint64_t>(chunk, P - r0);
        const double* src = p->weights_recheck + r0 * (uint64_t)N;
        if (p->space == MCP_HOST) {
            double* d_rows = nullptr;
            MCP_CHECK(mcp_dev_reserve(h, 11, chunk * (size_t)N * 8, (void**)&d_rows));
            MCP_CUDA(h, cudaMemcpyAsync(d_rows, src, rows * (size_t)N * 8, cudaMemcpyHostToDevice, st));
            src = d_rows;
        }
        j.first = p->first_index + r0;
        j.P = rows;
        j.w_in = src;
        MCP_CHECK(j.n <= PF_SMALL_MAX_N ? pf_small_launch(h, j) : pf_large_launch(h, j));
        MCP_CHECK(pf_reduce_launch(h, d_cand_scratch, j.blocks_used, running, first ? 0 : 1, st));
        first = 0;
        if (p->space == MCP_HOST) MCP_CUDA(h, cudaStreamSynchronize(st));      // the scratch rows are reused by the next chunk
    }
    PfCand best;
    MCP_CUDA(h, cudaMemcpyAsync(&best, running, sizeof best, cudaMemcpyDeviceToHost, st));
    MCP_CUDA(h, cudaStreamSynchronize(st));
    if (best.idx_s != MCP_NO_INDEX) out_idx[0] = best.idx_s;
    if (best.idx_d != MCP_NO_INDEX) out_idx[1] = best.idx_d;
    return MCP_OK;
}

}  // namespace mcp
