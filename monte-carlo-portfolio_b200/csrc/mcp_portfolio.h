// Internal (non-ABI) declarations shared by the portfolio-sweep translation units.
#pragma once
#include "mcp_context.h"

namespace mcp {

constexpr int PF_BLOCK = 256;          // threads per CTA of the sweep kernels
constexpr int PF_SMALL_MAX_N = 32;     // thread-per-portfolio register kernels up to this N

// Per-CTA (and final) selection candidates.  Keys are widened to double so that one
// reduction kernel serves both arithmetic types (float -> double is exact).
struct PfCand {
    double key_s;      // Sharpe
    uint64_t idx_s;    // global index, MCP_NO_INDEX if none
    double key_d;      // -|risk - target|
    uint64_t idx_d;
    double rmin, rmax; // risk range over accepted portfolios
};

// Selected-portfolio record written by the replay kernels: rec[0] = index bits,
// rec[1..4] = key, ret, risk, sharpe, rec[5 .. 5+N) = weights.
constexpr int PF_REC_HEADER = 5;

// One sweep launch over a contiguous global index range.  All pointers are device pointers.
struct PfJob {
    int n = 0;
    int dtype = MCP_F32;
    int max_tries = 100;
    int keep_last = 0;
    bool bounds = false;
    uint64_t seed = 0, first = 0, P = 0;
    double rf = 0, target = 0.30;
    const double* sigma = nullptr;     // host N x N
    const double* mu = nullptr;        // host N
    const double* lo = nullptr;        // host N or null
    const double* hi = nullptr;
    const void* w_in = nullptr;
    void* w_out = nullptr;
    void* ret_out = nullptr;
    void* risk_out = nullptr;
    void* sharpe_out = nullptr;
    uint8_t* acc_out = nullptr;
    PfCand* cands = nullptr;           // [max_blocks] per-CTA candidates
    int max_blocks = 0;
    unsigned long long* n_accepted = nullptr;
    // envelope (packed (return key << 32 | ~local index) per bin, FP32) -- see mcp_envelope
    int n_bins = 0;
    double env_lo = 0, env_hi = 0;
    unsigned long long* env_bins = nullptr;
    cudaStream_t stream = nullptr;
    int blocks_used = 0;               // out: grid size of the launch
    uint64_t tc_table_epoch = 0;       // != 0: the tcgen05 sweep's S' / mu table of THIS call sits in slot 6 under that epoch
    int tc_table_f16 = -1;             // operand split the table in slot 6 was built for (1: FP16 images, 0: TF32 + BF16)
    uint64_t lg_table_epoch = 0;       // != 0: the SIMT large-sweep constants (S' / Sigma, mu, bounds) of THIS call sit in slot 6 under that epoch
    int lg_table_kind = -1;            // which kernel family / dtype they were laid out for (0 tiled FP32, 1 generic FP32, 2 generic FP64)
    int rounds = 10;                   // Philox4x32 rounds (10 default, 7 optional)
    // list mode of the tiled SIMT sweep: evaluate the portfolios idx_list[0 .. *idx_count) (device memory) instead of a contiguous
    // range; outputs go to row (index - first).  Used for the rows the bounded tensor-core sweep defers.
    const uint64_t* idx_list = nullptr;
    const unsigned long long* idx_count = nullptr;
    const uint16_t* idx_attempt = nullptr;   // optional, per listed row: the attempt number of its first draw in this launch
    bool lgl_uploaded = false;         // the list-mode / bounded-route copy of the tiled kernel's constants (slot 20) is in place
    int tc_bounds_route = 0;           // bounded sweep at 32 < N <= 256: 1 once the tcgen05 route ran (the tiled SIMT kernel then keeps its
                                       // constants beside that kernel's table), 2 = the SIMT kernel only (replays)
};

// Replay of selected portfolios: regenerates (RNG mode) or re-reads (supplied mode, `rows`
// = device pointer to n_sel rows of N values) and evaluates them with the sweep's arithmetic.
struct PfReplay {
    int n_sel = 0;
    uint64_t idx[2] = {0, 0};
    const void* rows = nullptr;
    double* rec = nullptr;             // device, n_sel * (PF_REC_HEADER + N) doubles
};

// merges n per-CTA candidates (and the running record when accumulate != 0) into *acc
int pf_reduce_launch(mcp_context* h, const PfCand* cands, int n, PfCand* acc, int accumulate, cudaStream_t st);

// frontier envelope post-pass (mcp_envelope.cu): bins are (order-preserving return key, global index)
constexpr int ENV_MAX_BINS = 4096;
int env_reset(mcp_context* h, int K, unsigned long long* mx, unsigned long long* ix, cudaStream_t st);
int env_chunk(mcp_context* h, int dtype, const void* risk, const void* ret, uint64_t n, uint64_t base, double lo, double hi, int K,
              unsigned long long* cmax, unsigned long long* cidx, cudaStream_t st);
int env_fold(mcp_context* h, int K, const unsigned long long* cmax, const unsigned long long* cidx, unsigned long long* fmax,
             unsigned long long* fidx, cudaStream_t st);

// FP32 near-tie recheck (mcp_recheck.cu)
constexpr unsigned RC_CAP = 8192;
struct RcLists {
    unsigned int count[2];
    unsigned int pad[2];
    unsigned long long idx[2][RC_CAP];
    float key[2][RC_CAP];
};
int rc_collect_launch(mcp_context* h, const void* sharpe, const void* risk, uint64_t n, uint64_t base, const PfCand* running,
                      double rf_mu, double target, RcLists* lists, cudaStream_t st);
int rc_decide(mcp_context* h, const mcp_portfolio_params* p, const PfJob& job32, const PfCand& fin, double rf_mu,
              RcLists* d_lists, PfCand* d_cand_scratch, int max_blocks, unsigned long long* d_acc_scratch, uint64_t out_idx[2],
              int* overflow);
// overflow of the near-tie lists: the whole range re-evaluated in FP64 from weights_recheck (exact, slow, rare)
int rc_full_fp64(mcp_context* h, const mcp_portfolio_params* p, const PfJob& job32, PfCand* d_cand_scratch, int max_blocks,
                 unsigned long long* d_acc_scratch, uint64_t out_idx[2]);

int pf_small_launch(mcp_context* h, PfJob& job);
int pf_small_replay(mcp_context* h, const PfJob& job, const PfReplay& rp);
int pf_large_launch(mcp_context* h, PfJob& job);
int pf_large_replay(mcp_context* h, const PfJob& job, const PfReplay& rp);
// tcgen05 path of the large sweep (mcp_portfolio_large_tc.cu): FP32, 32 < N <= 256; Philox rows with or without bounds, supplied
// weights without bounds
bool pf_large_tc_eligible(const PfJob& job);
int pf_large_launch_tc(mcp_context* h, PfJob& job);
// Bounds rejection on the tensor-core sweep (app.py:700-707): the tcgen05 kernel evaluates every portfolio's attempt 0 and accepts
// the rows that are inside the bounds with a safety margin; the others (rejected, or too close to a bound to call in this kernel's
// summation order) are appended to a device list and go through the tiled SIMT kernel (pf_large_launch_list), which redraws with
// attempt + 1 exactly as before.  Accept / skip decisions and attempt numbers are therefore the SIMT kernel's.
int pf_large_launch_tc_bounded(mcp_context* h, PfJob& job);
int pf_large_launch_list(mcp_context* h, PfJob& job);            // tiled SIMT kernel over job.idx_list

// implemented per (type, padded N) in mcp_portfolio_small_inst.cu
template <typename T, int NP> int pf_small_launch_t(mcp_context* h, PfJob& job);
template <typename T, int NP> int pf_small_replay_t(mcp_context* h, const PfJob& job, const PfReplay& rp);

}  // namespace mcp
