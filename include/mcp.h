/*
 * mcp.h -- C ABI of libmcp.so: the B200 (sm_100a) Monte Carlo portfolio hot path.
 *
 * The reference (mohammadmarghzari/monte-carlo-portfolio, app.py) has no FFI or plugin
 * boundary: the path is inline script code plus one function-shaped twin.  Each entry
 * point below names the reference lines it replaces; the Python host
 * (monte-carlo-portfolio_b200/mcportfolio) binds exactly these symbols with ctypes and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: opaque handle, POD structs, pointers + sizes; no C++ / torch types.
 *   - every call returns MCP_OK (0) or a negative error class and never throws or exits;
 *     mcp_last_error(handle) returns the message of the last failure on that handle.
 *   - ownership: the caller allocates and frees every input / output buffer; the library
 *     owns only handle-internal scratch.  `space` says whether array arguments are host
 *     pointers (copies inside the call: page-locked buffers -- mcp_host_alloc -- are copied
 *     directly; pageable ones are staged through the handle's pinned double buffers, a host
 *     memcpy of chunk c overlapping the DMA of chunk c+1) or device pointers on the
 *     handle's device (no copies).
 *   - threading: one handle per device; a handle is not re-entrant.  Work is ordered on
 *     the handle's stream (mcp_set_stream) and every call returns after its results are
 *     complete (synchronous from the caller's view).
 *   - mu / Sigma / bounds / selection records always cross the boundary as host FP64,
 *     exactly the values the reference holds in `mean_returns` / `cov_matrix`
 *     (app.py:679-680); `dtype` selects the arithmetic and array element type on device.
 */
#ifndef MCP_H
#define MCP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCP_ABI_VERSION 2

#define MCP_OK               0
#define MCP_ERR_INVALID     (-1)  /* bad argument / unsupported shape          */
#define MCP_ERR_CUDA        (-2)  /* CUDA runtime failure (message has detail) */
#define MCP_ERR_NUMERIC     (-3)  /* Sigma not positive definite (Cholesky)    */
#define MCP_ERR_NOMEM       (-4)
#define MCP_ERR_COMM        (-5)  /* NCCL failure / no communicator (mcp_comm_*)   */

#define MCP_F32 0
#define MCP_F64 1

#define MCP_HOST   0
#define MCP_DEVICE 1

#define MCP_NO_INDEX UINT64_MAX
#define MCP_MAX_ALPHAS 8

typedef struct mcp_context* mcp_handle;

/* ---- lifecycle ------------------------------------------------------------------- */
int         mcp_abi_version(void);
int         mcp_create(int device, mcp_handle* out);
int         mcp_destroy(mcp_handle h);
const char* mcp_last_error(mcp_handle h);            /* h == NULL: last mcp_create failure */
/* NULL selects the handle's own stream, which is NON-BLOCKING: it does not synchronise with the legacy default stream
 * (handle 0).  To order the library's work behind a caller that uses the legacy default stream (torch's default), pass
 * cudaStreamLegacy ((void*)0x1), not 0.                                                                                */
int         mcp_set_stream(mcp_handle h, void* cuda_stream);
int         mcp_synchronize(mcp_handle h);
/* pinned host memory for the HOST-space fast path (plain malloc'ed buffers also work) */
int         mcp_host_alloc(size_t bytes, void** out);
int         mcp_host_free(void* p);

typedef struct {
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int32_t max_smem_per_block;
    uint64_t total_mem;
    char    name[64];
} mcp_device_info_t;
int mcp_device_info(mcp_handle h, mcp_device_info_t* out);

/* kernels launched by this handle since creation, and device time (CUDA events on the
 * handle's stream) of the dominant kernel(s) of the most recent call                    */
uint64_t mcp_launch_count(mcp_handle h);
double   mcp_last_kernel_ms(mcp_handle h);

/* ---- portfolio sweep ---------------------------------------------------------------
 * Replaces the sampling + evaluation loop app.py:699-717, the array materialisation
 * 719-722 and the selection 672/738/747 (twin: efficient_frontier, app.py:265-284).     */
typedef struct {
    int32_t  n_assets;          /* N >= 1                                                  */
    int32_t  dtype;             /* MCP_F32 | MCP_F64                                       */
    uint64_t n_portfolios;      /* P: units evaluated by this call (this rank's shard)     */
    uint64_t first_index;       /* global index of unit 0 (Philox counter, reported idx)   */
    uint64_t seed;              /* Philox key                                              */
    double   risk_free;         /* subtracted raw, app.py:711 (the app passes 3.0)         */
    double   risk_target;       /* nearest-risk pick target (README.md:4: 0.30)            */
    const double* min_weights;  /* host, N values, or NULL = no lower bounds (app.py:703)  */
    const double* max_weights;  /* host, N values, or NULL                                 */
    int32_t  max_tries;         /* rejection tries per portfolio (app.py:701: 100)         */
    int32_t  keep_last;         /* 1: keep the last draw on exhaustion (app.py:277)        */
    int32_t  space;             /* MCP_HOST | MCP_DEVICE for weights_in and output arrays  */
    int32_t  comm_merge;        /* 1: this call evaluates ONE RANK'S SHARD of a job; selections, counts, risk range and
                                   envelope bins are merged across the handle's communicator (mcp_comm_init) inside
                                   the call, so every rank returns the whole job's picks (arrays stay local)          */
    const void*   weights_in;      /* supplied-weights mode: [P, N] dtype; NULL = Philox   */
    const double* weights_recheck; /* optional FP64 copy of weights_in (same space) used to
                                      re-rank FP32 near-ties in FP64; NULL = off           */
    /* frontier envelope (replaces the scatter app.py:726-736 at scale); 0 bins = off      */
    int32_t  n_bins;
    int32_t  philox_rounds;     /* 0 or 10: Philox4x32-10 (the default, and the oracle's generator); 7: Philox4x32-7,
                                   Random123's documented Crush-resistant minimum (a different, cheaper stream)       */
    double   risk_lo, risk_hi;
} mcp_portfolio_params;

typedef struct {
    uint64_t index;             /* global index, MCP_NO_INDEX if nothing was accepted      */
    double   key;               /* Sharpe (max-Sharpe pick) or |risk - target| (risk pick) */
    double   ret, risk, sharpe;
    double*  weights;           /* host, N doubles, caller-allocated; may be NULL          */
} mcp_selection;

typedef struct {
    /* optional arrays in params.space, element type params.dtype; NULL = no write-back.
     * Rows of skipped portfolios (app.py:706-707) hold NaN metrics and accepted = 0.      */
    void*    weights;           /* [P, N]  all_weights  (app.py:721)                       */
    void*    returns;           /* [P]     all_returns  (app.py:720)                       */
    void*    risks;             /* [P]     all_risks    (app.py:719)                       */
    void*    sharpes;           /* [P]     all_metrics  (app.py:722, metric = sharpe)      */
    uint8_t* accepted;          /* [P]                                                      */
    /* envelope outputs (host), n_bins entries each, when params.n_bins > 0                */
    double*   bin_best_return;
    uint64_t* bin_best_index;
    /* results */
    uint64_t n_accepted;
    double   risk_min, risk_max;   /* over accepted portfolios                              */
    mcp_selection max_sharpe;      /* np.argmax(sharpe), first occurrence (app.py:672)      */
    mcp_selection target_risk;     /* argmin |risk - target|, first occurrence              */
    double   kernel_ms;            /* device time of the sweep kernel(s)                    */
    uint64_t n_accepted_global;    /* = n_accepted, or the sum over ranks with comm_merge    */
    int32_t  recheck_overflow;     /* weights_recheck: 1 = more FP32 near-ties than the screen holds were found and the
                                      tie set was re-evaluated by a full FP64 pass instead (result still exact)      */
    int32_t  reserved;
} mcp_portfolio_out;

/* Kernel selection (internal, by shape): N <= 32 register kernels (FP32 packed FFMA2 / FP64);
 * 32 < N <= 256 FP32 without bounds: tcgen05 tensor-core sweep -- FP16 operand split for in-kernel RNG rows,
 * TF32 + BF16 split for supplied weights (environment MCP_LARGE_TC=0 forces the SIMT kernel, MCP_LARGE_TC_F16=0
 * the TF32 split, for A/B measurements); bounds rejection at N > 32: tiled SIMT kernel; FP64 N > 32 and
 * FP32 N > 256: warp-per-portfolio kernel.  All produce the same portfolios for a given (seed, index).   */
int mcp_portfolios(mcp_handle h, const mcp_portfolio_params* params,
                   const double* mu_host, const double* sigma_host,
                   mcp_portfolio_out* out);

/* ---- correlated-path simulator -------------------------------------------------------
 * Not in the reference.  Spec: L = chol(Sigma); per step
 * r = mu dt + sqrt(dt) L z; V *= (1 + r) per asset (app.py:253 compounding);
 * terminal[m] = w . V_T - 1.                                                            */
typedef struct {
    int32_t  n_assets;
    int32_t  dtype;
    uint64_t n_paths;
    uint64_t first_index;
    uint64_t seed;
    int32_t  n_steps;
    int32_t  space;             /* of normals_in and terminal_out                          */
    double   dt;
    const void* normals_in;     /* supplied-normals mode: [M, S, N] dtype; NULL = Philox   */
    int32_t  philox_rounds;     /* 0 or 10 (default) | 7, as in mcp_portfolio_params       */
    int32_t  reserved;
} mcp_path_params;

/* Kernel selection (internal): FP32 Philox paths with N <= 256 run on the tensor cores (tcgen05: the step's normals go to
 * tensor memory as the A operand, L' sqrt(dt) is the B operand in shared memory, split operands with FP32 accumulation:
 * TF32 for N <= 32; for N > 32 two stages of FP16 pairs, the normals drawn and stored 16 at a time, and for N > 128 two
 * blocks of 128 assets; MCP_PATHS_TC=0 forces the SIMT kernels, MCP_PATHS_TC_WIDE16=0 the one-stage TF32 kernel above
 * N = 32, for A/B measurements); supplied normals, FP64 and wider universes use the thread-per-path register kernels
 * (N <= 32) or the warp-per-path kernel (N <= 1024).                                                                   */
int mcp_paths(mcp_handle h, const mcp_path_params* params,
              const double* mu_host, const double* sigma_host, const double* weights_host,
              void* terminal_out, double* kernel_ms);

/* Paths and their VaR / CVaR in ONE call (C4): the path kernel also fills the first radix histogram of the terminal
 * values, the remaining select passes, the interpolation and the tail sums run back to back on the handle's stream and
 * the results come back with a single copy.  terminal_out may be NULL (the values then live in library scratch).
 * comm_merge = 1: `params` describes this rank's shard, n_total is the whole job's path count, and histograms / tail
 * sums are all-reduced over the handle's communicator inside the call (exact global order statistics on every rank). */
typedef struct {
    int32_t  n_alphas;          /* 1..MCP_MAX_ALPHAS                                         */
    int32_t  comm_merge;
    uint64_t n_total;           /* global path count (ignored unless comm_merge)             */
    double   alphas[MCP_MAX_ALPHAS];
    double   var[MCP_MAX_ALPHAS];    /* out */
    double   cvar[MCP_MAX_ALPHAS];   /* out */
    double   kernel_ms;              /* out: path kernel                                     */
    double   quantile_ms;            /* out: select + tail passes (incl. their all-reduces)  */
} mcp_path_stats;
int mcp_paths_stats(mcp_handle h, const mcp_path_params* params,
                    const double* mu_host, const double* sigma_host, const double* weights_host,
                    void* terminal_out, mcp_path_stats* stats);

/* ---- VaR / CVaR ------------------------------------------------------------------------
 * app.py:258-263 conventions: VaR = np.percentile(x, (1-alpha)*100) (linear), CVaR =
 * mean(x[x <= VaR]) (VaR if empty), on exact order statistics (radix select).
 * `allreduce` (may be NULL) is called between radix passes with a DEVICE buffer of
 * `count` uint64 (kind 0) or double (kind 1) values that must be summed in place across
 * ranks; `n_total` is the global element count (= n when allreduce is NULL).             */
typedef int (*mcp_allreduce_fn)(void* device_buffer, size_t count, int kind, void* user);
/* pass this as `allreduce` to sum over the handle's own communicator (mcp_comm_init): NCCL all-reduces issued by the
 * library on the handle's stream, always stream-ordered                                                              */
#define MCP_ALLREDUCE_COMM ((mcp_allreduce_fn)(uintptr_t)1)
/* By default the callback is SYNCHRONOUS: libmcp drains its stream before calling it and expects the sums to be
 * in place when it returns.  With mcp_set_allreduce_stream_ordered(h, 1) the caller promises that the callback only
 * ENQUEUES the reduction on the handle's stream (mcp_set_stream; e.g. an NCCL all-reduce on that stream): the radix
 * passes, their all-reduces and the device-side digit selection then run back to back without host round trips.     */
int mcp_set_allreduce_stream_ordered(mcp_handle h, int on);

int mcp_quantiles(mcp_handle h, const void* values, int space, int dtype, uint64_t n,
                  uint64_t n_total, const double* alphas, int n_alphas,
                  double* var_out, double* cvar_out,
                  mcp_allreduce_fn allreduce, void* user);

/* Host-side state machine of the distributed exact radix select that mcp_quantiles runs
 * internally (no CUDA calls; usable for CPU tests of the multi-rank merge).  Per pass: every
 * rank histograms the next `mcp_select_pass_bits` key bits of its local values for each slot
 * prefix, the histograms are summed across ranks, and mcp_select_advance consumes the sum.  */
#define MCP_MAX_TARGETS 16
typedef struct {
    int32_t  key_bits;                       /* 32 (float) | 64 (double)                     */
    int32_t  bits_done;
    int32_t  n_targets;
    int32_t  n_slots;                        /* distinct prefixes in the coming pass         */
    uint64_t rank[MCP_MAX_TARGETS];          /* remaining 0-based rank inside the prefix     */
    uint64_t prefix[MCP_MAX_TARGETS];        /* decided high key bits, right-aligned         */
    int32_t  slot_of[MCP_MAX_TARGETS];
    uint64_t slot_prefix[MCP_MAX_TARGETS];
} mcp_select_state;

int mcp_select_init(mcp_select_state* s, int key_bits, const uint64_t* ranks, int n_targets);
int mcp_select_pass_bits(const mcp_select_state* s);          /* 0 when finished            */
int mcp_select_advance(mcp_select_state* s, const uint64_t* hist /* [n_slots][1 << bits] */);
/* one local radix pass on the device: hist_dev[n_slots << bits] (uint64) is overwritten     */
int mcp_select_hist(mcp_handle h, const void* values_dev, int dtype, uint64_t n,
                    const mcp_select_state* s, uint64_t* hist_dev);
/* order-preserving key <-> value (bit patterns), so callers can turn final prefixes into values */
double mcp_key_to_value(uint64_t key, int dtype);

/* ---- per-portfolio historical VaR / CVaR (app.py:710-713) and the 'VaR' / 'CVaR' method
 * selections (app.py:673-674, 717, 747)                                                  */
typedef struct {
    int32_t  n_assets;
    int32_t  n_periods;         /* T rows of the returns matrix                            */
    int32_t  dtype;
    int32_t  space;
    uint64_t n_portfolios;
    uint64_t first_index;
    double   alpha;             /* 0.95 (app.py:684)                                       */
    const void* weights_in;     /* [P, N] dtype                                            */
    int32_t  negate;            /* 1: the arrays hold -var / -cvar, the app's metric (app.py:717)                     */
    int32_t  recheck;           /* FP32 only, 1: portfolios whose FP32 VaR (CVaR) is within rounding of the best are
                                   re-evaluated in FP64 (series, order statistics, tail mean) and the pick is the
                                   FP64 argmax with the lowest index -- the index np.argmin(-var) returns on the
                                   reference's FP64 values (app.py:673-674, 747)                                      */
} mcp_hist_params;

typedef struct {
    void*    var;               /* [P] dtype, optional                                     */
    void*    cvar;              /* [P] dtype, optional                                     */
    uint64_t best_var_index;    /* argmin(-var)  = first index of the largest VaR          */
    uint64_t best_cvar_index;   /* argmin(-cvar)                                           */
    double   best_var, best_cvar;
    double   kernel_ms;
} mcp_hist_out;

int mcp_historical_var(mcp_handle h, const mcp_hist_params* params,
                       const double* returns_matrix_host /* [T, N] */, mcp_hist_out* out);

/* ---- per-asset statistics (app.py:231-263 as combined by calc_asset_stats, 286-335) ---------
 * returns_host: [T, N] FP64 row-major periodic returns.  stats_out: [N][MCP_STATS_FIELDS] FP64:
 * 0 sharpe, 1 sortino, 2 volatility_ann, 3 total_return_ann, 4 mean_ann, 5 mean_period,
 * 6 std_period (ddof=1), 7 min_period, 8 max_period, 9 max_drawdown, 10 var, 11 cvar.          */
#define MCP_STATS_FIELDS 12
int mcp_asset_stats(mcp_handle h, const double* returns_host, int n_periods, int n_assets, double risk_free,
                    double annual_factor, double alpha, double* stats_out);

/* ---- mu / Sigma estimation (app.py:679-680): mu = mean(R) * A, Sigma = cov(R, ddof = 1) * A ----------------------------
 * returns_host: [T, N] FP64 row-major (the app's returns_df, leading fillna(0) row included); mu_out: N, sigma_out: N x N
 * (symmetric), host FP64.  FP64 on the device: one CTA per (i, j >= i) column pair, two-pass (means first).  T = 1 gives
 * NaN covariances, as pandas does.                                                                                      */
int mcp_moments(mcp_handle h, const double* returns_host, int n_periods, int n_assets, double annual_factor,
                double* mu_out, double* sigma_out);

/* ---- frontier envelope of metrics that are already on the device ------------------------
 * Bins n (risk, return) pairs (DEVICE arrays of `dtype`; NaN rows = skipped portfolios are ignored)
 * exactly like mcp_portfolios does with n_bins > 0: K equal risk bins over [risk_lo, risk_hi], per bin
 * the largest return and the first global index (first_index + row) attaining it.  Lets a caller sweep
 * ONCE with the risk / return arrays kept in HBM (8 bytes per portfolio), read the attained risk range
 * from that sweep (all-reduce it across ranks), and bin afterwards -- instead of sweeping twice
 * (replaces the scatter of app.py:726-736).
 * bin_best_return / bin_best_index: host, n_bins entries; empty bins get -inf / MCP_NO_INDEX.          */
int mcp_envelope_arrays(mcp_handle h, int dtype, const void* risks_dev, const void* returns_dev, uint64_t n,
                        uint64_t first_index, double risk_lo, double risk_hi, int n_bins,
                        double* bin_best_return, uint64_t* bin_best_index);

/* ---- multi-GPU: one NCCL communicator per handle (one handle per GPU) --------------------------------------------------
 * The path shards by index range (first_index / n_portfolios, first_index / n_paths); the Philox counter is the GLOBAL
 * index, so the union over ranks is the same set of portfolios / paths for any rank count.  Only results cross NVLink:
 * selection records (all-gather), radix-select histograms and tail sums (all-reduce), risk ranges, envelope bins -- all
 * issued by the library on the handle's stream when an entry point is called with comm_merge = 1 / MCP_ALLREDUCE_COMM.
 * NCCL (libnccl.so.2) is bound at run time; a process that never calls mcp_comm_* does not need it.
 *   one process per GPU : rank 0 calls mcp_comm_unique_id and ships the 128 bytes to the others (any transport), then
 *                         every rank calls mcp_comm_init(handle, id, rank, nranks)        [ncclCommInitRank]
 *   one process, n GPUs : mcp_comm_init_all(handles, n); then drive each handle from its own host thread
 * After every host wait that follows a collective the library checks ncclCommGetAsyncError and returns MCP_ERR_COMM.   */
#define MCP_COMM_ID_BYTES 128
int mcp_comm_unique_id(void* id_out /* MCP_COMM_ID_BYTES */);
int mcp_comm_init(mcp_handle h, const void* id, int rank, int nranks);
int mcp_comm_init_all(mcp_handle* handles, int n);
int mcp_comm_destroy(mcp_handle h);
int mcp_comm_info(mcp_handle h, int* rank, int* nranks);          /* nranks = 0: no communicator */
int mcp_comm_nccl_version(int* version);
/* small host-buffer collectives for callers that merge their own results (e.g. envelope bins): every rank contributes
 * `bytes` and receives nranks * bytes in rank order / reduces `count` 8-byte elements in place                          */
#define MCP_REDUCE_U64_SUM 0
#define MCP_REDUCE_F64_SUM 1
#define MCP_REDUCE_F64_MIN 2
#define MCP_REDUCE_F64_MAX 3
#define MCP_REDUCE_U64_MAX 4
int mcp_comm_allgather(mcp_handle h, const void* send_host, size_t bytes, void* recv_host);
int mcp_comm_allreduce(mcp_handle h, void* inout_host, size_t count, int kind);

/* ---- one call, n GPUs of this process (the reference is ONE Streamlit process, Procfile:1) -------------------------------------
 * `handles`: n handles, one per GPU, joined by mcp_comm_init_all (handle r = rank r).  The job described by `params` -- the global
 * index range [first_index, first_index + n_portfolios) or n_paths -- is cut into n contiguous blocks (block r holds
 * total / n + (r < total % n) units) and every handle runs the ordinary entry point on its block from its own host thread inside
 * the library, with comm_merge set: `out` / `stats` receive the whole job's results, identical to a one-GPU run of the same job.
 * mcp_portfolios_multi: arrays (weights_in, weights_recheck, every array of `out`) must be HOST space and describe the WHOLE job;
 * each device reads / fills its slice.  out->n_accepted is the whole job's count, out->kernel_ms the slowest device's;
 * kernel_ms_per_device (n doubles) may be NULL.  mcp_paths_stats_multi: Philox paths only, terminal values stay on the devices.
 * On failure the message of the failing device is available through mcp_last_error(handles[0]).                                 */
int mcp_portfolios_multi(mcp_handle* handles, int n, const mcp_portfolio_params* params,
                         const double* mu_host, const double* sigma_host, mcp_portfolio_out* out, double* kernel_ms_per_device);
int mcp_paths_stats_multi(mcp_handle* handles, int n, const mcp_path_params* params,
                          const double* mu_host, const double* sigma_host, const double* weights_host, mcp_path_stats* stats);

/* ---- microbenchmarks used as roofline denominators (bench.py) ------------------------- */
int mcp_measure_fma_peak(mcp_handle h, int dtype /* MCP_F32 | MCP_F64 | 2 = packed FP32x2 (FFMA2) */, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MCP_H */
