"""GPU tier: the CUDA portfolio sweep (through the C ABI) against the oracle.

Tolerances (BASELINE.json north_star): supplied-weights mode within 1e-6 relative in FP64
and 1e-4 in FP32, with the same selected index; in-kernel RNG is checked value-by-value
against the numpy restatement of the generator and statistically against np.random.dirichlet.
"""
import numpy as np
import pytest

from conftest import load_golden, synthetic_inputs
from oracle import philox_np, reference_np as ref

pytestmark = pytest.mark.gpu

RTOL = {"float64": 1e-6, "float32": 1e-4}


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


def check_against_oracle(r, W, mu, sigma, rf, target, dtype, same_index=True):
    want = ref.evaluate(W, mu, sigma, rf, target)
    tol = RTOL[dtype]
    assert r.n_accepted == len(W)
    assert np.allclose(r.returns, want["returns"], rtol=tol, atol=tol * 1e-3)
    assert np.allclose(r.risks, want["risks"], rtol=tol)
    assert np.allclose(r.sharpes, want["sharpes"], rtol=tol, atol=tol * np.abs(want["sharpes"]).max())
    assert np.allclose(r.weights, W, rtol=tol, atol=1e-7)
    for pick in ("max_sharpe", "target_risk"):
        got, exp = getattr(r, pick), want[pick]
        if same_index:
            assert got["index"] == exp["index"], pick
        assert np.isclose(got["ret"], exp["ret"], rtol=tol) and np.isclose(got["risk"], exp["risk"], rtol=tol)
        assert np.allclose(got["weights"], exp["weights"], rtol=tol, atol=1e-7)
    return want


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("rf,opt", [(3.0, 1451), (0.03, 4593)])
def test_c1_supplied_weights_matches_reference_lines(mcp, c1, dtype, rf, opt):
    """C1: the 10k legacy-seed-42 Dirichlet draws, evaluated by the reference's own lines."""
    W = c1["weights"]
    r = mcp.simulate_portfolios(c1["mu"], c1["sigma"], len(W), weights=W, risk_free=rf, dtype=dtype)
    tol = RTOL[dtype]
    tag = "rf3" if rf == 3.0 else "rf003"
    assert np.allclose(r.risks, c1["risks"], rtol=tol)
    assert np.allclose(r.returns, c1["returns"], rtol=tol)
    assert np.allclose(r.sharpes, c1[f"sharpes_{tag}"], rtol=tol)
    assert r.max_sharpe["index"] == opt == int(c1[f"opt_sharpe_{tag}"])
    assert r.target_risk["index"] == int(c1["opt_target30"])
    check_against_oracle(r, W, c1["mu"], c1["sigma"], rf, 0.30, dtype)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_c2_14_assets_supplied(mcp, c2, dtype):
    W = c2["weights"]
    r = mcp.simulate_portfolios(c2["mu"], c2["sigma"], len(W), weights=W, risk_free=0.03, dtype=dtype)
    assert np.allclose(r.risks, c2["risks"], rtol=RTOL[dtype])
    assert np.allclose(r.sharpes, c2["sharpes"], rtol=RTOL[dtype], atol=RTOL[dtype])
    assert r.max_sharpe["index"] == int(c2["opt_sharpe"])


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 13, 16, 17, 24, 31, 32])
def test_supplied_weights_all_small_n(mcp, n, dtype):
    mu, sigma = synthetic_inputs(n, seed=n)
    rng = np.random.RandomState(n)
    P = 3001                                    # ragged: not a multiple of the CTA tile
    W = rng.dirichlet(np.ones(n), size=P)
    r = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03, risk_target=0.25, dtype=dtype)
    check_against_oracle(r, W, mu, sigma, 0.03, 0.25, dtype)


def test_edge_sizes(mcp):
    mu, sigma = synthetic_inputs(4)
    r = mcp.simulate_portfolios(mu, sigma, 0)
    assert r.n_accepted == 0 and r.max_sharpe is None and r.risks.shape == (0,)
    W = np.array([[0.25, 0.25, 0.25, 0.25]])
    r = mcp.simulate_portfolios(mu, sigma, 1, weights=W, dtype="float64")
    assert r.max_sharpe["index"] == 0 and r.target_risk["index"] == 0
    # zero-risk portfolio: sharpe = 0 (app.py:711 `if port_std > 0 else 0`)
    r = mcp.simulate_portfolios(mu, np.zeros((4, 4)), 1, weights=W, dtype="float64")
    assert r.risks[0] == 0.0 and r.sharpes[0] == 0.0


def test_first_occurrence_tie_break(mcp):
    """np.argmax / np.argmin return the first index among ties."""
    mu, sigma = synthetic_inputs(4)
    rng = np.random.RandomState(0)
    base = rng.dirichlet(np.ones(4), size=50)
    W = np.tile(base, (40, 1))                 # every row repeated 40 times -> exact ties
    for dtype in ("float64", "float32"):
        r = mcp.simulate_portfolios(mu, sigma, len(W), weights=W, dtype=dtype, risk_free=0.03)
        assert r.max_sharpe["index"] < 50 and r.target_risk["index"] < 50
        want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
        assert r.max_sharpe["index"] == want["max_sharpe"]["index"]
        assert r.target_risk["index"] == want["target_risk"]["index"]


@pytest.mark.parametrize("dtype,atol", [("float32", 2e-6), ("float64", 1e-12)])
@pytest.mark.parametrize("n", [2, 5, 16, 20, 32])
def test_rng_mode_matches_generator_restatement(mcp, n, dtype, atol):
    """In-kernel Philox weights == oracle/philox_np.py on the same counters; metrics follow."""
    mu, sigma = synthetic_inputs(n, seed=3)
    P, first, seed = 5000, 123_456_789_012, 0xDEADBEEFCAFE
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=seed, first_index=first, dtype=dtype)
    W, valid = philox_np.dirichlet_weights(first, P, n, seed, dtype)
    assert valid.all() and r.n_accepted == P
    # MUFU lg2 has ~2^-22 ABSOLUTE error near U = 1, so a weight is exact to ~2^-22 / sum(e)
    s = philox_np.exponentials(np.arange(first, first + P, dtype=np.uint64), 0, n, seed, dtype).sum(1, keepdims=True)
    assert np.all(np.abs(r.weights - W) <= atol + (6e-7 / s if dtype == "float32" else 0.0))
    tol = RTOL[dtype]
    want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    ok = (s[:, 0] > 0.05) if dtype == "float32" else np.ones(P, bool)     # see the lg2 note above
    assert ok.mean() > 0.99
    assert np.allclose(r.risks[ok], want["risks"][ok], rtol=tol) and np.allclose(r.returns[ok], want["returns"][ok], rtol=tol)
    assert np.allclose(r.sharpes[ok], want["sharpes"][ok], rtol=tol, atol=tol)
    # the kernel's metrics are exact (to tolerance) for the weights it actually produced
    own = ref.evaluate(np.asarray(r.weights, dtype=np.float64), mu, sigma, 0.03, 0.30)
    assert np.allclose(r.risks, own["risks"], rtol=tol) and np.allclose(r.sharpes, own["sharpes"], rtol=tol, atol=tol)
    # picks: the selected portfolio must be (within tolerance) the oracle's optimum
    assert np.isclose(r.max_sharpe["sharpe"], want["max_sharpe"]["sharpe"], rtol=tol)
    assert r.max_sharpe["global_index"] == first + r.max_sharpe["index"]
    if dtype == "float64":
        assert r.max_sharpe["index"] == want["max_sharpe"]["index"]
        assert r.target_risk["index"] == want["target_risk"]["index"]
    # the record is re-evaluated with the sweep's arithmetic: identical to the array entries
    i = r.max_sharpe["index"]
    assert r.max_sharpe["sharpe"] == float(r.sharpes[i]) and r.max_sharpe["risk"] == float(r.risks[i])


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_rng_bounds_rejection_and_skip(mcp, dtype):
    """app.py:700-707: <= max_tries draws, skip on exhaustion -> arrays shorter than P."""
    n = 4
    mu, sigma = synthetic_inputs(n)
    lo, hi = np.array([0.2, 0.0, 0.0, 0.0]), np.array([0.3, 1.0, 1.0, 0.5])
    P, seed = 4000, 11
    r = mcp.simulate_portfolios(mu, sigma, P, min_weights=lo, max_weights=hi, seed=seed, dtype=dtype,
                                max_tries=4, risk_free=0.03)
    W, valid = philox_np.dirichlet_weights(0, P, n, seed, dtype, lo, hi, max_tries=4)
    assert 0 < valid.sum() < P
    # acceptance can differ only where a weight sits on a bound within rounding
    acc = r.accepted.astype(bool)
    margin = np.minimum(np.abs(W - lo).min(1), np.abs(W - hi).min(1))
    differ = acc != valid
    assert differ.sum() <= 2 and np.all(margin[differ] < 1e-5)
    both = acc & valid
    assert np.allclose(r.weights[np.cumsum(acc)[both] - 1], W[both], atol=2e-6 if dtype == "float32" else 1e-12)
    assert np.all(r.weights >= lo - 1e-6) and np.all(r.weights <= hi + 1e-6)
    assert r.n_accepted == acc.sum() == len(r.risks)
    # picks index the accepted-only arrays, like the reference's opt_idx (app.py:747)
    assert r.max_sharpe["index"] == int(np.argmax(r.sharpes))
    # keep_last (efficient_frontier semantics, app.py:277): nothing is skipped
    r2 = mcp.simulate_portfolios(mu, sigma, P, min_weights=lo, max_weights=hi, seed=seed, dtype=dtype,
                                 max_tries=4, keep_last=True)
    assert r2.n_accepted == P and np.allclose(r2.weights, W, atol=2e-6)
    # impossible bounds: everything skipped, empty arrays, no pick
    r3 = mcp.simulate_portfolios(mu, sigma, 500, min_weights=np.full(n, 0.6), seed=1, dtype=dtype, max_tries=3)
    assert r3.n_accepted == 0 and r3.max_sharpe is None and len(r3.risks) == 0


def test_sharding_invariance(mcp, synth16):
    """Counter = global index: any split of the range gives the same portfolios and picks."""
    mu, sigma = synth16
    P = 300_000
    whole = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=5, return_arrays=False)
    cuts = [0, 70_001, 150_000, 299_999, P]
    parts = [mcp.simulate_portfolios(mu, sigma, b - a, risk_free=0.03, seed=5, first_index=a, return_arrays=False)
             for a, b in zip(cuts[:-1], cuts[1:])]
    best = max(parts, key=lambda r: (r.max_sharpe["key"], -r.max_sharpe["global_index"]))
    assert best.max_sharpe["global_index"] == whole.max_sharpe["global_index"]
    assert best.max_sharpe["sharpe"] == whole.max_sharpe["sharpe"]
    near = min(parts, key=lambda r: (r.target_risk["key"], r.target_risk["global_index"]))
    assert near.target_risk["global_index"] == whole.target_risk["global_index"]
    assert np.array_equal(near.target_risk["weights"], whole.target_risk["weights"])
    assert sum(p.n_accepted for p in parts) == whole.n_accepted == P
    # and the no-write-back run agrees with the run that materialises the arrays
    full = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=5)
    assert full.max_sharpe["global_index"] == whole.max_sharpe["global_index"] == int(np.argmax(full.sharpes))
    assert whole.target_risk["global_index"] == int(np.argmin(np.abs(full.risks - 0.30)))


@pytest.mark.parametrize("rounds", [10, 7])      # 7: the optional Philox4x32-7 stream must pass the same statistical tier
def test_rng_statistics_against_numpy_dirichlet(mcp, synth16, rounds):
    """Flat Dirichlet: mean 1/N, var (N-1)/(N^2 (N+1)); (risk, return) envelope vs numpy's sampler."""
    mu, sigma = synth16
    N, P = 16, 2_000_000
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=99, philox_rounds=rounds)
    W = r.weights
    assert np.allclose(W.sum(1), 1.0, atol=1e-5)
    v = (N - 1) / (N * N * (N + 1))
    assert np.allclose(W.mean(0), 1 / N, atol=5 * np.sqrt(v / P))
    assert np.allclose(W.var(0), v, rtol=0.01)
    c = np.cov(W.astype(np.float64), rowvar=False)                      # all 2e6 rows: s.e. of an entry ~ v / sqrt(P) = 2.4e-6
    off = c[~np.eye(N, dtype=bool)]
    assert np.allclose(off, -1 / (N * N * (N + 1)), atol=5 * v / np.sqrt(P))          # Dirichlet covariance, 5 sigma
    # adjacent assets share Philox words in the FP32 24-bit field packing: their dependence must be the Dirichlet one only
    adj = np.array([c[i, i + 1] for i in range(N - 1)])
    assert np.abs(adj + 1 / (N * N * (N + 1))).max() < 5 * v / np.sqrt(P)
    Wn = np.random.RandomState(1).dirichlet(np.ones(N), size=P)
    _, risk_n, sharpe_n = ref.portfolio_metrics(Wn, mu, sigma, 0.03)
    for q in (0.001, 0.01, 0.5, 0.99, 0.999):
        assert np.isclose(np.quantile(r.risks, q), np.quantile(risk_n, q), rtol=5e-3)
        assert np.isclose(np.quantile(r.sharpes, q), np.quantile(sharpe_n, q), rtol=1e-2, atol=1e-3)
    # a single coordinate is Beta(1, N-1): KS distance against the exact CDF
    x = np.sort(W[:500_000, 3].astype(np.float64))
    cdf = 1 - (1 - x) ** (N - 1)
    ks = np.abs(cdf - (np.arange(len(x)) + 0.5) / len(x)).max()
    assert ks < 1.63 / np.sqrt(len(x)) * 1.5


def test_efficient_frontier_drop_in(mcp, c1):
    """Same signature and return layout as app.py:265-284."""
    res, W = mcp.efficient_frontier(c1["mu"], c1["sigma"], points=200, seed=4)
    assert res.shape == (3, 200) and W.shape == (200, 2) and res.dtype == np.float64
    Wp, _ = philox_np.dirichlet_weights(0, 200, 2, 4, "float64")
    assert np.allclose(W, Wp, atol=1e-12)
    ret, risk, sh = ref.portfolio_metrics(W, c1["mu"], c1["sigma"], 0.0)
    assert np.allclose(res[0], risk, rtol=1e-9) and np.allclose(res[1], ret, rtol=1e-9)
    assert np.allclose(res[2], ret / risk, rtol=1e-9)                 # Sharpe without rf (app.py:282)
    import pandas as pd                                               # the app passes pandas objects
    res2, W2 = mcp.efficient_frontier(pd.Series(c1["mu"]), pd.DataFrame(c1["sigma"]), 200, seed=4)
    assert np.array_equal(res2, res) and np.array_equal(W2, W)
    # bounded: keeps the last draw on exhaustion, so all `points` rows come back
    res3, W3 = mcp.efficient_frontier(c1["mu"], c1["sigma"], 64, np.array([0.49, 0.0]), np.array([0.5, 1.0]), seed=4)
    assert W3.shape == (64, 2)
    inside = (W3[:, 0] >= 0.49) & (W3[:, 0] <= 0.5)
    assert 0 < inside.sum() < 64


def test_device_space_torch_tensors(mcp, synth16):
    import torch
    mu, sigma = synth16
    W = np.random.RandomState(2).dirichlet(np.ones(16), size=10_000)
    Wt = torch.from_numpy(W).cuda()
    r = mcp.simulate_portfolios(mu, sigma, len(W), weights=Wt, risk_free=0.03, dtype="float64")
    assert r.risks.is_cuda
    want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    assert np.allclose(r.risks.cpu().numpy(), want["risks"], rtol=1e-9)
    assert r.max_sharpe["index"] == want["max_sharpe"]["index"]
    r2 = mcp.simulate_portfolios(mu, sigma, 50_000, seed=1, return_arrays="device")
    assert r2.weights.is_cuda and r2.weights.shape == (50_000, 16)
    assert torch.allclose(r2.weights.sum(1), torch.ones(50_000, device="cuda"), atol=1e-5)


def test_host_pipeline_many_chunks(mcp):
    """HOST space with more rows than one pipeline slot holds: chunks, both slots, ragged tail."""
    n = 32
    mu, sigma = synthetic_inputs(n, seed=1)
    P = 2_000_003
    W = np.random.RandomState(3).dirichlet(np.ones(n), size=P).astype(np.float32)
    r = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03, dtype="float32")
    ret, risk, sharpe = ref.portfolio_metrics(W.astype(np.float64), mu, sigma, 0.03)
    assert np.allclose(r.risks, risk, rtol=1e-4) and np.allclose(r.sharpes, sharpe, rtol=1e-4, atol=1e-4)
    assert np.array_equal(r.weights, W)
    assert np.isclose(r.max_sharpe["sharpe"], sharpe.max(), rtol=1e-5)
    assert r.max_sharpe["index"] == int(np.argmax(r.sharpes))
    assert r.target_risk["index"] == int(np.argmin(np.abs(r.risks - np.float32(0.30))))


# ---- N > 32: tiled FP32 kernel (N <= 256) and the generic warp-per-portfolio kernel ----------

@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("n", [33, 48, 64, 100, 200, 256, 300])
def test_large_n_supplied_weights(mcp, n, dtype):
    mu, sigma = synthetic_inputs(n, seed=n)
    P = 1000 if n <= 256 else 300
    W = np.random.RandomState(n).dirichlet(np.ones(n), size=P)
    r = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03, risk_target=0.03, dtype=dtype)
    check_against_oracle(r, W, mu, sigma, 0.03, 0.03, dtype)


@pytest.mark.parametrize("dtype,atol", [("float32", 1e-6), ("float64", 1e-12)])
@pytest.mark.parametrize("n", [40, 256, 260])
def test_large_n_rng_matches_generator_restatement(mcp, n, dtype, atol):
    mu, sigma = synthetic_inputs(n, seed=1)
    P, first, seed = 777, 9_876_543_210, 42
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=seed, first_index=first, dtype=dtype)
    W, valid = philox_np.dirichlet_weights(first, P, n, seed, dtype)
    assert r.n_accepted == P and np.allclose(r.weights, W, atol=atol)
    tol = RTOL[dtype]
    want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    assert np.allclose(r.risks, want["risks"], rtol=tol) and np.allclose(r.sharpes, want["sharpes"], rtol=tol, atol=tol)
    assert np.isclose(r.max_sharpe["sharpe"], want["max_sharpe"]["sharpe"], rtol=tol)
    i = r.max_sharpe["index"]
    assert r.max_sharpe["sharpe"] == float(r.sharpes[i]) and r.max_sharpe["risk"] == float(r.risks[i])
    assert np.allclose(r.max_sharpe["weights"], r.weights[i], atol=0, rtol=0)
    j = r.target_risk["index"]
    assert j == int(np.argmin(np.abs(r.risks - r.risks.dtype.type(0.30))))
    if dtype == "float64":
        assert i == want["max_sharpe"]["index"] and j == want["target_risk"]["index"]


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_large_n_bounds_rejection(mcp, dtype):
    n = 64
    mu, sigma = synthetic_inputs(n)
    hi = np.full(n, 0.07)                      # max weight 7 %: a flat-Dirichlet draw at N=64 often violates it
    P, seed = 900, 3
    r = mcp.simulate_portfolios(mu, sigma, P, max_weights=hi, seed=seed, dtype=dtype, max_tries=3, risk_free=0.03)
    W, valid = philox_np.dirichlet_weights(0, P, n, seed, dtype, None, hi, max_tries=3)
    assert 0 < valid.sum() < P
    acc = r.accepted.astype(bool)
    margin = np.abs(W - hi).min(1)
    differ = acc != valid
    assert differ.sum() <= 2 and np.all(margin[differ] < 1e-5)
    both = acc & valid
    assert np.allclose(r.weights[np.cumsum(acc)[both] - 1], W[both], atol=1e-6 if dtype == "float32" else 1e-12)
    assert r.weights.max() <= 0.07 + 1e-6 and r.n_accepted == acc.sum()
    want = ref.evaluate(np.asarray(r.weights, dtype=np.float64), mu, sigma, 0.03, 0.30)
    assert np.allclose(r.sharpes, want["sharpes"], rtol=RTOL[dtype], atol=RTOL[dtype])
    assert r.max_sharpe["index"] == int(np.argmax(r.sharpes))


def test_large_n_sharding_invariance_and_tail_tiles(mcp):
    n = 256
    mu, sigma = synthetic_inputs(n)
    P = 20_011                                  # not a multiple of the 64-portfolio tile
    whole = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=8, return_arrays=False)
    cuts = [0, 63, 64, 10_000, P]
    parts = [mcp.simulate_portfolios(mu, sigma, b - a, risk_free=0.03, seed=8, first_index=a, return_arrays=False)
             for a, b in zip(cuts[:-1], cuts[1:])]
    best = max(parts, key=lambda r: (r.max_sharpe["key"], -r.max_sharpe["global_index"]))
    assert best.max_sharpe["global_index"] == whole.max_sharpe["global_index"]
    assert best.max_sharpe["sharpe"] == whole.max_sharpe["sharpe"]
    near = min(parts, key=lambda r: (r.target_risk["key"], r.target_risk["global_index"]))
    assert near.target_risk["global_index"] == whole.target_risk["global_index"]
    assert sum(p.n_accepted for p in parts) == P


# ---- frontier envelope (config C5) -------------------------------------------

def _envelope_oracle(risks, returns, n_bins, lo, hi, dtype):
    """oracle.paths_np.envelope evaluated in the kernel's arithmetic type for the bin index."""
    T = np.float32 if dtype == "float32" else np.float64
    from oracle import paths_np
    r = np.asarray(risks, dtype=T)
    scale = T(n_bins / (hi - lo))
    b = np.floor((r - T(lo)) * scale).astype(np.int64)
    b[(r == T(hi)) | (b >= n_bins)] = n_bins - 1
    ok = (r >= T(lo)) & (r <= T(hi))
    best = np.full(n_bins, -np.inf)
    idx = np.full(n_bins, -1, dtype=np.int64)
    ret = np.asarray(returns, dtype=np.float64)
    for i in np.nonzero(ok)[0]:
        if ret[i] > best[b[i]]:
            best[b[i]], idx[b[i]] = ret[i], i
    return best, idx


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n,P", [(16, 200_000), (64, 20_000)])
def test_envelope_matches_oracle(mcp, n, P, dtype):
    mu, sigma = synthetic_inputs(n)
    full = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=2, dtype=dtype)
    lo, hi = float(full.risks.min()), float(full.risks.max())
    K = 64
    r = mcp.frontier_envelope(mu, sigma, P, K, risk_free=0.03, seed=2, dtype=dtype)
    assert r.risk_range == (lo, hi)
    env = r.extra["envelope"]
    best, idx = _envelope_oracle(full.risks, full.returns, K, lo, hi, dtype)
    assert np.array_equal(env["best_index"], idx)
    assert np.array_equal(env["best_return"], best)
    assert (idx >= 0).sum() > K // 2 and env["edges"].shape == (K + 1,)
    # the FP64 run is also the plain numpy spec (oracle.paths_np.envelope)
    if dtype == "float64":
        from oracle import paths_np
        b2, i2 = paths_np.envelope(full.risks, full.returns, K, lo, hi)
        assert np.array_equal(i2, idx) and np.array_equal(b2, best)
    # explicit range narrower than the data: out-of-range portfolios are ignored
    mid = (lo + hi) / 2
    r2 = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=2, dtype=dtype, n_bins=8, risk_range=(lo, mid),
                                 return_arrays=False)
    best2, idx2 = _envelope_oracle(full.risks, full.returns, 8, lo, mid, dtype)
    assert np.array_equal(r2.extra["envelope"]["best_index"], idx2)


def test_envelope_with_arrays_and_host_chunks(mcp):
    """Envelope while arrays stream back through the two-slot HOST pipeline (several chunks)."""
    n = 32
    mu, sigma = synthetic_inputs(n, seed=4)
    P = 1_500_001
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=9, n_bins=32, risk_range=(0.05, 0.5))
    best, idx = _envelope_oracle(r.risks, r.returns, 32, 0.05, 0.5, "float32")
    assert np.array_equal(r.extra["envelope"]["best_index"], idx)
    assert np.array_equal(r.extra["envelope"]["best_return"], best)


# ---- FP32 screen + FP64 decision: the same selected index as the FP64 reference ------------------

@pytest.mark.parametrize("n", [2, 16, 64])
@pytest.mark.parametrize("space", ["host", "device"])
def test_fp32_near_ties_pick_the_fp64_index(mcp, n, space):
    """Rows that are identical after rounding to FP32 but differ in FP64: np.argmax on the FP64
    values picks the later, slightly better row; the plain FP32 sweep would pick the earlier one."""
    mu, sigma = synthetic_inputs(n, seed=5)
    rng = np.random.RandomState(7)
    P = 50_000
    W = rng.dirichlet(np.ones(n), size=P)
    base = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    i_s, i_d = base["max_sharpe"]["index"], base["target_risk"]["index"]
    # plant FP64-only improvements at higher indices
    def nudge(row, key_fn, better):
        best, out = key_fn(row), row
        for _ in range(200):
            cand = row * (1.0 + rng.standard_normal(n) * 2e-10)      # far below the FP32 ulp (6e-8 relative)
            cand = cand / cand.sum()
            if np.array_equal(cand.astype(np.float32), row.astype(np.float32)) and better(key_fn(cand), best):
                best, out = key_fn(cand), cand
        return out
    sharpe_of = lambda w: ref.portfolio_metrics(w[None], mu, sigma, 0.03)[2][0]
    dist_of = lambda w: abs(ref.portfolio_metrics(w[None], mu, sigma, 0.03)[1][0] - 0.30)
    W[P - 10] = nudge(W[i_s], sharpe_of, lambda a, b: a > b)
    W[P - 5] = nudge(W[i_d], dist_of, lambda a, b: a < b)
    want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    assert want["max_sharpe"]["index"] == P - 10 and want["target_risk"]["index"] == P - 5
    if space == "device":
        import torch
        arg = torch.from_numpy(W).cuda()
    else:
        arg = W
    r = mcp.simulate_portfolios(mu, sigma, P, weights=arg, risk_free=0.03, dtype="float32")
    assert r.max_sharpe["index"] == P - 10
    assert r.target_risk["index"] == P - 5
    assert np.isclose(r.max_sharpe["sharpe"], want["max_sharpe"]["sharpe"], rtol=1e-12)     # FP64 record
    assert np.allclose(r.max_sharpe["weights"], W[P - 10], rtol=0, atol=0)
    # without the FP64 copy the FP32 screen alone decides (first occurrence among FP32 ties)
    r32 = mcp.simulate_portfolios(mu, sigma, P, weights=W.astype(np.float32), risk_free=0.03, dtype="float32")
    assert r32.max_sharpe["index"] == i_s and r32.target_risk["index"] == i_d


def test_default_app_bounds_cost_nothing(mcp, synth16):
    """The app always passes min = 0 / max = 1 (app.py:453-454): identical results to no bounds."""
    mu, sigma = synth16
    a = mcp.simulate_portfolios(mu, sigma, 100_000, risk_free=0.03, seed=6)
    b = mcp.simulate_portfolios(mu, sigma, 100_000, risk_free=0.03, seed=6, min_weights=np.zeros(16), max_weights=np.ones(16))
    assert np.array_equal(a.weights, b.weights) and np.array_equal(a.sharpes, b.sharpes)
    assert a.max_sharpe["index"] == b.max_sharpe["index"] and b.n_accepted == 100_000
    # supplied rows outside [0, 1] are still checked against explicit bounds
    W = np.array([[0.5, 0.5] + [0.0] * 14, [1.5, -0.5] + [0.0] * 14])
    r = mcp.simulate_portfolios(mu, sigma, 2, weights=W, min_weights=np.zeros(16), max_weights=np.ones(16), dtype="float64")
    assert r.n_accepted == 1 and len(r.risks) == 1


@pytest.mark.parametrize("rounds", [10, 7])      # 7: the optional Philox4x32-7 stream must pass the same statistical tier
def test_rng_weight_histogram_chi_square(mcp, rounds):
    """A single coordinate of a flat Dirichlet(N) is Beta(1, N-1): chi-square over 64 equal-mass bins,
    for N = 4 and N = 16, and pairwise sums (w_i + w_j ~ Beta(2, N-2)) to catch cross-asset correlation."""
    for N in (4, 16):
        mu, sigma = synthetic_inputs(N)
        P = 4_000_000
        W = mcp.simulate_portfolios(mu, sigma, P, seed=21, philox_rounds=rounds).weights.astype(np.float64)
        edges = 1 - (1 - np.linspace(0, 1, 65)) ** (1 / (N - 1))                   # Beta(1, N-1) quantiles
        for col in (0, N - 1):
            counts = np.histogram(W[:, col], bins=edges)[0]
            chi2 = ((counts - P / 64) ** 2 / (P / 64)).sum()
            assert chi2 < 63 + 5 * np.sqrt(2 * 63), (N, col, chi2)
        s = W[:, 0] + W[:, 1]                                                       # Beta(2, N-2): mean 2/N
        assert abs(s.mean() - 2 / N) < 5 * np.sqrt(2 * (N - 2) / (N * N * (N + 1)) / P)
        assert np.isclose(s.var(), 2 * (N - 2) / (N * N * (N + 1)), rtol=0.01)


@pytest.mark.parametrize("n,dtype", [(16, "float32"), (100, "float32"), (40, "float64")])
def test_single_sweep_envelope_equals_two_sweeps(mcp, n, dtype):
    """frontier_envelope keeps (risk, return) in HBM and bins them afterwards (mcp_envelope_arrays); the bins,
    picks and range must be those of the two-sweep route (range sweep, then a binning sweep with n_bins set)."""
    mu, sigma = synthetic_inputs(n, seed=5)
    P, K = 300_000, 97
    one = mcp.frontier_envelope(mu, sigma, P, K, risk_free=0.03, seed=6, dtype=dtype, first_index=12345)
    two = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=6, dtype=dtype, first_index=12345, return_arrays=False,
                                  n_bins=K, risk_range=one.risk_range)
    a, b = one.extra["envelope"], two.extra["envelope"]
    assert np.array_equal(a["best_index"], b["best_index"]) and np.array_equal(a["best_return"], b["best_return"])
    assert np.array_equal(a["edges"], b["edges"]) and (a["best_index"] >= 0).sum() > K // 2
    assert one.risks is None and one.n_accepted == two.n_accepted == P
    for pick in ("max_sharpe", "target_risk"):
        x, y = getattr(one, pick), getattr(two, pick)
        assert x["global_index"] == y["global_index"] and x["sharpe"] == y["sharpe"] and x["risk"] == y["risk"]
        assert np.array_equal(x["weights"], y["weights"])
    # the public binning entry on arrays the caller holds
    import torch
    full = mcp.simulate_portfolios(mu, sigma, 50_000, risk_free=0.03, seed=6, dtype=dtype, return_arrays="device")
    lo, hi = full.risk_range
    env = mcp.envelope_from_arrays(full.risks, full.returns, 32, (lo, hi), first_index=7)
    best, idx = _envelope_oracle(full.risks.cpu().numpy(), full.returns.cpu().numpy(), 32, lo, hi, dtype)
    assert np.array_equal(env["best_return"], best) and np.array_equal(env["best_index"], np.where(idx >= 0, idx + 7, -1))
    with pytest.raises(TypeError):
        mcp.envelope_from_arrays(np.zeros(4), np.zeros(4), 4, (0.0, 1.0))
    with pytest.raises(ValueError):
        mcp.envelope_from_arrays(full.risks, full.returns, 4, (1.0, 1.0))
