"""GPU tier: the tcgen05 sweep (32 < N <= 256, FP32, in-kernel RNG, no bounds) against the FP64 oracle
and against the SIMT kernel it replaces on that path.

The tensor cores see TF32 / BF16 operands; the split (hi + lo, Shi + Slo) must give FP32-class results:
the north-star FP32 tolerance is 1e-4 relative, the tests below also pin the much tighter agreement
with the plain-FP32 SIMT kernel so that a lost correction term (which would still pass 1e-4 on benign
inputs) is caught.
"""
import os

import numpy as np
import pytest

from conftest import synthetic_inputs
from oracle import philox_np, reference_np as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


class simt_kernel:
    """MCP_LARGE_TC=0 routes the same call to the SIMT large_sweep (read per launch)."""

    def __enter__(self):
        self.old = os.environ.get("MCP_LARGE_TC")
        os.environ["MCP_LARGE_TC"] = "0"

    def __exit__(self, *exc):
        if self.old is None:
            del os.environ["MCP_LARGE_TC"]
        else:
            os.environ["MCP_LARGE_TC"] = self.old


class tf32_split:
    """MCP_LARGE_TC_F16=0 makes Philox rows use the TF32 + BF16 operand split that supplied weights always use."""

    def __enter__(self):
        self.old = os.environ.get("MCP_LARGE_TC_F16")
        os.environ["MCP_LARGE_TC_F16"] = "0"

    def __exit__(self, *exc):
        if self.old is None:
            del os.environ["MCP_LARGE_TC_F16"]
        else:
            os.environ["MCP_LARGE_TC_F16"] = self.old


def wild_sigma(n, seed):
    """Covariance with variances over six decades and strong common factors: cancellation in w'Sw and
    entries whose TF32 rounding error alone would exceed the FP32 tolerance."""
    rng = np.random.default_rng(seed)
    F = rng.standard_normal((n, 3))
    vol = 10.0 ** rng.uniform(-3, 0, n)
    corr = F @ F.T + np.diag(rng.uniform(0.05, 0.5, n))
    d = np.sqrt(np.diag(corr))
    corr = corr / d[:, None] / d[None, :]
    sigma = corr * vol[:, None] * vol[None, :]
    mu = rng.uniform(-0.2, 0.6, n)
    return mu, sigma


@pytest.mark.parametrize("n", [33, 50, 64, 65, 96, 128, 160, 200, 255, 256])
def test_tc_matches_oracle_and_simt(mcp, n):
    mu, sigma = synthetic_inputs(n, seed=n)
    P, first, seed = 1000, 7_000_000_123, 11
    tc = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=seed, first_index=first, dtype="float32")
    with simt_kernel():
        sm = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=seed, first_index=first, dtype="float32")
    W, valid = philox_np.dirichlet_weights(first, P, n, seed, "float32")
    assert tc.n_accepted == P and valid.all()
    assert np.allclose(tc.weights, W, atol=1e-6) and np.allclose(tc.weights, sm.weights, atol=2e-8, rtol=2e-6)
    assert np.abs(tc.weights.sum(1) - 1).max() < 2e-6
    want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    # north-star FP32 tolerance against the FP64 oracle
    assert np.allclose(tc.risks, want["risks"], rtol=1e-4) and np.allclose(tc.returns, want["returns"], rtol=1e-4)
    assert np.allclose(tc.sharpes, want["sharpes"], rtol=1e-4, atol=1e-4)
    # FP32-class agreement with the SIMT kernel: a missing lo / Slo term would show up as ~2e-4
    assert np.abs(tc.risks / sm.risks - 1).max() < 5e-6
    assert np.abs(tc.returns / sm.returns - 1).max() < 5e-6
    # the record of a pick is the sweep's own value, bit for bit
    i = tc.max_sharpe["index"]
    assert tc.max_sharpe["sharpe"] == float(tc.sharpes[i]) and tc.max_sharpe["risk"] == float(tc.risks[i])
    assert np.array_equal(tc.max_sharpe["weights"], tc.weights[i])
    j = tc.target_risk["index"]
    assert j == int(np.argmin(np.abs(tc.risks - np.float32(0.30))))


@pytest.mark.parametrize("n", [48, 256])
def test_tc_accuracy_on_ill_scaled_covariance(mcp, n):
    mu, sigma = wild_sigma(n, seed=n)
    P, seed = 4000, 5
    tc = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.0, seed=seed, dtype="float32")
    W, _ = philox_np.dirichlet_weights(0, P, n, seed, "float32")
    want = ref.evaluate(W, mu, sigma, 0.0, 0.30)
    err = np.abs(tc.risks / want["risks"] - 1).max()
    assert err < 2e-5, err                                  # 1e-4 is the bar; the split leaves FP32-class error
    assert np.allclose(tc.returns, want["returns"], rtol=1e-4, atol=1e-7)
    # the FP64 pick is inside the FP32 near-tie set
    i64 = want["max_sharpe"]["index"]
    assert tc.sharpes[i64] >= tc.max_sharpe["sharpe"] * (1 - 1e-5) - 1e-6


def test_tc_tail_tiles_and_index_ranges(mcp):
    """P around the 128-row tile, ranges cut anywhere: a portfolio's values do not depend on its tile slot."""
    n = 100
    mu, sigma = synthetic_inputs(n, seed=2)
    whole = mcp.simulate_portfolios(mu, sigma, 1000, risk_free=0.03, seed=4, first_index=10**10, dtype="float32")
    for P in (1, 2, 127, 128, 129, 255, 257):
        r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=4, first_index=10**10, dtype="float32")
        assert r.n_accepted == P
        assert np.array_equal(r.risks, whole.risks[:P]) and np.array_equal(r.sharpes, whole.sharpes[:P])
        assert np.array_equal(r.weights, whole.weights[:P])
    for a, b in ((1, 130), (127, 129), (500, 1000), (999, 1000)):
        r = mcp.simulate_portfolios(mu, sigma, b - a, risk_free=0.03, seed=4, first_index=10**10 + a, dtype="float32")
        assert np.array_equal(r.risks, whole.risks[a:b]) and np.array_equal(r.returns, whole.returns[a:b])
        assert np.array_equal(r.weights, whole.weights[a:b])


def test_tc_many_tiles_selection_matches_arrays(mcp):
    """A range long enough that every CTA loops over several tiles; no-array run picks what the array run holds."""
    n = 64
    mu, sigma = synthetic_inputs(n, seed=9)
    P = 148 * 128 * 3 + 77
    full = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=1, dtype="float32")
    lean = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=1, dtype="float32", return_arrays=False)
    assert lean.n_accepted == P == full.n_accepted
    i = int(np.argmax(full.sharpes))
    assert lean.max_sharpe["global_index"] == i and lean.max_sharpe["sharpe"] == float(full.sharpes[i])
    j = int(np.argmin(np.abs(full.risks - np.float32(0.30))))
    assert lean.target_risk["global_index"] == j and lean.target_risk["risk"] == float(full.risks[j])
    assert lean.risk_range == (float(full.risks.min()), float(full.risks.max()))
    with simt_kernel():
        sm = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=1, dtype="float32", return_arrays=False)
    assert np.isclose(sm.max_sharpe["sharpe"], lean.max_sharpe["sharpe"], rtol=5e-6)


@pytest.mark.parametrize("n", [33, 64, 100, 254, 256])
def test_tc_supplied_weights_parity_mode(mcp, n):
    """Supplied weights through the tensor-core kernel: identical inputs as the FP64 oracle, 1e-4, same picks."""
    mu, sigma = wild_sigma(n, seed=100 + n) if n in (64, 256) else synthetic_inputs(n, seed=n)
    P = 3000
    W = np.random.RandomState(n).dirichlet(np.ones(n) * 0.5, size=P)          # sparse-ish weights: some near 1, many near 0
    W[0] = 0; W[0, n - 1] = 1.0                                               # a one-asset portfolio (last column: the zero-padded chunk)
    W[1] = 1.0 / n
    tc = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.01, risk_target=0.2, dtype="float32")      # FP64 host weights: the FP32 sweep is a screen, near-ties are decided in FP64
    want = ref.evaluate(W, mu, sigma, 0.01, 0.2)
    assert tc.n_accepted == P
    assert np.allclose(tc.risks, want["risks"], rtol=1e-4) and np.allclose(tc.returns, want["returns"], rtol=1e-4, atol=1e-7)
    assert np.allclose(tc.sharpes, want["sharpes"], rtol=1e-4, atol=1e-4 * np.abs(want["sharpes"]).max())
    assert np.array_equal(tc.weights, W.astype(np.float32))
    assert np.isclose(tc.risks[0], np.sqrt(sigma[n - 1, n - 1]), rtol=2e-6)
    for pick in ("max_sharpe", "target_risk"):
        assert getattr(tc, pick)["index"] == want[pick]["index"], pick        # FP32 screen + FP64 decision
    with simt_kernel():
        sm = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.01, risk_target=0.2, dtype="float32")
    # both kernels are FP32-class: neither may be further from the FP64 oracle than a few FP32 roundings of a
    # cancelling sum (the ill-scaled covariances lose ~1e-5 in either), and they agree with each other likewise
    e_tc, e_sm = np.abs(tc.risks / want["risks"] - 1).max(), np.abs(sm.risks / want["risks"] - 1).max()
    assert e_tc < 2e-5 and e_tc < 3 * e_sm + 2e-6, (e_tc, e_sm)
    assert np.abs(tc.risks / sm.risks - 1).max() < 2e-5


@pytest.mark.parametrize("n,scale", [(40, 1.0), (256, 1.0), (96, 1e-12), (200, 1e9), (64, 3e-30)])
def test_fp16_split_agrees_with_tf32_split_and_oracle_at_any_covariance_scale(mcp, n, scale):
    """Philox rows run the FP16 operand split (S' stored times a power of two so that FP16's range is never the issue):
    same weights, FP32-class agreement with the TF32 split and with the FP64 oracle, whatever the units of Sigma."""
    mu, sigma = wild_sigma(n, seed=7 * n)
    sigma = sigma * scale
    P, seed = 3000, 21
    f16 = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.0, seed=seed, dtype="float32")
    with tf32_split():
        t32 = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.0, seed=seed, dtype="float32")
    assert np.array_equal(f16.weights, t32.weights)                       # the generator is the same code
    W, _ = philox_np.dirichlet_weights(0, P, n, seed, "float32")
    want = ref.evaluate(W, mu, sigma, 0.0, 0.30)
    e16, e32 = np.abs(f16.risks / want["risks"] - 1).max(), np.abs(t32.risks / want["risks"] - 1).max()
    assert e16 < 2e-5 and e32 < 2e-5, (e16, e32)                          # 1e-4 is the bar
    assert np.abs(f16.risks / t32.risks - 1).max() < 1e-5
    assert np.allclose(f16.returns, t32.returns, rtol=2e-6, atol=1e-9)
    assert np.isfinite(f16.sharpes).all()
    i = f16.max_sharpe["index"]
    assert f16.max_sharpe["risk"] == float(f16.risks[i]) and f16.max_sharpe["sharpe"] == float(f16.sharpes[i])


def test_fp16_split_zero_and_tiny_covariance(mcp):
    """Sigma = 0 (every risk exactly 0, Sharpe 0 by the reference's guard, app.py:711) and a covariance of denormal size."""
    n = 48
    mu = np.linspace(0.01, 0.2, n)
    z = mcp.simulate_portfolios(mu, np.zeros((n, n)), 500, risk_free=0.0, seed=1, dtype="float32")
    assert np.all(z.risks == 0) and np.all(z.sharpes == 0) and np.isfinite(z.returns).all()
    _, sigma = wild_sigma(n, seed=3)
    t = mcp.simulate_portfolios(mu, sigma * 1e-36, 500, risk_free=0.0, seed=1, dtype="float32")
    W, _ = philox_np.dirichlet_weights(0, 500, n, 1, "float32")
    want = ref.evaluate(W, mu, sigma * 1e-36, 0.0, 0.30)
    assert np.allclose(t.risks, want["risks"], rtol=1e-4)
