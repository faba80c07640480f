"""GPU tier: randomised shapes (hypothesis) through the C ABI against the oracle -- every kernel
family (register kernel N <= 32, tiled N <= 256, generic), ragged sizes, both dtypes."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from conftest import synthetic_inputs
from oracle import paths_np, philox_np, reference_np as ref

pytestmark = pytest.mark.gpu
RTOL = {"float64": 1e-6, "float32": 1e-4}


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 70), P=st.integers(1, 700), dtype=st.sampled_from(["float32", "float64"]),
       seed=st.integers(0, 2**40), first=st.integers(0, 2**45), rf=st.sampled_from([0.0, 0.03, 3.0]))
def test_rng_sweep_any_shape(mcp, n, P, dtype, seed, first, rf):
    mu, sigma = synthetic_inputs(n, seed=n)
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=rf, seed=seed, first_index=first, dtype=dtype)
    W, _ = philox_np.dirichlet_weights(first, P, n, seed, dtype)
    s = philox_np.exponentials(np.arange(first, first + P, dtype=np.uint64), 0, n, seed, dtype).sum(1, keepdims=True)
    assert np.all(np.abs(r.weights - W) <= (2e-6 + 6e-7 / s if dtype == "float32" else 1e-12))
    own = ref.evaluate(np.asarray(r.weights, dtype=np.float64), mu, sigma, rf, 0.30)
    tol = RTOL[dtype]
    assert np.allclose(r.risks, own["risks"], rtol=tol) and np.allclose(r.returns, own["returns"], rtol=tol, atol=tol * 1e-2)
    assert np.allclose(r.sharpes, own["sharpes"], rtol=tol, atol=tol * (1 + abs(rf)) * 10)
    assert r.max_sharpe["index"] == int(np.argmax(r.sharpes))
    assert r.target_risk["index"] == int(np.argmin(np.abs(r.risks - r.risks.dtype.type(0.30))))
    assert r.max_sharpe["global_index"] == first + r.max_sharpe["index"] and r.n_accepted == P


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 70), P=st.integers(1, 500), dtype=st.sampled_from(["float32", "float64"]), seed=st.integers(0, 1000))
def test_supplied_sweep_any_shape(mcp, n, P, dtype, seed):
    mu, sigma = synthetic_inputs(n, seed=n + 1)
    W = np.random.RandomState(seed).dirichlet(np.ones(n), size=P)
    r = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03, risk_target=0.2, dtype=dtype)
    want = ref.evaluate(W, mu, sigma, 0.03, 0.2)
    tol = RTOL[dtype]
    assert np.allclose(r.risks, want["risks"], rtol=tol) and np.allclose(r.sharpes, want["sharpes"], rtol=tol, atol=tol)
    assert r.max_sharpe["index"] == want["max_sharpe"]["index"]
    assert r.target_risk["index"] == want["target_risk"]["index"]


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 32), M=st.integers(1, 300), S=st.integers(1, 40), dtype=st.sampled_from(["float32", "float64"]),
       seed=st.integers(0, 2**30), first=st.integers(0, 2**40))
def test_paths_any_shape(mcp, n, M, S, dtype, seed, first):
    mu, sigma = synthetic_inputs(n, seed=n + 2)
    w = np.random.default_rng(seed).dirichlet(np.ones(n))
    out = mcp.simulate_paths(mu, sigma, w, M, S, seed=seed, first_index=first, dtype=dtype, alphas=(0.95, 0.5))
    Z = philox_np.normals(first, M, S, n, seed, dtype)
    want = paths_np.terminal_returns(mu, sigma, w, Z)
    assert np.allclose(out["terminal"] + 1, want + 1, rtol=3e-4 if dtype == "float32" else 1e-9)
    x = out["terminal"].astype(np.float64)
    for a, (v, c) in out["stats"].items():
        assert v == ref.var(x, a) and np.isclose(c, ref.cvar(x, a), rtol=1e-12, atol=1e-300)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(33, 256), P=st.integers(1, 400), seed=st.integers(0, 2**40), first=st.integers(0, 2**45),
       supplied=st.booleans())
def test_tensor_core_sweep_any_shape(mcp, n, P, seed, first, supplied):
    """FP32, 32 < N <= 256: the tcgen05 kernel (Philox rows or supplied rows) against the FP64 oracle."""
    mu, sigma = synthetic_inputs(n, seed=n + 3)
    if supplied:
        W = np.random.RandomState(seed % 2**31).dirichlet(np.ones(n), size=P)
        r = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03, dtype="float32")
        assert np.array_equal(r.weights, W.astype(np.float32))
    else:
        r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=seed, first_index=first, dtype="float32")
        W, _ = philox_np.dirichlet_weights(first, P, n, seed, "float32")
        assert np.allclose(r.weights, W, atol=2e-6)
    want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    assert r.n_accepted == P
    assert np.allclose(r.risks, want["risks"], rtol=1e-4) and np.allclose(r.returns, want["returns"], rtol=1e-4)
    assert np.allclose(r.sharpes, want["sharpes"], rtol=1e-4, atol=1e-3)
    i = r.max_sharpe["index"]
    assert r.max_sharpe["sharpe"] == float(r.sharpes[i]) or supplied       # supplied FP64 rows: the pick is decided in FP64
    assert np.isclose(r.sharpes[i], r.sharpes.max(), rtol=1e-5)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(33, 256), P=st.integers(1, 3000), seed=st.integers(0, 2**40), first=st.integers(0, 2**45),
       hi_scale=st.floats(1.5, 6.0), lo_scale=st.sampled_from([0.0, 0.0, 0.02, 0.2]), tries=st.integers(1, 6), keep_last=st.booleans())
def test_bounded_tensor_core_sweep_any_shape(mcp, n, P, seed, first, hi_scale, lo_scale, tries, keep_last):
    """Bounds rejection at 32 < N <= 256 (app.py:700-707): the tcgen05 rounds + SIMT list route against the SIMT-only sweep
    (same accept / skip decisions, same attempts) and against the generator restatement with the reference's rejection loop."""
    import os
    mu, sigma = synthetic_inputs(n, seed=n + 5)
    hi = np.full(n, hi_scale / n)                        # the largest of n flat-Dirichlet weights is ~ (ln n + 0.58) / n
    lo = np.full(n, lo_scale / (n * n)) if lo_scale else None      # the smallest is ~ 1 / n^2
    kw = dict(min_weights=lo, max_weights=hi, seed=seed, first_index=first, max_tries=tries, keep_last=keep_last, risk_free=0.03)
    a = mcp.simulate_portfolios(mu, sigma, P, **kw)
    os.environ["MCP_LARGE_TC_BOUNDS"] = "0"
    try:
        b = mcp.simulate_portfolios(mu, sigma, P, **kw)
    finally:
        os.environ.pop("MCP_LARGE_TC_BOUNDS")
    assert a.n_accepted == b.n_accepted and np.array_equal(a.accepted, b.accepted)
    if keep_last:
        assert a.n_accepted == P
    if a.n_accepted:
        assert np.allclose(a.weights, b.weights, rtol=0, atol=3e-7)
        assert np.allclose(a.risks, b.risks, rtol=2e-5) and np.allclose(a.sharpes, b.sharpes, rtol=2e-5, atol=2e-5)
        assert a.max_sharpe["index"] == int(np.argmax(a.sharpes))
        W, valid = philox_np.dirichlet_weights(first, P, n, seed, "float32", lo, hi, max_tries=tries, keep_last=keep_last)
        acc = a.accepted.astype(bool)
        edge = np.abs(W - hi).min(1) if lo is None else np.minimum(np.abs(W - hi).min(1), np.abs(W - lo).min(1))
        differ = acc != valid
        assert differ.sum() <= 2 and np.all(edge[differ] < 1e-6)
        both = acc & valid
        assert np.allclose(a.weights[np.cumsum(acc)[both] - 1], W[both], atol=2e-6)
    else:
        assert a.max_sharpe is None


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(T=st.integers(2, 520), n=st.integers(1, 24), P=st.integers(1, 300), alpha=st.sampled_from([0.9, 0.95, 0.99, 0.5]),
       decimals=st.sampled_from([2, 3, 8]), dtype=st.sampled_from(["float32", "float64"]), seed=st.integers(0, 10**6))
def test_historical_var_cvar_any_shape(mcp, T, n, P, alpha, decimals, dtype, seed):
    """Per-portfolio historical VaR / CVaR (app.py:710-713): fast FP32 kernel, plain kernel (FP64, deep ranks, short series)
    and ties (rounded returns) against the oracle's np.percentile restatement."""
    rng = np.random.default_rng(seed)
    R = np.round(rng.standard_normal((T, n)) * 0.03, decimals)
    W = rng.dirichlet(np.ones(n), size=P)
    out = mcp.historical_var_cvar(R, W, alpha, dtype=dtype)
    v, c = ref.historical_var_cvar(R, W, alpha)
    if dtype == "float64":
        assert np.allclose(out["var"], v, rtol=1e-10, atol=1e-15) and np.allclose(out["cvar"], c, rtol=1e-10, atol=1e-15)
    else:
        # FP32 series: an order statistic can swap with a neighbour that differs by FP32 rounding only -> compare values
        scale = np.abs(R).max() + 1e-12
        assert np.allclose(out["var"], v, rtol=1e-4, atol=2e-6 * scale + 1e-9)
        # CVaR = mean of the tail INCLUDING the VaR element(s): with tied values an FP32 rounding can move a whole tie group
        # in or out of the tail, so the FP32 check needs room for one element of the tail mean
        k = max(1, int((T - 1) * (1 - alpha)) + 1)
        assert np.all(np.abs(out["cvar"] - c) <= 1e-4 * np.abs(c) + (np.abs(v - c) + 4e-6 * scale) * (T if decimals < 8 else 1) / k + 1e-9)
    assert out["best_var"]["index"] == int(np.argmax(out["var"])) and out["best_cvar"]["index"] == int(np.argmax(out["cvar"]))
