"""GPU tier: correlated-path simulator and exact VaR / CVaR against the oracle."""
import numpy as np
import pytest

from conftest import synthetic_inputs
from oracle import paths_np, philox_np, reference_np as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-6), ("float32", 1e-4)])
@pytest.mark.parametrize("n,steps", [(16, 252), (3, 17), (32, 5), (14, 60)])
def test_supplied_normals_parity(mcp, n, steps, dtype, tol):
    mu, sigma = synthetic_inputs(n, seed=2)
    rng = np.random.default_rng(n)
    M = 777
    Z = rng.standard_normal((M, steps, n))
    w = rng.dirichlet(np.ones(n))
    out = mcp.simulate_paths(mu, sigma, w, M, steps, normals=Z, dtype=dtype, alphas=(0.95, 0.99))
    Zc = Z.astype(np.float32).astype(np.float64) if dtype == "float32" else Z
    want = paths_np.terminal_returns(mu, sigma, w, Zc)
    # tolerance is relative to the terminal VALUE w.V_T (= 1 + return): returns cross zero
    assert np.allclose(out["terminal"] + 1.0, want + 1.0, rtol=tol)
    st = paths_np.risk_stats(want)
    for a, (v, c) in out["stats"].items():
        assert np.isclose(v + 1, st[a][0] + 1, rtol=tol) and np.isclose(c + 1, st[a][1] + 1, rtol=tol)


@pytest.mark.parametrize("dtype,tol", [("float32", 2e-4), ("float64", 1e-9)])
def test_rng_paths_match_generator_restatement(mcp, dtype, tol):
    n, steps, M = 16, 12, 1000
    mu, sigma = synthetic_inputs(n)
    w = np.full(n, 1 / n)
    first, seed = 5_000_000_000, 77
    out = mcp.simulate_paths(mu, sigma, w, M, steps, seed=seed, first_index=first, dtype=dtype)
    Z = philox_np.normals(first, M, steps, n, seed, dtype)
    want = paths_np.terminal_returns(mu, sigma, w, Z)
    assert np.allclose(out["terminal"] + 1, want + 1, rtol=tol)


@pytest.mark.parametrize("rounds", [10, 7])      # 7: the optional Philox4x32-7 stream must pass the same statistical tier
def test_rng_paths_statistics(mcp, rounds):
    """Mean of the terminal value is analytic: E[V_T,i] = (1 + mu_i dt)^S; sharding invariance."""
    n, S, M = 16, 252, 400_000
    mu, sigma = synthetic_inputs(n)
    w = np.random.default_rng(1).dirichlet(np.ones(n))
    out = mcp.simulate_paths(mu, sigma, w, M, S, seed=1, return_terminal=True, philox_rounds=rounds)
    x = out["terminal"].astype(np.float64)
    want = w @ ((1 + mu / 252) ** S) - 1
    assert abs(x.mean() - want) < 4.5 * x.std() / np.sqrt(M)
    # variance of log terminal value of a single asset ~ sigma_ii (weights on one asset)
    e0 = np.zeros(n); e0[0] = 1
    x0 = mcp.simulate_paths(mu, sigma, e0, 200_000, S, seed=2, return_terminal=True, philox_rounds=rounds)["terminal"].astype(np.float64)
    assert np.isclose(np.log1p(x0).var(), sigma[0, 0], rtol=0.03)
    # a shard [a, b) of the same seed reproduces the same terminal values bit for bit
    part = mcp.simulate_paths(mu, sigma, w, 1000, S, seed=1, first_index=5000, return_terminal=True, philox_rounds=rounds)["terminal"]
    assert np.array_equal(part, out["terminal"][5000:6000])
    for a, (v, c) in out["stats"].items():
        assert v == ref.var(x, a) and np.isclose(c, ref.cvar(x, a), rtol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 21, 100, 4097, 1_000_003])
def test_quantiles_exact_vs_numpy(mcp, n, dtype):
    """Exact order statistics + numpy's lerp: bit-equal VaR, CVaR to FP64 rounding."""
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 0.3).astype(dtype)
    if n > 50:
        x[::7] = np.round(x[::7], 1)            # ties
        x[5] = -0.0; x[6] = 0.0
    alphas = (0.95, 0.99, 0.5, 0.0, 1.0)
    for src in ("host", "device"):
        if src == "device":
            import torch
            arg = torch.from_numpy(x).cuda()
        else:
            arg = x
        st = mcp.quantile_stats(arg, alphas)
        xd = x.astype(np.float64)
        for a in alphas:
            v, c = st[a]
            assert v == ref.var(xd, a), (n, a, src)
            assert np.isclose(c, ref.cvar(xd, a), rtol=1e-12, atol=1e-300)


def test_quantiles_sharded_callback(mcp):
    """The multi-rank VaR merge on one GPU: three uneven shards of one vector, one libmcp handle
    and host thread each, histograms / tail sums summed through the all-reduce callback."""
    import torch
    from mcportfolio.dist import emulate_sharded_quantiles
    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.float64):
        x = rng.standard_normal(300_001).astype(dtype)
        parts = [torch.from_numpy(p).cuda() for p in np.split(x, [100_000, 100_007])]
        st = emulate_sharded_quantiles(parts, (0.95, 0.99))
        xd = x.astype(np.float64)
        for a, (v, c) in st.items():
            assert v == ref.var(xd, a) and np.isclose(c, ref.cvar(xd, a), rtol=1e-12)


def test_cholesky_failure_is_reported(mcp):
    mu = np.array([0.1, 0.2])
    bad = np.array([[1.0, 2.0], [2.0, 1.0]])
    with pytest.raises(mcp.McpError, match="positive definite"):
        mcp.simulate_paths(mu, bad, np.array([0.5, 0.5]), 10, 5)


@pytest.mark.parametrize("rounds", [10, 7])      # 7: the optional Philox4x32-7 stream must pass the same statistical tier
def test_rng_normals_distribution(mcp, rounds):
    """Statistical tier for the in-kernel Box-Muller normals: with L = I, one step and dt = 1 the
    terminal value of asset i is z_i, so the kernel's normals can be tested directly: moments,
    independence across assets and a KS test of the marginal against the exact normal CDF."""
    from math import erf
    n, M = 16, 1_000_000
    sigma = np.eye(n)
    mu = np.zeros(n)
    cols = []
    for i in (0, 1, 7, 15):
        w = np.zeros(n); w[i] = 1.0
        x = mcp.simulate_paths(mu, sigma, w, M, 1, dt=1.0, seed=11, return_terminal=True, philox_rounds=rounds)["terminal"].astype(np.float64)
        cols.append(x)
        assert abs(x.mean()) < 4.5 / np.sqrt(M) and abs(x.var() - 1) < 5 * np.sqrt(2 / M)
        assert abs(((x - x.mean()) ** 3).mean()) < 0.02 and abs(((x - x.mean()) ** 4).mean() - 3) < 0.05
        xs = np.sort(x[:200_000])
        cdf = 0.5 * (1 + np.vectorize(erf)(xs / np.sqrt(2)))
        ks = np.abs(cdf - (np.arange(len(xs)) + 0.5) / len(xs)).max()
        assert ks < 1.63 / np.sqrt(len(xs)) * 1.3
    c = np.corrcoef(np.stack(cols))
    assert np.abs(c - np.eye(4)).max() < 5 / np.sqrt(M)          # (z0, z1) share a Box-Muller pair: still uncorrelated
    # consecutive steps are independent: two-step compounding variance of log(1 + r) adds up
    w = np.zeros(n); w[3] = 1.0
    x2 = mcp.simulate_paths(mu, 1e-4 * sigma, w, M, 2, dt=1.0, seed=12, return_terminal=True, philox_rounds=rounds)["terminal"].astype(np.float64)
    assert np.isclose(x2.var(), 2e-4, rtol=0.01)
