"""Tiny invocation of every kernel for compute-sanitizer (memcheck): small sizes, ragged tails."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import torch
import mcportfolio as mcp
from bench import synthetic_inputs

rng = np.random.default_rng(0)
for n in (2, 16, 20, 33, 100, 256):      # 33 / 100 / 256 in RNG mode without bounds run the tcgen05 sweep
    mu, sigma = synthetic_inputs(n)
    P = 1337
    for dtype in ("float32", "float64"):
        r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=1, dtype=dtype)
        W = np.asarray(r.weights, dtype=np.float64)
        mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03, dtype=dtype)
        mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=1, dtype=dtype, return_arrays=False,
                                max_weights=np.full(n, 0.9), max_tries=3)
        mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=1, dtype=dtype, n_bins=16, risk_range=(0.01, 1.0))
        if n <= 32:
            mcp.simulate_portfolios(mu, sigma, P, weights=torch.from_numpy(W).cuda(), risk_free=0.03, dtype=dtype)
mu, sigma = synthetic_inputs(16)
w = np.full(16, 1 / 16)
for dtype in ("float32", "float64"):
    mcp.simulate_paths(mu, sigma, w, 777, 9, seed=2, dtype=dtype)
    mcp.simulate_paths(mu, sigma, w, 333, 5, normals=rng.standard_normal((333, 5, 16)), dtype=dtype)
    R = rng.standard_normal((100, 16)) * 0.03
    mcp.historical_var_cvar(R, rng.dirichlet(np.ones(16), size=500), 0.95, dtype=dtype)
mcp.quantile_stats(rng.standard_normal(100_003).astype(np.float32), (0.95, 0.99))
mcp.quantile_stats(rng.standard_normal(5_001), (0.95, 0.5))
mcp.asset_stats(rng.standard_normal((365, 5)) * 0.05, annual_factor=52, risk_free=0.03)
print("sanitize target finished", mcp.get_engine().launch_count(), "launches")
