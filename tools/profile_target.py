"""Short, fixed workload for ncu: every hot kernel launched a few times (one GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import torch
import mcportfolio as mcp
from bench import synthetic_inputs

mu, sigma = synthetic_inputs(16)
w = np.full(16, 1 / 16)
P = int(os.environ.get("PROFILE_P", 1_000_000_000))
for _ in range(3):
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=0, return_arrays=False)
print("sweep", P / r.kernel_ms * 1e3, "pf/s")
for _ in range(2):
    rw = mcp.simulate_portfolios(mu, sigma, 20_000_000, risk_free=0.03, seed=0, return_arrays="device")
print("write-back", 20_000_000 / rw.kernel_ms * 1e3, "pf/s")
for _ in range(2):
    rs = mcp.simulate_portfolios(mu, sigma, 20_000_000, weights=rw.weights, risk_free=0.03)
print("supplied", 20_000_000 / rs.kernel_ms * 1e3, "pf/s")
for _ in range(2):
    o = mcp.simulate_paths(mu, sigma, w, 1_000_000, 252, seed=0, return_terminal=False)
print("paths", 1_000_000 * 252 / o["kernel_ms"] * 1e3, "path-steps/s", o["stats"])
r64 = mcp.simulate_portfolios(mu, sigma, 100_000_000, risk_free=0.03, seed=0, return_arrays=False, dtype="float64")
print("sweep f64", 100_000_000 / r64.kernel_ms * 1e3, "pf/s")

mu256, sigma256 = synthetic_inputs(256)
for _ in range(2):
    r256 = mcp.simulate_portfolios(mu256, sigma256, 4_000_000, risk_free=0.03, seed=0, return_arrays=False)
print("sweep N=256", 4_000_000 / r256.kernel_ms * 1e3, "pf/s")
