"""ncu target: the packed path kernel only (C4 shape: 16 assets, 252 steps; one GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import mcportfolio as mcp
from bench import synthetic_inputs
mu, sigma = synthetic_inputs(16)
w = np.full(16, 1 / 16)
M = int(os.environ.get("PROFILE_M", 2_000_000))
for _ in range(3):
    o = mcp.simulate_paths(mu, sigma, w, M, 252, seed=0, return_terminal=False)
    print(f"paths M={M}: {M * 252 / o['kernel_ms'] * 1e3:.4g} path-steps/s")
