"""Timing target: the per-portfolio historical VaR/CVaR kernel (T=365, N=16), fast FP32 path vs MCP_HIST_FAST=0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import torch
import mcportfolio as mcp

rng = np.random.default_rng(0)
T, n, P = 365, 16, 4_000_000
R = rng.standard_normal((T, n)) * 0.04
W = torch.from_numpy(rng.dirichlet(np.ones(n), size=P).astype(np.float32)).cuda()
CASES = [("fast", None, None), ("fast row=0", None, "0"), ("fast row=1", None, "1"), ("fast row=2", None, "2"), ("plain", "0", None)]
if os.environ.get("HIST_PROFILE_ONLY_FAST"):
    CASES = CASES[:1]
for tag, env, row in CASES:
    if env is None:
        os.environ.pop("MCP_HIST_FAST", None)
    else:
        os.environ["MCP_HIST_FAST"] = env
    if row is None:
        os.environ.pop("MCP_HIST_ROW", None)
    else:
        os.environ["MCP_HIST_ROW"] = row
    for alpha in (0.95, 0.99):
        for _ in range(3):
            out = mcp.historical_var_cvar(R, W, alpha, dtype="float32")
        ms = mcp.get_engine().last_kernel_ms()
        print(f"{tag} alpha={alpha}: {P / ms * 1e3:.4g} pf/s  kernel_ms={ms:.3f}  best_var={out['best_var']}")
