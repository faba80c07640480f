"""Bounded sweep at N = 256: tcgen05 + deferred rows vs the SIMT kernel (MCP_LARGE_TC_BOUNDS=0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import mcportfolio as mcp
from bench import synthetic_inputs

for n, hi, lo in ((256, 0.03, None), (256, 0.024, None), (256, 0.03, 1e-6), (64, 0.08, None)):
    mu, sigma = synthetic_inputs(n)
    P = 50_000_000 if n == 256 else 200_000_000
    kw = dict(max_weights=np.full(n, hi), min_weights=None if lo is None else np.full(n, lo), risk_free=0.03, seed=0, return_arrays=False)
    for mode in ("1", "0"):
        os.environ["MCP_LARGE_TC_BOUNDS"] = mode
        for _ in range(2):
            r = mcp.simulate_portfolios(mu, sigma, P, **kw)
        print(f"N={n} hi={hi} lo={lo} tc_bounds={mode}: {P / r.kernel_ms * 1e3:.4g} pf/s  kernel_ms={r.kernel_ms:.2f}  accepted={r.n_accepted / P:.4f} "
              f"idx={r.max_sharpe['global_index']}", flush=True)
