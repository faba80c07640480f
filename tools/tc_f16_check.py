"""A/B of the two operand splits of the tcgen05 sweep (FP16 x2 vs TF32 + BF16) at N = 256: rate and agreement."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import mcportfolio as mcp
from bench import synthetic_inputs

for n in (256, 64):
    mu, sigma = synthetic_inputs(n)
    out = {}
    for mode in ("1", "0"):
        os.environ["MCP_LARGE_TC_F16"] = mode
        r = mcp.simulate_portfolios(mu, sigma, 20000, risk_free=0.03, seed=0, dtype="float32")
        P = 200_000_000 if n == 256 else 400_000_000
        best = 1e9
        for _ in range(3):
            b = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=0, return_arrays=False)
            best = min(best, b.kernel_ms)
        out[mode] = (r, b)
        print(f"N={n} f16={mode}: {P / best * 1e3:.4g} pf/s  picks {b.max_sharpe['global_index']} {b.target_risk['global_index']}", flush=True)
    a, t = out["1"][0], out["0"][0]
    print("  risk rel diff f16 vs tf32:", np.abs(a.risks / t.risks - 1).max(), " sharpe:", np.abs(a.sharpes - t.sharpes).max(),
          " weights equal:", np.array_equal(a.weights, t.weights))
