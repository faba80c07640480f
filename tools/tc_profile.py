"""ncu target: the tcgen05 N=256 sweep only (a few launches, one GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import mcportfolio as mcp
from bench import synthetic_inputs

n = int(os.environ.get("PROFILE_N", 256))
P = int(os.environ.get("PROFILE_P", 4_000_000))
mu, sigma = synthetic_inputs(n)
for _ in range(3):
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=0, return_arrays=False)
    print(f"sweep N={n} P={P}: {P / r.kernel_ms * 1e3:.4g} pf/s  kernel_ms={r.kernel_ms:.3f}")
