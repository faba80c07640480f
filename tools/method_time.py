import os, sys, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "monte-carlo-portfolio_b200")]
import numpy as np, torch
import mcportfolio as mcp
n, Th, Ph = 16, 365, 1_000_000
Rh = np.random.default_rng(0).standard_normal((Th, n)) * 0.05
for m in ("CVaR", "VaR"):
    for _ in range(3):
        mo = mcp.simulate_method(Rh, m, Ph, annual_factor=52, risk_free=0.03, seed=0)
    t0 = time.perf_counter()
    for _ in range(10):
        mo = mcp.simulate_method(Rh, m, Ph, annual_factor=52, risk_free=0.03, seed=0)
    dt = (time.perf_counter() - t0) / 10
    print(m, f"{dt*1e3:.3f} ms  {Ph/dt:.4g} pf/s  opt={mo['opt_idx']}", mo["metrics"][:3], mo["weights"][mo["opt_idx"]][:3])
