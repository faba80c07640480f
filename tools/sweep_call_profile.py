"""Host-side cost of one simulate_portfolios(return_arrays='device') call of 1e6 portfolios: cProfile by function."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np, torch
import mcportfolio as mcp
from bench import synthetic_inputs
mu, sigma = synthetic_inputs(16)
P = 1_000_000
f = lambda: mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=0, return_arrays="device")
for _ in range(5): r = f()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): r = f()
torch.cuda.synchronize()
print(f"per call {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms, kernel {r.kernel_ms:.3f}, launches/call {mcp.get_engine().launch_count()/55:.1f}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): r = f()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
