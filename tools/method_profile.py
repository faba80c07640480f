"""Where the time of simulate_method('CVaR', 1e6 portfolios) goes: stage by stage, host wall clock with a device sync after each."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import torch
import mcportfolio as mcp
from mcportfolio import api

n, T, P = 16, 365, 1_000_000
R = np.random.default_rng(0).standard_normal((T, n)) * 0.05


def t(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


ms, mom = t(lambda: mcp.estimate_moments(R, 52))
print(f"estimate_moments           {ms:8.3f} ms")
mu, sigma = mom
ms, r = t(lambda: mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=0, return_arrays="device"))
print(f"sweep -> device arrays     {ms:8.3f} ms   (kernel {r.kernel_ms:.3f})")
for rc in (True, False):
    ms, hv = t(lambda: mcp.historical_var_cvar(R, r.weights, 0.95, negate=True, recheck=rc))
    print(f"historical recheck={rc!s:5}  {ms:8.3f} ms   (kernel {hv['kernel_ms']:.3f})")
host = {k: api._result_empty(s, np.float32) for k, s in (("w", (P, n)), ("a", (P,)), ("b", (P,)), ("c", (P,)))}


def d2h():
    for k, src in (("w", r.weights), ("a", r.returns), ("b", r.risks), ("c", hv["cvar"])):
        torch.from_numpy(host[k]).copy_(src, non_blocking=True)
    torch.cuda.current_stream().synchronize()


ms, _ = t(d2h)
print(f"4 x D2H into pinned        {ms:8.3f} ms   ({(P * (n + 3) * 4) / ms / 1e6:.1f} GB/s)")
ms, _ = t(lambda: [api._result_empty(s, np.float32) for s in ((P, n), (P,), (P,), (P,))])
print(f"4 x _result_empty          {ms:8.3f} ms")
for m in ("CVaR", "Monte Carlo"):
    ms, o = t(lambda: mcp.simulate_method(R, m, P, annual_factor=52, risk_free=0.03, seed=0))
    print(f"simulate_method({m!r:13}) {ms:8.3f} ms")
ms, o = t(lambda: mcp.simulate_method(R, "CVaR", P, annual_factor=52, risk_free=0.03, seed=0, moments=mom))
print(f"simulate_method(CVaR, moments given) {ms:8.3f} ms")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    mcp.simulate_method(R, "CVaR", P, annual_factor=52, risk_free=0.03, seed=0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
