"""Full-size sharding invariance: one sweep over the whole index range vs the merge of two arbitrary sub-ranges
(N=256: 1e10 portfolios on the tcgen05 kernel; N=16: 2^39, the per-call limit).  Prints rates and the picks."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import mcportfolio as mcp
from bench import synthetic_inputs
def best(parts, name, larger):
    key = (lambda r: (-getattr(r, name)["key"], getattr(r, name)["global_index"])) if larger else (lambda r: (getattr(r, name)["key"], getattr(r, name)["global_index"]))
    return getattr(min(parts, key=key), name)
for n, P in ((256, 10_000_000_000), (16, 1 << 39)):
    mu, sigma = synthetic_inputs(n)
    t0 = time.perf_counter()
    whole = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=3, first_index=2**40, return_arrays=False)
    t1 = time.perf_counter()
    cut = P // 3 + 12345
    parts = [mcp.simulate_portfolios(mu, sigma, b - a, risk_free=0.03, seed=3, first_index=2**40 + a, return_arrays=False) for a, b in ((0, cut), (cut, P))]
    ms, tr = best(parts, "max_sharpe", True), best(parts, "target_risk", False)
    ok = (ms["global_index"] == whole.max_sharpe["global_index"] and ms["sharpe"] == whole.max_sharpe["sharpe"]
          and tr["global_index"] == whole.target_risk["global_index"] and tr["risk"] == whole.target_risk["risk"]
          and sum(p.n_accepted for p in parts) == whole.n_accepted == P)
    print(f"N={n} P={P:.3e}: {P / (t1 - t0):.4g} pf/s wall, whole == merged halves: {ok}; max Sharpe idx {whole.max_sharpe['global_index']} "
          f"({whole.max_sharpe['sharpe']:.6f}), target idx {whole.target_risk['global_index']} (risk {whole.target_risk['risk']:.8f})")
