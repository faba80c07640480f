"""Diagnostic: bounded tensor-core sweep, host-space vs device-space vs SIMT-only results row by row."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import mcportfolio as mcp
from bench import synthetic_inputs

n, P = 64, 70_001
mu, sigma = synthetic_inputs(n)
kw = dict(max_weights=np.full(n, 0.07), seed=11, max_tries=3, risk_free=0.03)
a = mcp.simulate_portfolios(mu, sigma, P, **kw)
a2 = mcp.simulate_portfolios(mu, sigma, P, **kw)
d = mcp.simulate_portfolios(mu, sigma, P, return_arrays="device", **kw)
d2 = mcp.simulate_portfolios(mu, sigma, P, return_arrays="device", **kw)
os.environ["MCP_LARGE_TC_BOUNDS"] = "0"
b = mcp.simulate_portfolios(mu, sigma, P, **kw)
os.environ.pop("MCP_LARGE_TC_BOUNDS")
ds, d2s = d.sharpes.cpu().numpy(), d2.sharpes.cpu().numpy()
print("n_acc", a.n_accepted, d.n_accepted, b.n_accepted)
print("host==host2", np.array_equal(a.sharpes, a2.sharpes), " dev==dev2", np.array_equal(ds, d2s))
for name, x, y in (("host vs dev", a.sharpes, ds), ("host vs simt", a.sharpes, b.sharpes), ("dev vs simt", ds, b.sharpes)):
    diff = np.nonzero(x != y)[0]
    print(f"{name}: {diff.size} rows differ of {x.size}; first {diff[:8]}, last {diff[-8:]}; max rel {np.abs(x / y - 1).max():.3e}")
acc = a.accepted.astype(bool)
pos = np.nonzero(acc)[0]
diff = np.nonzero(a.sharpes != ds)[0]
print("global indices of differing rows (host vs dev):", pos[diff][:20], "...", pos[diff][-5:])
print("weights host vs dev equal:", np.array_equal(a.weights, d.weights.cpu().numpy()), " max abs", np.abs(a.weights - d.weights.cpu().numpy()).max())
