import sys, os; sys.path[:0]=["/root/repo","/root/repo/monte-carlo-portfolio_b200"]
import numpy as np, mcportfolio as mcp
from bench import synthetic_inputs
mu,sigma=synthetic_inputs(16)
P=2_000_000_000
for _ in range(3):
    r=mcp.simulate_portfolios(mu,sigma,P,risk_free=0.03,return_arrays=False)
print("K env", os.environ.get("MCP_SWEEP_K"), "pf/s", P/r.kernel_ms*1e3, r.max_sharpe["global_index"])
