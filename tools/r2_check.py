"""Scratch diagnostics for the round-2 kernels (one GPU): tcgen05 path kernel vs the SIMT kernel and the FP64 oracle,
mcp_paths_stats vs numpy, tile-count sweep."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import mcportfolio as mcp
from bench import synthetic_inputs
from oracle import paths_np, philox_np, reference_np as ref

what = sys.argv[1] if len(sys.argv) > 1 else "parity"
mu, sigma = synthetic_inputs(16)
w = np.random.default_rng(1).dirichlet(np.ones(16))

if what == "parity":
    for n, steps, M in ((16, 12, 1000), (16, 252, 3000), (5, 40, 777), (32, 20, 500), (21, 33, 300)):
        mu_n, sigma_n = synthetic_inputs(n)
        wn = np.random.default_rng(n).dirichlet(np.ones(n))
        first, seed = 5_000_000_000, 77
        os.environ["MCP_PATHS_TC"] = "1"
        a = mcp.simulate_paths(mu_n, sigma_n, wn, M, steps, seed=seed, first_index=first, return_terminal=True)
        os.environ["MCP_PATHS_TC"] = "0"
        b = mcp.simulate_paths(mu_n, sigma_n, wn, M, steps, seed=seed, first_index=first, return_terminal=True)
        os.environ["MCP_PATHS_TC"] = "1"
        Z = philox_np.normals(first, M, steps, n, seed, "float32")
        want = paths_np.terminal_returns(mu_n, sigma_n, wn, Z)
        ea = np.abs((a["terminal"] + 1.0) / (want + 1.0) - 1).max()
        eb = np.abs((b["terminal"] + 1.0) / (want + 1.0) - 1).max()
        x = a["terminal"].astype(np.float64)
        ok = all(v == ref.var(x, al) and abs(c - ref.cvar(x, al)) < 1e-9 for al, (v, c) in a["stats"].items())
        print(f"n={n} S={steps} M={M}: tc rel err {ea:.3e}  simt rel err {eb:.3e}  stats exact {ok}", flush=True)
elif what == "perf":
    M = int(float(os.environ.get("M", 1e7)))
    for tiles in os.environ.get("TILES", "4").split(","):
        pass
    for mode in os.environ.get("MODES", "1,0").split(","):
        os.environ["MCP_PATHS_TC"] = mode
        for _ in range(3):
            o = mcp.simulate_paths(mu, sigma, w, M, 252, seed=0, return_terminal=False)
        t0 = time.perf_counter()
        o = mcp.simulate_paths(mu, sigma, w, M, 252, seed=0, return_terminal=False)
        dt = time.perf_counter() - t0
        print(f"TC={mode} wg={os.environ.get('MCP_PATHS_TC_WG', 'default')} ppt={os.environ.get('MCP_PATHS_TC_PPT', 'default')} stages={os.environ.get('MCP_PATHS_TC_STAGES', 'default')} mma={os.environ.get('MCP_PATHS_TC_MMA', 'default')}: {M * 252 / o['kernel_ms'] * 1e3:.4g} path-steps/s kernel_ms={o['kernel_ms']:.3f} "
              f"quantile_ms={o['quantile_ms']:.3f} e2e_ms={dt * 1e3:.3f} stats={o['stats']}", flush=True)
    if os.environ.get("R7"):
        os.environ["MCP_PATHS_TC"] = "1"
        for _ in range(3):
            o7 = mcp.simulate_paths(mu, sigma, w, M, 252, seed=0, return_terminal=False, philox_rounds=7)
        print(f"rounds=7 (TC): {M * 252 / o7['kernel_ms'] * 1e3:.4g} path-steps/s stats={o7['stats']}")
