"""ncu target (one GPU): the four headline kernels at the launch sizes bench.py uses, two launches each.

    PROFILE_ONLY=sweep|paths|tc|tcb|hist  restricts the run to one kernel family (default: all five).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import torch
import mcportfolio as mcp
from bench import synthetic_inputs

only = os.environ.get("PROFILE_ONLY", "")
mu, sigma = synthetic_inputs(16)
w = np.full(16, 1 / 16)
if only in ("", "sweep"):
    P = int(float(os.environ.get("PROFILE_P", 1e10)))
    for _ in range(2):
        r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=0, return_arrays=False)
    print(f"sweep N=16 P={P:.0e}: {P / r.kernel_ms * 1e3:.4g} pf/s kernel_ms={r.kernel_ms:.3f}")
if only in ("", "paths"):
    M = int(float(os.environ.get("PROFILE_M", 1e7)))
    for _ in range(2):
        o = mcp.simulate_paths(mu, sigma, w, M, 252, seed=0, return_terminal=False)
    print(f"paths M={M:.0e}: {M * 252 / o['kernel_ms'] * 1e3:.4g} path-steps/s kernel_ms={o['kernel_ms']:.3f}")
if only in ("", "tc"):
    mu256, sigma256 = synthetic_inputs(256)
    P256 = int(float(os.environ.get("PROFILE_P256", 2e7)))
    for _ in range(2):
        r256 = mcp.simulate_portfolios(mu256, sigma256, P256, risk_free=0.03, seed=0, return_arrays=False)
    print(f"sweep N=256 P={P256:.0e}: {P256 / r256.kernel_ms * 1e3:.4g} pf/s kernel_ms={r256.kernel_ms:.3f}")
if only in ("", "tcb"):
    mu256, sigma256 = synthetic_inputs(256)
    Pb = int(float(os.environ.get("PROFILE_P256B", 2e7)))
    for _ in range(2):
        rb = mcp.simulate_portfolios(mu256, sigma256, Pb, risk_free=0.03, seed=0, return_arrays=False, max_weights=np.full(256, 0.03))
    print(f"bounded sweep N=256 hi=0.03 P={Pb:.0e}: {Pb / rb.kernel_ms * 1e3:.4g} pf/s kernel_ms={rb.kernel_ms:.3f} accepted={rb.n_accepted}")
if only in ("", "hist"):
    rng = np.random.default_rng(0)
    T, n, Ph = 365, 16, 1_000_000
    R = rng.standard_normal((T, n)) * 0.05
    W = torch.from_numpy(rng.dirichlet(np.ones(n), size=Ph).astype(np.float32)).cuda()
    for _ in range(2):
        hv = mcp.historical_var_cvar(R, W, 0.95, dtype="float32", return_arrays=False)
    print(f"hist T={T} P={Ph:.0e}: {Ph / hv['kernel_ms'] * 1e3:.4g} pf/s kernel_ms={hv['kernel_ms']:.3f}")
