"""Quick on-box perf probe (not the bench): prints kernel-level rates as JSON lines."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import torch
import mcportfolio as mcp


def synthetic(n, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    return rng.uniform(0.05, 0.60, n) if False else None, A


def inputs(n):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((n, n))
    sigma = A @ A.T / n * 0.2 + 1e-6 * np.eye(n)
    mu = rng.uniform(0.05, 0.60, n)
    return mu, sigma


def main():
    eng = mcp.get_engine(0)
    out = {"device": eng.info()}
    out["fma_peak_f32_tflops"] = eng.measure_fma_peak("float32")
    out["fma_peak_f64_tflops"] = eng.measure_fma_peak("float64")
    out["fma_peak_f32x2_tflops"] = eng.measure_fma_peak("float32x2")
    print(json.dumps(out), flush=True)
    for n in (16, 8, 32, 14):
        mu, sigma = inputs(n)
        for dtype in ("float32", "float64"):
            P = 2_000_000_000 if dtype == "float32" else 200_000_000
            if n == 32:
                P //= 4
            mcp.simulate_portfolios(mu, sigma, P // 20, risk_free=0.03, return_arrays=False, dtype=dtype)
            best = None
            for _ in range(3):
                r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, return_arrays=False, dtype=dtype)
                best = r.kernel_ms if best is None else min(best, r.kernel_ms)
            print(json.dumps({"kernel": "sweep_rng_nowrite", "n": n, "dtype": dtype, "P": P, "ms": best,
                              "pf_per_s": P / best * 1e3, "idx": r.max_sharpe["global_index"]}), flush=True)
    mu, sigma = inputs(16)
    # write-back, device space
    P = 50_000_000
    for _ in range(3):
        r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, return_arrays="device")
    print(json.dumps({"kernel": "sweep_rng_write_device", "n": 16, "P": P, "ms": r.kernel_ms, "pf_per_s": P / r.kernel_ms * 1e3,
                      "GBps": P * 76 / r.kernel_ms / 1e6}), flush=True)
    W = r.weights
    for _ in range(3):
        r2 = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03)
    print(json.dumps({"kernel": "sweep_supplied_device", "n": 16, "P": P, "ms": r2.kernel_ms, "pf_per_s": P / r2.kernel_ms * 1e3,
                      "GBps": P * (64 + 76) / r2.kernel_ms / 1e6}), flush=True)
    del r, r2, W
    torch.cuda.empty_cache()
    # host space end to end
    P = 4_000_000
    t0 = time.perf_counter()
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03)
    t1 = time.perf_counter()
    print(json.dumps({"kernel": "sweep_rng_write_host_e2e", "P": P, "wall_ms": (t1 - t0) * 1e3, "kernel_ms": r.kernel_ms}), flush=True)
    # paths
    w = np.full(16, 1 / 16)
    for dtype, M in (("float32", 2_000_000), ("float64", 200_000)):
        mcp.simulate_paths(mu, sigma, w, M // 10, 252, dtype=dtype, return_terminal=False)
        o = mcp.simulate_paths(mu, sigma, w, M, 252, dtype=dtype, return_terminal=False)
        print(json.dumps({"kernel": "paths_rng", "dtype": dtype, "M": M, "ms": o["kernel_ms"],
                          "pathsteps_per_s": M * 252 / o["kernel_ms"] * 1e3, "quantile_ms": o["quantile_ms"],
                          "stats": {str(k): v for k, v in o["stats"].items()}}), flush=True)
    x = torch.randn(10_000_000, device="cuda")
    for _ in range(3):
        t0 = time.perf_counter()
        mcp.quantile_stats(x, (0.95, 0.99))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    print(json.dumps({"kernel": "quantiles_1e7", "wall_ms": (t1 - t0) * 1e3, "kernel_ms": eng.last_kernel_ms()}), flush=True)


if __name__ == "__main__":
    main()
