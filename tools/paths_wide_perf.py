"""Throughput of the path kernels by universe size: tcgen05 (N <= 32; 32 < N <= 128 with chunked normals) and the SIMT warp-per-path
kernel (N > 128, or MCP_PATHS_TC=0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import mcportfolio as mcp
from bench import synthetic_inputs

for n, M, tc in ((16, 4_000_000, "1"), (32, 2_000_000, "1"), (64, 1_000_000, "1"), (64, 1_000_000, "32"), (64, 200_000, "0"),
                 (128, 400_000, "1"), (128, 400_000, "32"), (128, 100_000, "0"), (192, 400_000, "1"), (192, 400_000, "32"),
                 (256, 400_000, "1"), (256, 400_000, "32"), (256, 100_000, "0")):
    os.environ["MCP_PATHS_TC"] = "0" if tc == "0" else "1"                  # tc: 1 = tcgen05 (16-bit split, two stages above N = 32),
    os.environ["MCP_PATHS_TC_WIDE16"] = "0" if tc == "32" else "1"          # 32 = the one-stage TF32-split wide kernel, 0 = SIMT
    mu, sigma = synthetic_inputs(n)
    w = np.full(n, 1 / n)
    for _ in range(2):
        o = mcp.simulate_paths(mu, sigma, w, M, 252, seed=0, return_terminal=False)
    flop = n * n + 3 * n
    print(f"N={n:4d} tc={tc} M={M:.0e}: {M * 252 / o['kernel_ms'] * 1e3:.4g} path-steps/s  kernel_ms={o['kernel_ms']:.2f}  "
          f"{M * 252 * flop / o['kernel_ms'] * 1e3 / 1e12:.2f} TFLOP/s algorithmic  VaR95={o['stats'][0.95][0]:.5f}", flush=True)
