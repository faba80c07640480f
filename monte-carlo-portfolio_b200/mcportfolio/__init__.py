"""mcportfolio: B200-native Monte Carlo portfolio hot path (drop-in for the path in
mohammadmarghzari/monte-carlo-portfolio's app.py).  All compute is in libmcp.so
(sm_100a CUDA) behind the C ABI of include/mcp.h; see DESIGN.md."""
from ._lib import McpError, build, lib
from . import app_adapter
from .api import (METHODS, asset_stats, Engine, PortfolioResult, efficient_frontier, envelope_from_arrays, frontier_envelope, get_engine,
                  comm_unique_id, estimate_moments, historical_var_cvar,
                  pinned_empty, quantile_stats, simulate_method, simulate_paths, simulate_portfolios)

__all__ = ["app_adapter", "asset_stats", "comm_unique_id", "estimate_moments", "Engine", "McpError", "PortfolioResult", "build", "efficient_frontier", "envelope_from_arrays", "frontier_envelope", "get_engine",
           "historical_var_cvar", "lib", "pinned_empty", "simulate_method", "METHODS", "quantile_stats", "simulate_paths", "simulate_portfolios"]
