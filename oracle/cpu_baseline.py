"""CPU baselines timed by bench.py (``cpu_baseline`` leg and ``--impl reference``).

TEST / MEASUREMENT INFRASTRUCTURE (see ``oracle/__init__.py``): never on the product path.

* ``verbatim_rate``   -- the reference's own per-portfolio loop semantics (app.py:699-717:
  one legacy Dirichlet call per portfolio, pandas operands, per-portfolio var / cvar), 1 core.
* ``vectorised_rate`` -- "host numpy": the vectorised restatement (oracle.reference_np.
  vectorised_sweep) run as one process per host core, the strongest numpy form of the path.
* ``paths_rate``      -- numpy path simulator (standard_normal @ L.T, cumprod, percentile).
"""
from __future__ import annotations

import os
import time

import numpy as np


def _worker_sweep(args):
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    mu, sigma, n, seed, chunk = args
    from oracle.reference_np import vectorised_sweep
    t0 = time.perf_counter()
    res = vectorised_sweep(mu, sigma, n, 0.03, 0.30, seed=seed, chunk=chunk)
    return n, time.perf_counter() - t0, res


def _worker_paths(args):
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    mu, sigma, w, m, steps, seed = args
    from oracle.paths_np import risk_stats, terminal_returns
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    done = 0
    xs = []
    while done < m:
        k = min(10_000, m - done)
        Z = rng.standard_normal((k, steps, len(mu)))
        xs.append(terminal_returns(mu, sigma, w, Z))
        done += k
    st = risk_stats(np.concatenate(xs))
    return m, time.perf_counter() - t0, st


def _pool(cores):
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    return ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn"))


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def vectorised_rate(mu, sigma, per_core: int, cores: int | None = None, chunk: int = 250_000, pool=None):
    """portfolios/s of the vectorised numpy path with one process per core; wall-clock over the
    whole pool (includes nothing but the sweeps: workers are warmed up first)."""
    cores = cores or host_cores()
    own = pool is None
    pool = pool or _pool(cores)
    try:
        list(pool.map(_worker_sweep, [(mu, sigma, 1000, 0, 1000)] * cores))          # spawn + import warm-up
        t0 = time.perf_counter()
        out = list(pool.map(_worker_sweep, [(mu, sigma, per_core, 100 + i, chunk) for i in range(cores)]))
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.shutdown()
    total = sum(o[0] for o in out)
    return {"value": total / wall, "unit": "portfolios/s", "cores": cores, "kind": "port",
            "sample": f"{total} portfolios (N={len(mu)}), vectorised numpy (oracle.reference_np.vectorised_sweep), "
                      f"{cores} processes x {per_core}, chunk {chunk}, wall {wall:.2f} s"}


def verbatim_rate(n_portfolios: int = 2000):
    """portfolios/s of the reference's loop semantics (pandas operands, per-portfolio var/cvar)."""
    import pandas as pd
    from oracle.reference_np import cvar, var
    rng = np.random.default_rng(0)
    T, N = 365, 16
    returns_df = pd.DataFrame(rng.standard_normal((T, N)) * 0.05)
    mean_returns = returns_df.mean() * 52
    cov_matrix = returns_df.cov() * 52
    lo, hi = np.zeros(N), np.ones(N)
    np.random.seed(0)
    t0 = time.perf_counter()
    acc = []
    for _ in range(n_portfolios):                       # app.py:699-717, statement for statement
        for _ in range(100):
            ws = np.random.dirichlet(np.ones(N), size=1)[0]
            if np.all(ws >= lo) and np.all(ws <= hi):
                break
        port_return = np.dot(ws, mean_returns)
        port_std = np.sqrt(np.dot(ws.T, np.dot(cov_matrix, ws)))
        port_series = returns_df @ ws
        sharpe = (port_return - 0.03) / port_std if port_std > 0 else 0
        acc.append((port_std, port_return, sharpe, var(port_series, 0.95), cvar(port_series, 0.95)))
    wall = time.perf_counter() - t0
    return {"value": n_portfolios / wall, "unit": "portfolios/s", "cores": 1,
            "sample": f"{n_portfolios} portfolios, reference loop semantics (app.py:699-717) with pandas operands"}


def paths_rate(mu, sigma, w, per_core: int, steps: int = 252, cores: int | None = None, pool=None):
    cores = cores or host_cores()
    own = pool is None
    pool = pool or _pool(cores)
    try:
        list(pool.map(_worker_paths, [(mu, sigma, w, 10, 4, 0)] * cores))
        t0 = time.perf_counter()
        out = list(pool.map(_worker_paths, [(mu, sigma, w, per_core, steps, 7 + i) for i in range(cores)]))
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.shutdown()
    total = sum(o[0] for o in out)
    return {"value": total * steps / wall, "unit": "path-steps/s", "cores": cores, "kind": "port",
            "sample": f"{total} paths x {steps} steps (N={len(mu)}), numpy (oracle.paths_np), {cores} processes, wall {wall:.2f} s"}
