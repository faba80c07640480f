"""numpy FP64 oracle for the correlated-path simulator, VaR/CVaR over paths and
the frontier envelope.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

**Parity unpinned**: none of this exists in ``/root/reference/app.py`` (no
Cholesky, no normal draws, no simulated paths; SURVEY.md section 8 rows a10,
a12 and a11-on-paths).  The spec is BASELINE.json's north_star text, fixed
here with the reference's own conventions:

* volatility / covariance definition      app.py:709  (sigma = sqrt(w' Sigma w), Sigma annualised, 680)
* arithmetic compounding ``prod(1 + r)``  app.py:253 (``np.cumprod(1+returns)``), 249
* VaR = percentile((1-alpha)*100), linear; CVaR = mean(x[x <= VaR])   app.py:258-263
* ``np.linalg.cholesky``: lower-triangular L, Sigma = L L'
"""
from __future__ import annotations

import numpy as np

from .reference_np import cvar, var


def step_returns(mu, sigma, normals, dt=1.0 / 252):
    """r[m, s, :] = mu*dt + sqrt(dt) * L z[m, s, :],  L = chol(Sigma_annual)."""
    mu = np.asarray(mu, dtype=np.float64)
    L = np.linalg.cholesky(np.asarray(sigma, dtype=np.float64))
    Z = np.asarray(normals, dtype=np.float64)
    return mu * dt + np.sqrt(dt) * (Z @ L.T)


def terminal_returns(mu, sigma, weights, normals, dt=1.0 / 252):
    """x[m] = w . prod_s(1 + r[m, s, :]) - 1  (per-asset cumulative product, then the portfolio)."""
    r = step_returns(mu, sigma, normals, dt)
    V = np.cumprod(1.0 + r, axis=1)[:, -1, :]
    return V @ np.asarray(weights, dtype=np.float64) - 1.0


def risk_stats(x, alphas=(0.95, 0.99)):
    """{alpha: (VaR, CVaR)} with app.py:258-263 conventions (signed lower-tail return)."""
    x = np.asarray(x, dtype=np.float64)
    return {a: (var(x, a), cvar(x, a)) for a in alphas}


def envelope(risks, returns, n_bins, lo, hi):
    """Frontier envelope: per risk bin the max return and its first (lowest) index.

    bin = floor((risk - lo) * n_bins / (hi - lo)); risk == hi goes to the last bin;
    risks outside [lo, hi] are ignored.  Empty bins: return -inf, index -1.
    Replaces the reference's per-point scatter (app.py:726-736) at sizes where a
    scatter is meaningless (C5).
    """
    risks = np.asarray(risks, dtype=np.float64)
    returns = np.asarray(returns, dtype=np.float64)
    scale = n_bins / (hi - lo)
    b = np.floor((risks - lo) * scale).astype(np.int64)
    b[risks == hi] = n_bins - 1
    ok = (risks >= lo) & (risks <= hi) & (b >= 0) & (b < n_bins)
    best = np.full(n_bins, -np.inf)
    best_idx = np.full(n_bins, -1, dtype=np.int64)
    for i in np.nonzero(ok)[0]:
        k = b[i]
        if returns[i] > best[k]:
            best[k] = returns[i]
            best_idx[k] = i
    return best, best_idx
