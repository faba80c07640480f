"""The Monte Carlo tab of the reference as one call.

``/root/reference/app.py:655-783`` is the body of ``tabs[2]``: it builds ``returns_df`` (657-668, option overlays
included), defines the ``simulation_methods`` table (671-677), estimates mu / Sigma (679-680), and then, per method,
runs the sampling + evaluation loop (682-717), materialises ``all_risks / all_returns / all_weights / all_metrics``
(719-722), plots them (726-736), picks ``opt_idx`` (738, 747) and draws the allocation (763-783).

``run_monte_carlo_tab`` replaces 679-722 and the picks for all five methods and hands back, per method, exactly the names
the plotting code below it reads, so the maintainer's change is::

    tab = mcp.app_adapter.run_monte_carlo_tab(returns_df, asset_names, annual_factor, user_rf, min_weights, max_weights)
    for method, config in simulation_methods.items():
        all_risks, all_returns, all_weights, all_metrics = tab[method].arrays()
        ...                                   # app.py:724-783 unchanged (opt_idx = config['opt_crit'](...) still works:
                                              # all_metrics holds sharpe | -var_95 | -cvar_95, as app.py:717 builds it)

mu / Sigma are estimated once on the device and shared by the five methods; every method's arrays come back through
pooled page-locked memory.  ``tests/test_app_adapter_*.py`` exec the app's own lines 671-677 and 738-747 around this
module (with a stub ``st`` / ``go``) and compare the picks with the golden vectors of the reference loop.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import api
from .overlay import option_overlay_returns

#: app.py:672-676 -- metric name and pick rule per method (the pick itself is computed by the library)
METHOD_METRIC = {"Monte Carlo": "sharpe", "VaR": "var_95", "CVaR": "cvar_95", "MPT": "sharpe", "Equal Weight": "sharpe"}


@dataclass
class MethodResult:
    """One pass of the per-method loop body: the four arrays of app.py:719-722 and the pick of app.py:747."""
    method: str
    metric: str
    all_risks: np.ndarray
    all_returns: np.ndarray
    all_weights: np.ndarray
    all_metrics: np.ndarray          # sharpe, or -var_95 / -cvar_95 (app.py:717)
    opt_idx: int
    opt_weights: np.ndarray          # all_weights[opt_idx] as float64 (app.py:765)

    def arrays(self):
        return self.all_risks, self.all_returns, self.all_weights, self.all_metrics

    def capital_allocation_line(self, user_rf, points=100):
        """app.py:738-744 (MPT only in the app): x = linspace(0, 1.3 max risk %, points), y = rf % + sharpe* x."""
        sharpe_star = self.all_metrics[self.opt_idx]
        cal_x = np.linspace(0, self.all_risks.max() * 1.3 * 100, points)
        return cal_x, user_rf * 100 + sharpe_star * cal_x


class TabResult(dict):
    """{method: MethodResult}, plus the mu / Sigma all methods shared (app.py:679-680)."""
    mean_returns: np.ndarray
    cov_matrix: np.ndarray


def build_returns_df(resampled_prices, asset_names, option_rows_dict=None):
    """app.py:657-668: per asset the option-overlay series when legs are configured, else pct_change().fillna(0);
    then one frame with NaN rows dropped."""
    import pandas as pd
    option_rows_dict = option_rows_dict or {}
    cols = {}
    for name in asset_names:
        price = resampled_prices[name]
        legs = option_rows_dict.get(name, [])
        if legs:
            cols[name] = pd.Series(option_overlay_returns(legs, np.asarray(price, dtype=np.float64)), index=getattr(price, "index", None))
        else:
            cols[name] = price.pct_change().fillna(0)
    return pd.DataFrame(cols).dropna()


def run_monte_carlo_tab(returns_df, asset_names, annual_factor, user_rf, min_weights, max_weights, n_portfolios=2500, *,
                        methods=api.METHODS, seed=0, dtype="float32", device=None, weights=None, cvar_alpha=0.95):
    """{method: MethodResult} for the methods of app.py:671-677, in the app's order.

    returns_df      the app's frame (T x N, leading fillna(0) row included) or a (T, N) array
    asset_names     column order (app.py:686: 1/len(asset_names)); checked against the frame
    annual_factor, user_rf, min_weights, max_weights: the app's globals, passed through untouched (user_rf is
                    subtracted raw, app.py:711)
    n_portfolios    app.py:681 (2500)
    seed            Philox key; every method draws the SAME portfolios for a given seed (the app re-draws per method from
                    the global MT19937 stream)
    weights         optional (P, N) rows evaluated instead of in-kernel draws -- e.g. the app's own seeded numpy draws,
                    which reproduces its arrays and picks
    """
    cols = list(getattr(returns_df, "columns", []))
    names = list(asset_names)
    if cols and cols != names:
        returns_df = returns_df[names]                      # the app's frame is built in asset_names order (658-667)
    R = np.ascontiguousarray(np.asarray(returns_df, dtype=np.float64))
    if R.ndim != 2 or R.shape[1] != len(names):
        raise ValueError(f"returns_df must be (T, {len(names)}), got {R.shape}")
    lo = None if min_weights is None else np.asarray(min_weights, dtype=np.float64)
    hi = None if max_weights is None else np.asarray(max_weights, dtype=np.float64)
    moments = api.estimate_moments(R, annual_factor, device=device)               # app.py:679-680, once for all methods
    out = TabResult()
    out.mean_returns, out.cov_matrix = moments
    for method in methods:
        if method not in METHOD_METRIC:
            raise KeyError(method)
        try:
            o = api.simulate_method(R, method, n_portfolios, annual_factor=annual_factor, risk_free=user_rf, min_weights=lo,
                                    max_weights=hi, alpha=cvar_alpha, seed=seed, dtype=dtype, device=device, weights=weights,
                                    moments=moments)
        except IndexError:
            if method != "Equal Weight":
                raise
            # app.py:687: 1/N outside the bounds leaves the four arrays empty; the app then fails at all_risks[opt_idx]
            # (747-749).  The arrays are returned as the app builds them; opt_idx stays 0 (config['opt_crit'] returns 0).
            n = len(names)
            e = np.empty(0, dtype=np.float64)
            out[method] = MethodResult(method, METHOD_METRIC[method], e, e.copy(), np.empty((0, n)), e.copy(), 0, np.empty(0))
            continue
        out[method] = MethodResult(method, METHOD_METRIC[method], o["risks"], o["returns"], o["weights"], o["metrics"],
                                   int(o["opt_idx"]), o["opt_weights"])
    return out
