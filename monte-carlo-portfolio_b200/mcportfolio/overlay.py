"""Option / hedging overlay: per-period return series of a leg basket.

Host-side mirror of ``calc_option_return`` / ``calc_options_series`` (app.py:164-193): when the
user configures option legs for an asset, this series replaces the asset's ``pct_change`` column
of the returns frame that feeds the Monte Carlo tab (app.py:661-664).  It is O(T x legs) scalar
arithmetic that runs once per asset per session, so it stays on the host -- vectorised here
instead of the reference's Python double loop.

A leg is ``(kind, strike, premium, qty)`` exactly as the app stores it; ``kind`` is one of the
app's seven Persian tags or the English alias below.  Unknown kinds contribute 0 (app.py:179-180).
"""
from __future__ import annotations

import numpy as np

KINDS = {
    "خرید دارایی": "long_asset", "فروش دارایی": "short_asset",
    "خرید کال": "long_call", "فروش کال": "short_call",
    "خرید پوت": "long_put", "فروش پوت": "short_put",
    "فروش فیوچرز": "short_future",
}
_ALIASES = set(KINDS.values())


def _kind(tag: str) -> str | None:
    k = KINDS.get(tag, tag)
    return k if k in _ALIASES else None


def option_overlay_returns(option_rows, prices) -> np.ndarray:
    """``calc_options_series(option_rows, prices)`` (app.py:182-193) as a float64 array.

    rets[0] = 0; rets[i] = sum_legs qty * leg_return(price_i, price_{i-1}, strike, premium)
    with every leg return divided by the previous price (0 when that price is 0).
    """
    p = np.asarray(prices, dtype=np.float64)
    if p.ndim != 1:
        raise ValueError("prices must be a 1-D series")
    rets = np.zeros(p.shape[0])
    if p.shape[0] < 2:
        return rets
    prev, cur = p[:-1], p[1:]
    ok = prev != 0
    safe = np.where(ok, prev, 1.0)
    total = np.zeros(cur.shape[0])
    for tag, strike, premium, qty in option_rows:
        kind = _kind(tag)
        if kind is None:
            continue
        if kind == "long_asset":
            num = cur - prev
        elif kind in ("short_asset", "short_future"):
            num = prev - cur
        elif kind == "long_call":
            num = np.maximum(cur - strike, 0) - premium
        elif kind == "short_call":
            num = premium - np.maximum(cur - strike, 0)
        elif kind == "long_put":
            num = np.maximum(strike - cur, 0) - premium
        else:  # short_put
            num = premium - np.maximum(strike - cur, 0)
        total += qty * np.where(ok, num / safe, 0.0)
    rets[1:] = total
    return rets


def returns_with_overlays(prices_frame, overlays=None):
    """The returns frame of app.py:658-667: per asset either the overlay series (when legs are
    configured) or ``pct_change().fillna(0)``; rows with NaN dropped.  `prices_frame` is the
    resampled price DataFrame (`mcportfolio.data.price_frame`), `overlays` maps asset name -> legs."""
    import pandas as pd
    overlays = overlays or {}
    cols = {}
    for name in prices_frame.columns:
        legs = overlays.get(name)
        if legs:
            cols[name] = pd.Series(option_overlay_returns(legs, prices_frame[name].to_numpy()), index=prices_frame.index)
        else:
            cols[name] = prices_frame[name].pct_change().fillna(0)
    return pd.DataFrame(cols).dropna()
