"""numpy FP64 restatement of the reference's Monte Carlo portfolio path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
``/root/reference/app.py`` lines it follows.  The reference is a Streamlit
script whose hot loop is inline code, so the loop is restated here as
functions; the reference's own callable pieces (``var``, ``cvar``,
``efficient_frontier``) are additionally executed verbatim by
``oracle/ref_loader.py`` to pin this file (``tests/golden``).

Parity: pinned by golden vectors generated from the reference's own lines for
a1-a8, a11, a13; **parity unpinned** for a9 (30 %-risk pick, README.md:4 only).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# a11  risk statistics                                    app.py:258-263
# --------------------------------------------------------------------------

def lower_tail_percent(alpha: float) -> float:
    """``(1-alpha)*100`` exactly as written at app.py:259 (FP64: 0.95 -> 5.000000000000004)."""
    return (1 - alpha) * 100


def percentile_linear(x: np.ndarray, q: float) -> float:
    """``np.percentile(x, q)`` with the default ``method='linear'`` spelled out.

    h = q/100 * (n-1); v = x[floor h] + (h - floor h) * (x[floor h + 1] - x[floor h])
    on the sorted values.  This is the third-party arithmetic behind app.py:259
    (numpy 2.3.5 in this image; requirements.txt pins no version).  numpy's own
    lerp switches form at t >= 0.5 (``b - (b-a)*(1-t)``), reproduced here so the
    result is bit-equal to ``np.percentile``.
    """
    xs = np.sort(np.asarray(x, dtype=np.float64))
    n = xs.shape[0]
    h = (q / 100.0) * (n - 1)
    lo = int(np.floor(h))
    hi = min(lo + 1, n - 1)
    t = h - lo
    a, b = xs[lo], xs[hi]
    if t >= 0.5:
        return float(b - (b - a) * (1 - t))
    return float(a + (b - a) * t)


def var(returns, alpha: float = 0.95) -> float:
    """app.py:258-259."""
    return float(np.percentile(returns, lower_tail_percent(alpha)))


def cvar(returns, alpha: float = 0.95) -> float:
    """app.py:261-263: mean of the values <= VaR (inclusive); VaR itself if none."""
    returns = np.asarray(returns)
    v = var(returns, alpha)
    mask = returns <= v
    return float(returns[mask].mean()) if np.any(mask) else v


# --------------------------------------------------------------------------
# f2  per-asset statistics                                app.py:231-256, 286-335
# --------------------------------------------------------------------------

def asset_stats(returns, risk_free=0.0, ann_factor=12, alpha=0.95):
    """The statistics `calc_asset_stats` (app.py:286-335) derives from one returns series."""
    r = np.asarray(returns, dtype=np.float64)
    excess = r - risk_free / ann_factor                                   # app.py:232
    std = np.std(r, ddof=1) if len(r) > 1 else np.nan
    sharpe = 0.0 if np.std(excess, ddof=1) == 0 else np.mean(excess) / np.std(excess, ddof=1) * np.sqrt(ann_factor)
    neg = excess[excess < 0]                                              # app.py:242-243
    down = np.std(neg, ddof=1) if len(neg) > 0 else 0.0001
    cum = np.cumprod(1 + r)                                               # app.py:253-256
    peak = np.maximum.accumulate(cum)
    return {"sharpe": sharpe, "sortino": np.mean(excess) / down * np.sqrt(ann_factor),
            "volatility_ann": std * np.sqrt(ann_factor),
            "total_return_ann": np.prod(1 + r) ** (ann_factor / len(r)) - 1,      # app.py:249
            "mean_ann": np.mean(r) * ann_factor, "mean_month": np.mean(r), "std_month": std,
            "min_month": np.min(r), "max_month": np.max(r), "max_drawdown": np.min((cum - peak) / peak),
            "var_95": var(r, alpha), "cvar_95": cvar(r, alpha)}


# --------------------------------------------------------------------------
# a1  mu / Sigma estimation                               app.py:658-667, 679-680
# --------------------------------------------------------------------------

def pct_change_fillna0(prices: np.ndarray) -> np.ndarray:
    """app.py:666 ``price.pct_change().fillna(0)``: first row is 0 and is kept."""
    prices = np.asarray(prices, dtype=np.float64)
    r = np.zeros_like(prices)
    r[1:] = prices[1:] / prices[:-1] - 1.0
    return r


def estimate_mu_sigma(returns: np.ndarray, annual_factor: float):
    """app.py:679-680: ``mean()*A`` and ``cov()*A`` (pandas cov: ddof=1)."""
    returns = np.asarray(returns, dtype=np.float64)
    mu = returns.mean(axis=0) * annual_factor
    sigma = np.atleast_2d(np.cov(returns, rowvar=False, ddof=1)) * annual_factor
    return mu, sigma


# --------------------------------------------------------------------------
# a4-a6  return / volatility / Sharpe                     app.py:708-711 (twin 278-282)
# --------------------------------------------------------------------------

def portfolio_metrics(weights, mu, sigma, risk_free: float = 0.0):
    """Vectorised app.py:708-711 for a (P, N) weight matrix.

    ret = w.mu (708); risk = sqrt(w^T (Sigma w)) (709);
    sharpe = (ret - risk_free)/risk if risk > 0 else 0 (711) -- ``risk_free`` is
    subtracted raw (the app passes the widget value 3.0 unchanged, app.py:428).
    ``risk_free=0`` gives the dead-code twin's ``ret/std`` (app.py:282).
    """
    W = np.atleast_2d(np.asarray(weights, dtype=np.float64))
    mu = np.asarray(mu, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    ret = W @ mu
    risk = np.sqrt(np.einsum("pi,pi->p", W @ sigma.T, W))
    with np.errstate(divide="ignore", invalid="ignore"):
        sharpe = np.where(risk > 0, (ret - risk_free) / np.where(risk > 0, risk, 1.0), 0.0)
    return ret, risk, sharpe


# --------------------------------------------------------------------------
# a2-a3  weight sampling + bounds rejection               app.py:699-707 (twin 269-277)
# --------------------------------------------------------------------------

def draw_weights_reference(n_assets, n_portfolios, min_weights=None, max_weights=None,
                           keep_last=False, max_tries=100, rng=None):
    """The reference's sampling loop, one legacy-RNG call per try.

    ``np.random.dirichlet(np.ones(N), size=1)[0]`` per try (app.py:702 / 271), at
    most 100 tries (701 / 270).  Inline loop: a portfolio with no valid draw is
    skipped (706-707) -> ``valid[i] = False``.  Dead-code twin
    (``keep_last=True``): the last draw is kept even if invalid (277).
    ``rng`` defaults to the global legacy ``np.random`` state, as in the app.
    """
    rng = np.random if rng is None else rng
    lo = None if min_weights is None else np.asarray(min_weights, dtype=np.float64)
    hi = None if max_weights is None else np.asarray(max_weights, dtype=np.float64)
    W = np.empty((n_portfolios, n_assets))
    valid = np.zeros(n_portfolios, dtype=bool)
    ones = np.ones(n_assets)
    for i in range(n_portfolios):
        ok = False
        for _ in range(max_tries):
            w = rng.dirichlet(ones, size=1)[0]
            if lo is not None and not np.all(w >= lo):
                continue
            if hi is not None and not np.all(w <= hi):
                continue
            ok = True
            break
        W[i] = w
        valid[i] = ok or keep_last
    return W, valid


def efficient_frontier(mean_returns, cov_matrix, points=200, min_weights=None, max_weights=None,
                       rng=None):
    """Restatement of the dead-code twin app.py:265-284 (same return layout).

    results[0]=std, [1]=ret, [2]=ret/std (no risk-free); weights (points, N).
    """
    mu = np.asarray(mean_returns, dtype=np.float64)
    W, _ = draw_weights_reference(len(mu), points, min_weights, max_weights,
                                  keep_last=True, rng=rng)
    ret, risk, sharpe = portfolio_metrics(W, mu, cov_matrix, 0.0)
    return np.vstack([risk, ret, sharpe]), W


def inline_loop(mean_returns, cov_matrix, n_portfolios, risk_free, min_weights, max_weights,
                returns_matrix=None, metric="sharpe", alpha=0.95, rng=None):
    """Restatement of one method's pass of the inline loop app.py:699-722.

    Returns ``all_risks, all_returns, all_weights, all_metrics`` (719-722) over the
    accepted portfolios only.  ``metric``: 'sharpe' | 'var_95' | 'cvar_95' (717);
    the VaR/CVaR metrics are ``-var`` / ``-cvar`` of ``returns_matrix @ w`` (710-713).
    """
    mu = np.asarray(mean_returns, dtype=np.float64)
    W, valid = draw_weights_reference(len(mu), n_portfolios, min_weights, max_weights,
                                      keep_last=False, rng=rng)
    W = W[valid]
    ret, risk, sharpe = portfolio_metrics(W, mu, cov_matrix, risk_free)
    if metric == "sharpe":
        metrics = sharpe
    else:
        hv, hc = historical_var_cvar(returns_matrix, W, alpha)
        metrics = -hv if metric == "var_95" else -hc
    return risk, ret, W, metrics


def historical_var_cvar(returns_matrix, weights, alpha=0.95):
    """Per-portfolio VaR/CVaR of the historical series ``R @ w`` (app.py:710-713)."""
    R = np.asarray(returns_matrix, dtype=np.float64)
    W = np.atleast_2d(np.asarray(weights, dtype=np.float64))
    series = R @ W.T                                   # (T, P)
    q = lower_tail_percent(alpha)
    v = np.percentile(series, q, axis=0)
    c = np.empty_like(v)
    for p in range(W.shape[0]):
        col = series[:, p]
        m = col <= v[p]
        c[p] = col[m].mean() if m.any() else v[p]
    return v, c


# --------------------------------------------------------------------------
# a8, a9, a13  selection                                  app.py:672-676, 738, 747
# --------------------------------------------------------------------------

def select_max_sharpe(sharpes) -> int:
    """app.py:672 ``np.argmax(x['sharpe'])`` -- first occurrence wins ties."""
    return int(np.argmax(sharpes))


def select_target_risk(risks, target: float = 0.30) -> int:
    """30 %-risk pick.  NOT IN THE REFERENCE CODE (README.md:4 only) -- parity unpinned.

    Spec (BASELINE.json north_star): argmin_i |risk_i - target|, first occurrence.
    """
    return int(np.argmin(np.abs(np.asarray(risks) - target)))


def select_method(metrics, method: str) -> int:
    """app.py:672-676 as called at 747: every key of the dict maps to ``all_metrics``.

    'Monte Carlo'/'MPT' -> argmax(metric); 'VaR'/'CVaR' -> argmin(metric) where the
    metric already holds ``-var_95`` / ``-cvar_95`` (717); 'Equal Weight' -> 0.
    """
    if method in ("Monte Carlo", "MPT"):
        return int(np.argmax(metrics))
    if method in ("VaR", "CVaR"):
        return int(np.argmin(metrics))
    if method == "Equal Weight":
        if len(metrics) == 0:
            raise IndexError("equal weights violate the bounds (app.py:687)")
        return 0
    raise KeyError(method)


def selection_record(idx, weights, ret, risk, sharpe):
    return {"index": int(idx), "weights": np.asarray(weights[idx], dtype=np.float64).copy(),
            "ret": float(ret[idx]), "risk": float(risk[idx]), "sharpe": float(sharpe[idx])}


def evaluate(weights, mu, sigma, risk_free=0.0, risk_target=0.30):
    """Supplied-weights oracle: metrics + both picks, the shape `simulate_portfolios` returns."""
    ret, risk, sharpe = portfolio_metrics(weights, mu, sigma, risk_free)
    W = np.atleast_2d(np.asarray(weights, dtype=np.float64))
    i1 = select_max_sharpe(sharpe)
    i2 = select_target_risk(risk, risk_target)
    return {"risks": risk, "returns": ret, "weights": W, "sharpes": sharpe,
            "max_sharpe": selection_record(i1, W, ret, risk, sharpe),
            "target_risk": selection_record(i2, W, ret, risk, sharpe)}


# --------------------------------------------------------------------------
# "host numpy" vectorised sweep -- the CPU baseline of BASELINE.md section 4.2
# --------------------------------------------------------------------------

def vectorised_sweep(mu, sigma, n_portfolios, risk_free=0.03, risk_target=0.30,
                     seed=0, chunk=1_000_000):
    """Vectorised restatement of app.py:702,708-711 + picks, in chunks.

    ``np.random.RandomState(seed).dirichlet(ones(N), size=chunk)`` (legacy generator,
    as app.py:702), ``W@mu``, ``sqrt(((W@Sigma)*W).sum(1))``, Sharpe, ``argmax`` and
    ``argmin|risk-target|`` with first-occurrence merging across chunks.
    Returns (best_sharpe, idx), (best_dist, idx).
    """
    rng = np.random.RandomState(seed)
    mu = np.asarray(mu, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    ones = np.ones(mu.shape[0])
    best_s, best_si = -np.inf, -1
    best_d, best_di = np.inf, -1
    done = 0
    while done < n_portfolios:
        m = int(min(chunk, n_portfolios - done))
        W = rng.dirichlet(ones, size=m)
        ret = W @ mu
        risk = np.sqrt(((W @ sigma) * W).sum(axis=1))
        sharpe = np.where(risk > 0, (ret - risk_free) / risk, 0.0)
        i = int(np.argmax(sharpe))
        if sharpe[i] > best_s:
            best_s, best_si = float(sharpe[i]), done + i
        d = np.abs(risk - risk_target)
        j = int(np.argmin(d))
        if d[j] < best_d:
            best_d, best_di = float(d[j]), done + j
        done += m
    return (best_s, best_si), (best_d, best_di)


# --------------------------------------------------------------------------
# f3  CSV ingest with the thousands-separator fix          app.py:89-134, 465-482
# --------------------------------------------------------------------------

def load_price_csv(path):
    """Comma-aware version of ``read_csv_file`` (app.py:89-134).

    Documented deviation: ``thousands=','`` -- the reference's ``pd.to_numeric``
    (app.py:127) turns "86,493.0" into NaN and rejects the BTC/ETH/XAU files.
    Price column choice follows app.py:115 (price/close/adj close/open, first hit).
    """
    import pandas as pd
    df = pd.read_csv(path, thousands=",", encoding="utf-8-sig")
    cols = {str(c).strip().lower(): c for c in df.columns}
    date_col = cols["date"]
    price_col = next(cols[k] for k in ("price", "close", "adj close", "open") if k in cols)
    out = df[[date_col, price_col]].dropna().rename(columns={date_col: "Date", price_col: "Price"})
    out["Date"] = pd.to_datetime(out["Date"], errors="coerce")
    out["Price"] = pd.to_numeric(out["Price"], errors="coerce")
    return out.dropna(subset=["Date", "Price"])


_RULE_ALIASES = {"M": "ME", "Q": "QE"}   # pandas >= 2.2 rejects the reference's 'M'/'Q' (app.py:426)


def build_returns(paths, names=None, rule="W", annual_factor=52):
    """app.py:465-482 + 658-667: inner join on Date, resample(rule).last().dropna(), pct_change().fillna(0).

    ``rule=None`` skips the resample (daily data, the C2 policy of SURVEY 8(d)).
    Returns (returns ndarray T x N, names, mu, sigma).
    """
    import pandas as pd
    names = names or [str(p) for p in paths]
    frames = []
    for p, n in zip(paths, names):
        df = load_price_csv(p).rename(columns={"Price": n}).set_index("Date")
        frames.append(df[[n]])
    prices = pd.concat(frames, axis=1, join="inner").sort_index()
    if rule is not None:
        prices = prices.resample(_RULE_ALIASES.get(rule, rule)).last().dropna()
    returns = prices.pct_change().fillna(0).dropna()
    mu = (returns.mean() * annual_factor).to_numpy()
    sigma = (returns.cov() * annual_factor).to_numpy()
    return returns.to_numpy(), names, mu, sigma
