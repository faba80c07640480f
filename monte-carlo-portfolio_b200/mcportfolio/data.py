"""Input side of the path: price CSVs -> aligned returns matrix -> (mu, Sigma).

Host-side mirror of the reference's ingest: `read_csv_file`
(app.py:89-134), the price frame (app.py:465-482) and the returns / mu / Sigma step
(app.py:658-667, 679-680).  It runs once per session on a few KB of text, so it stays on the
host (pandas); the device path starts at (mu, Sigma).

Two documented fixes relative to the reference (SURVEY.md "three facts", item 3):
  * thousands separators: `pd.to_numeric("86,493.0")` is NaN in the reference (app.py:127), so
    its loader rejects or truncates every file with prices >= 1000; here commas are stripped.
  * pandas >= 2.2 rejects the resample aliases 'M' / 'Q' the app uses (app.py:426): they are
    mapped to 'ME' / 'QE'.
"""
from __future__ import annotations

import io

import numpy as np

ANNUAL_FACTOR = {"M": 12, "Q": 4, "W": 52, "D": 252}          # app.py:427 (+ daily)
_RULE = {"M": "ME", "Q": "QE"}


def read_price_csv(source):
    """`read_csv_file` (app.py:89-134): returns a DataFrame with columns Date, Price.

    Header sniffing (a row holding a 'date' cell within the first 5 rows, app.py:100-104),
    price column = first of price / close / adj close / open (app.py:115), rows with an
    unparsable date or price dropped (app.py:126-128).  Raises ValueError where the app shows
    an error box."""
    import csv
    import pandas as pd
    if isinstance(source, (bytes, bytearray)):
        text = bytes(source).decode("utf-8-sig")
    elif hasattr(source, "read"):
        text = source.read()
        text = text.decode("utf-8-sig") if isinstance(text, bytes) else text.lstrip("\ufeff")
    else:
        with open(source, encoding="utf-8-sig", newline="") as fh:
            text = fh.read()
    rows = [r for r in csv.reader(io.StringIO(text)) if any(c.strip() for c in r)]
    header_idx = None
    for i, r in enumerate(rows[:6]):                  # the header line itself + 5 data rows (app.py:100)
        if any(c.strip().lower() == "date" for c in r):
            header_idx = i
            break
    if header_idx is None:
        raise ValueError("no header row with a 'date' column found (app.py:105-106)")
    header = [c.strip() for c in rows[header_idx]]
    width = len(header)
    body = [r[:width] + [None] * (width - len(r)) for r in rows[header_idx + 1:]]
    df = pd.DataFrame(body, columns=header, dtype=object)
    cols = [c.lower() for c in header]
    by_name = dict(zip(cols, df.columns))
    date_col = by_name["date"]
    price_col = next((by_name[k] for k in ("price", "close", "adj close", "open") if k in by_name), None)
    if price_col is None:
        others = [c for c in df.columns if c != date_col]
        if not others:
            raise ValueError("no price column found (app.py:118-119)")
        price_col = others[0]
    out = df[[date_col, price_col]].dropna().rename(columns={date_col: "Date", price_col: "Price"})
    out["Date"] = pd.to_datetime(out["Date"], errors="coerce")
    out["Price"] = pd.to_numeric(out["Price"].astype(str).str.replace(",", "", regex=False).str.strip(), errors="coerce")
    out = out.dropna(subset=["Date", "Price"])
    if out.empty:
        raise ValueError("no valid rows after type conversion (app.py:129-130)")
    return out.reset_index(drop=True)


def price_frame(frames, names=None, rule="M"):
    """app.py:465-482: inner join on Date, `resample(rule).last().dropna()`.  rule=None keeps
    the native frequency.  Duplicate names get ' (2)', ' (3)' ... suffixes (app.py:469-472)."""
    import pandas as pd
    names = list(names) if names is not None else [f"asset{i}" for i in range(len(frames))]
    seen, cols = {}, []
    for (df, base) in zip(frames, names):
        seen[base] = seen.get(base, 0) + 1
        name = base if seen[base] == 1 else f"{base} ({seen[base]})"
        cols.append(df.rename(columns={"Price": name}).dropna(subset=[name]).set_index("Date")[[name]])
    prices = pd.concat(cols, axis=1, join="inner").sort_index()
    if rule is not None:
        prices = prices.resample(_RULE.get(rule, rule)).last().dropna()
    return prices


def returns_matrix(prices):
    """app.py:666: `pct_change().fillna(0)` per asset (the leading all-zero row is kept)."""
    return prices.pct_change().fillna(0).dropna()


def mu_sigma(returns, annual_factor):
    """app.py:679-680: `mean() * A`, `cov() * A` (ddof = 1)."""
    r = np.asarray(returns, dtype=np.float64)
    return r.mean(axis=0) * annual_factor, np.atleast_2d(np.cov(r, rowvar=False, ddof=1)) * annual_factor


def load(paths, names=None, rule="M"):
    """CSV files -> (returns ndarray (T, N), names, mu, Sigma) with the app's annual factor."""
    frames = [read_price_csv(p) for p in paths]
    prices = price_frame(frames, names, rule)
    rets = returns_matrix(prices)
    mu, sigma = mu_sigma(rets.to_numpy(), ANNUAL_FACTOR["D" if rule is None else rule])
    return rets.to_numpy(), list(prices.columns), mu, sigma
