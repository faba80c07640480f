"""CPU tier: the ingest mirror (f3) -- investing.com CSV format, thousands separators, aliases."""
import io
import os

import numpy as np
import pytest

from conftest import load_golden

CSV_A = ('﻿"Date","Price","Open","High","Low","Vol.","Change %"\n'
         '"03/16/2025","86,493.0","86,397.0","95,283.0","81,740.0","10.27K","0.13%"\n'
         '"03/09/2025","86,382.5","96,721.0","96,767.0","78,651.0","18.69K","-10.69%"\n'
         '"03/02/2025","96,700.0","94,000.0","99,000.0","93,000.0","9.1K","2.9%"\n'
         '"02/23/2025","950.25","94,000.0","99,000.0","93,000.0","9.1K","2.9%"\n')
CSV_B = ('exported by somebody\nsecond junk line\nDate,Close\n2025-02-23,10.0\n2025-03-02,11.0\n2025-03-09,bad\n'
         '2025-03-16,12.1\nnot a date,5\n')


def test_read_price_csv_handles_bom_quotes_and_thousands():
    from mcportfolio import data
    df = data.read_price_csv(io.StringIO(CSV_A))
    assert list(df.columns) == ["Date", "Price"] and len(df) == 4
    assert df["Price"].tolist() == [86493.0, 86382.5, 96700.0, 950.25]          # the reference would keep only 950.25
    assert str(df["Date"].iloc[0].date()) == "2025-03-16"


def test_header_sniffing_and_bad_rows():
    from mcportfolio import data
    df = data.read_price_csv(io.StringIO(CSV_B))
    assert df["Price"].tolist() == [10.0, 11.0, 12.1] and len(df) == 3
    with pytest.raises(ValueError, match="date"):
        data.read_price_csv(io.StringIO("a,b\n1,2\n3,4\n"))
    with pytest.raises(ValueError, match="no valid rows"):
        data.read_price_csv(io.StringIO("Date,Price\nxx,yy\n"))


def test_price_frame_join_resample_and_returns():
    from mcportfolio import data
    a = data.read_price_csv(io.StringIO(CSV_A))
    b = data.read_price_csv(io.StringIO(CSV_B))
    prices = data.price_frame([a, b], ["A", "A"], rule="W")                     # duplicate names get a suffix
    assert list(prices.columns) == ["A", "A (2)"]
    assert len(prices) == 3                                                     # 03/09 has no valid B price -> inner join drops it
    r = data.returns_matrix(prices).to_numpy()
    assert np.all(r[0] == 0.0) and np.isclose(r[1, 1], 0.1)
    mu, sigma = data.mu_sigma(r, 52)
    assert mu.shape == (2,) and sigma.shape == (2, 2) and np.allclose(sigma, sigma.T)
    for rule in ("M", "Q"):                                                     # the app's aliases work on pandas >= 2.2
        assert len(data.price_frame([a, b], ["A", "B"], rule=rule)) >= 1


@pytest.mark.needs_reference
def test_c1_files_reproduce_the_golden_inputs():
    from mcportfolio import data
    d = "/root/reference/data"
    R, names, mu, sigma = data.load([os.path.join(d, "BTC_USD 7 Years Weekly.csv"), os.path.join(d, "ETH_USD 7 Years Weekly.csv")],
                                    ["BTC", "ETH"], rule="W")
    g = load_golden("c1_btc_eth.npz")
    assert R.shape == (365, 2) and np.allclose(R, g["returns_matrix"], rtol=0, atol=0)
    assert np.allclose(mu, g["mu"], rtol=1e-13) and np.allclose(sigma, g["sigma"], rtol=1e-13)


def test_option_overlay_matches_reference_lines():
    """f4: calc_options_series (app.py:182-193) outputs recorded from the reference's own lines."""
    import json
    from conftest import GOLDEN
    from mcportfolio.overlay import KINDS, option_overlay_returns
    with open(os.path.join(GOLDEN, "option_overlay.json"), encoding="utf-8") as fh:
        g = json.load(fh)
    prices = np.array(g["prices"])
    assert (prices == 0).any()                                   # the prev_price != 0 guard is exercised
    for case in g["cases"]:
        legs = [tuple(l) for l in case["legs"]]
        got = option_overlay_returns(legs, prices)
        assert np.allclose(got, case["returns"], rtol=1e-13, atol=1e-15), case["name"]
        english = [(KINDS.get(k, k), s, p, q) for k, s, p, q in legs]
        assert np.array_equal(option_overlay_returns(english, prices), got)
    assert option_overlay_returns([], prices).tolist() == [0.0] * len(prices)
    assert option_overlay_returns([("long_asset", 0, 0, 1)], prices[:1]).tolist() == [0.0]


def test_returns_with_overlays_feeds_the_frame():
    import pandas as pd
    from mcportfolio.overlay import returns_with_overlays
    idx = pd.date_range("2025-01-05", periods=5, freq="W")
    prices = pd.DataFrame({"A": [10.0, 11.0, 12.1, 11.0, 12.0], "B": [5.0, 5.5, 5.0, 5.0, 6.0]}, index=idx)
    r = returns_with_overlays(prices, {"A": [("long_asset", 0, 0, 1.0)]})
    assert np.allclose(r["A"].to_numpy(), prices["A"].pct_change().fillna(0).to_numpy())     # long asset == pct_change
    assert np.allclose(r["B"].to_numpy(), prices["B"].pct_change().fillna(0).to_numpy())
    r2 = returns_with_overlays(prices, {"B": [("short_asset", 0, 0, 1.0)]})
    assert np.allclose(r2["B"].to_numpy(), -prices["B"].pct_change().fillna(0).to_numpy())


@pytest.mark.needs_reference
def test_option_overlay_against_the_reference_function_randomised():
    """f4, dev container only: calc_options_series (app.py:182-193) exec'd from the reference on random legs / prices
    (including zero prices, unknown tags, negative quantities) against the vectorised host mirror."""
    import pandas as pd
    import textwrap
    from oracle import ref_loader
    from mcportfolio.overlay import KINDS, option_overlay_returns
    with open(ref_loader.REFERENCE_APP, encoding="utf-8") as fh:
        lines = fh.readlines()
    ns = {"np": np, "pd": pd}
    exec(compile(textwrap.dedent("".join(lines[163:193])), ref_loader.REFERENCE_APP, "exec"), ns)
    tags = list(KINDS) + ["something else"]
    rng = np.random.default_rng(42)
    for trial in range(40):
        T = int(rng.integers(1, 60))
        prices = np.round(50 * np.cumprod(1 + 0.05 * rng.standard_normal(T)), 2)
        if T > 3 and trial % 3 == 0:
            prices[rng.integers(0, T)] = 0.0
        legs = [(tags[rng.integers(len(tags))], float(np.round(rng.uniform(20, 90), 1)), float(np.round(rng.uniform(0, 5), 2)),
                 float(np.round(rng.uniform(-2, 2), 2))) for _ in range(int(rng.integers(0, 5)))]
        want = ns["calc_options_series"](legs, pd.Series(prices)).to_numpy()
        got = option_overlay_returns(legs, prices)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-15), (trial, legs)
