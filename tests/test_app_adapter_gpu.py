"""GPU tier, row f4: the Monte Carlo tab through mcportfolio.app_adapter against the golden vectors that the reference's own
lines produced (oracle/make_golden.py execs app.py:671-677, 679-680, 682-722 and the pick of 747).  The reference's seeded
numpy draws are passed in as `weights`, so arrays AND picks must be the reference's."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

# app.py:672-676 restated (the committed golden picks come from the app's own lambdas; here they re-derive opt_idx from
# the adapter's arrays the way app.py:747 does: one array under all three names)
OPT_CRIT = {"Monte Carlo": lambda x: np.argmax(x["sharpe"]), "VaR": lambda x: np.argmin(x["var_95"]),
            "CVaR": lambda x: np.argmin(x["cvar_95"]), "MPT": lambda x: np.argmax(x["sharpe"]), "Equal Weight": lambda x: 0}


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 1e-4)])
def test_tab_reproduces_the_reference_run_c1(mcp, c1, dtype, tol):
    names = ["BTC", "ETH"]
    df = pd.DataFrame(c1["returns_matrix"], columns=names)
    tab = mcp.app_adapter.run_monte_carlo_tab(df, names, 52, 3.0, np.zeros(2), np.ones(2), 10_000, weights=c1["weights"], dtype=dtype)
    assert list(tab) == list(mcp.METHODS)                                            # the app's order (app.py:671-677)
    assert np.allclose(tab.mean_returns, c1["mu"], rtol=1e-12) and np.allclose(tab.cov_matrix, c1["sigma"], rtol=1e-12)
    mc = tab["Monte Carlo"]
    assert np.allclose(mc.all_risks, c1["risks"], rtol=tol) and np.allclose(mc.all_returns, c1["returns"], rtol=tol, atol=tol * 1e-3)
    assert np.allclose(mc.all_metrics, c1["sharpes_rf3"], rtol=tol * 10, atol=tol) and np.allclose(mc.all_weights, c1["weights"], atol=1e-7)
    assert mc.opt_idx == int(c1["opt_sharpe_rf3"]) and tab["MPT"].opt_idx == mc.opt_idx
    assert np.allclose(tab["VaR"].all_metrics, c1["neg_var95"], rtol=tol, atol=tol * 0.1) and tab["VaR"].opt_idx == int(c1["opt_var"])
    assert np.allclose(tab["CVaR"].all_metrics, c1["neg_cvar95"], rtol=tol, atol=tol * 0.1) and tab["CVaR"].opt_idx == int(c1["opt_cvar"])
    ew = tab["Equal Weight"]
    assert ew.all_risks.shape == (1,) and ew.all_weights.shape == (1, 2) and ew.opt_idx == 0
    assert np.allclose([ew.all_risks[0], ew.all_returns[0], ew.all_metrics[0]], c1["ew_rf3"], rtol=tol)
    for method, r in tab.items():
        m = r.all_metrics
        # app.py:747: opt_idx = config['opt_crit']({'sharpe': m, 'var_95': m, 'cvar_95': m}) on the returned arrays
        want = int(OPT_CRIT[method]({"sharpe": m, "var_95": m, "cvar_95": m}))
        if dtype == "float64" or method in ("Equal Weight",):
            assert r.opt_idx == want
        assert np.array_equal(r.opt_weights, np.asarray(r.all_weights[r.opt_idx], dtype=np.float64))        # app.py:765
        assert r.all_risks.shape == r.all_returns.shape == r.all_metrics.shape == (len(r.all_weights),)
    cal_x, cal_y = tab["MPT"].capital_allocation_line(3.0)                                                   # app.py:738-744
    assert cal_x.shape == (100,) and np.isclose(cal_x[-1], tab["MPT"].all_risks.max() * 130)
    assert np.allclose(cal_y, 300 + tab["MPT"].all_metrics[tab["MPT"].opt_idx] * cal_x)


def test_tab_c2_fourteen_assets(mcp, c2):
    names = [str(x) for x in c2["names"]]
    df = pd.DataFrame(c2["returns_matrix"], columns=names)
    tab = mcp.app_adapter.run_monte_carlo_tab(df, names, 252, 0.03, [0.0] * 14, [1.0] * 14, 300, weights=c2["weights"], dtype="float64",
                                              methods=("Monte Carlo", "VaR", "CVaR"))
    assert np.allclose(tab["Monte Carlo"].all_metrics, c2["sharpes"], rtol=1e-8) and tab["Monte Carlo"].opt_idx == int(c2["opt_sharpe"])
    assert tab["VaR"].opt_idx == int(c2["opt_var"]) and tab["CVaR"].opt_idx == int(c2["opt_cvar"])
    # columns in another order than asset_names: the adapter follows asset_names (app.py:658-667 builds the frame that way)
    perm = df[df.columns[::-1]]
    tab2 = mcp.app_adapter.run_monte_carlo_tab(perm, names, 252, 0.03, [0.0] * 14, [1.0] * 14, 300, weights=c2["weights"], dtype="float64",
                                               methods=("Monte Carlo",))
    assert np.array_equal(tab2["Monte Carlo"].all_metrics, tab["Monte Carlo"].all_metrics)


def test_tab_in_kernel_draws_bounds_and_empty_equal_weight(mcp, c1):
    names = ["BTC", "ETH"]
    df = pd.DataFrame(c1["returns_matrix"], columns=names)
    lo, hi = np.array([0.50, 0.0]), np.array([0.505, 1.0])                  # the bounded golden run's bounds (c1_bounded.npz)
    tab = mcp.app_adapter.run_monte_carlo_tab(df, names, 52, 0.03, lo, hi, 2000, seed=5)
    mc = tab["Monte Carlo"]
    assert 0 < len(mc.all_risks) < 2000                                     # skipped portfolios are absent (app.py:706-707)
    assert (mc.all_weights[:, 0] >= 0.5).all() and (mc.all_weights[:, 0] <= np.float32(0.505)).all()
    assert mc.opt_idx == int(np.argmax(mc.all_metrics))
    for m in ("VaR", "CVaR"):
        assert np.array_equal(tab[m].all_weights, mc.all_weights)           # same seed: same portfolios for every method
        assert len(tab[m].all_metrics) == len(mc.all_risks)
    ew = tab["Equal Weight"]                                                # 1/N = 0.5 is inside [0.5, 0.505] x [0, 1]
    assert ew.all_weights.shape == (1, 2)
    tab = mcp.app_adapter.run_monte_carlo_tab(df, names, 52, 0.03, np.array([0.6, 0.0]), np.ones(2), 50, methods=("Equal Weight",))
    assert tab["Equal Weight"].all_risks.shape == (0,) and tab["Equal Weight"].all_weights.shape == (0, 2)      # app.py:687


def test_build_returns_df_with_overlay(mcp):
    rng = np.random.default_rng(0)
    idx = pd.date_range("2024-01-31", periods=30, freq="ME")
    prices = {"A": pd.Series(100 * np.cumprod(1 + 0.03 * rng.standard_normal(30)), index=idx),
              "B": pd.Series(50 * np.cumprod(1 + 0.05 * rng.standard_normal(30)), index=idx)}
    legs = {"B": [("خرید دارایی", 0, 0, 1.0), ("خرید پوت", 45.0, 1.5, 1.0)]}
    df = mcp.app_adapter.build_returns_df(prices, ["A", "B"], legs)
    assert list(df.columns) == ["A", "B"] and len(df) == 30 and df.iloc[0].tolist() == [0.0, 0.0]
    assert np.allclose(df["A"].to_numpy()[1:], prices["A"].pct_change().to_numpy()[1:])
    from mcportfolio.overlay import option_overlay_returns
    assert np.array_equal(df["B"].to_numpy(), option_overlay_returns(legs["B"], prices["B"].to_numpy()))
    tab = mcp.app_adapter.run_monte_carlo_tab(df, ["A", "B"], 12, 0.02, np.zeros(2), np.ones(2), 500, seed=1)
    assert len(tab["Monte Carlo"].all_risks) == 500
