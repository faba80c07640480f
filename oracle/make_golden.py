"""Generate ``tests/golden/*`` by running the reference's own lines in THIS container.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Run once here
(``python -m oracle.make_golden``); the outputs are committed because
``/root/reference`` does not travel to the GPU box.

What is executed verbatim from ``/root/reference/app.py`` (read at run time,
never copied into the repo):
  * 164-193  ``calc_option_return`` / ``calc_options_series`` (option overlay, row f4)
  * 231-284  ``var``, ``cvar``, ``efficient_frontier`` ... (via ``ref_loader``)
  * 671-677  the ``simulation_methods`` table (the ``opt_crit`` lambdas)
  * 679-680  mu / Sigma estimation
  * 682-722  the sampling + evaluation loop and array materialisation
  * 747      the ``opt_crit`` call (restated: same dict-of-aliases argument)

Documented deviations needed to make those lines run on the data files
(SURVEY.md section 8(d)): prices are read with ``thousands=','`` (the reference
loader turns "86,493.0" into NaN) and C2 uses the 14 non-weekly files on their
24 common dates with daily returns and annual factor 252.
"""
from __future__ import annotations

import glob
import json
import os
import textwrap

import numpy as np
import pandas as pd

from . import ref_loader
from .reference_np import build_returns

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Stub:
    """Stands in for ``st`` inside the exec'd slices (only ``st.markdown`` is reached)."""

    def __getattr__(self, name):
        return lambda *a, **k: None


def _slice(lines, a, b):
    return textwrap.dedent("".join(lines[a - 1:b]))


def run_reference_methods(returns, names, annual_factor, user_rf, n_portfolios, seed,
                          min_weights, max_weights, methods):
    """Exec app.py:671-677, 679-680, 682-722 per method; returns {method: dict of arrays}."""
    with open(ref_loader.REFERENCE_APP, encoding="utf-8") as fh:
        lines = fh.readlines()
    fns = ref_loader.load_reference_functions()
    ns = dict(fns)
    ns.update(st=_Stub(), returns_df=pd.DataFrame(returns, columns=names), asset_names=list(names),
              annual_factor=annual_factor, user_rf=user_rf,
              min_weights=np.asarray(min_weights, dtype=float),
              max_weights=np.asarray(max_weights, dtype=float))
    exec(compile(_slice(lines, 671, 677), "app.py:671-677", "exec"), ns)
    all_methods = ns["simulation_methods"]
    exec(compile(_slice(lines, 679, 680), "app.py:679-680", "exec"), ns)
    loop_src = _slice(lines, 682, 722)
    out = {}
    for method in methods:
        ns["simulation_methods"] = {method: all_methods[method]}
        ns["n_portfolios"] = n_portfolios
        np.random.seed(seed)
        exec(compile(loop_src, "app.py:682-722", "exec"), ns)
        m = ns["all_metrics"]
        # app.py:747
        opt_idx = all_methods[method]["opt_crit"]({"sharpe": m, "var_95": m, "cvar_95": m})
        out[method] = {"risks": ns["all_risks"].copy(), "returns": ns["all_returns"].copy(),
                       "weights": ns["all_weights"].copy(), "metrics": m.copy(),
                       "opt_idx": int(opt_idx)}
    out["_mu"] = ns["mean_returns"].to_numpy()
    out["_sigma"] = ns["cov_matrix"].to_numpy()
    return out


def make_option_overlay():
    """f4: the reference's calc_options_series (app.py:182-193) on a price walk with a planted 0."""
    with open(ref_loader.REFERENCE_APP, encoding="utf-8") as fh:
        lines = fh.readlines()
    ns = {"np": np, "pd": pd}
    exec(compile(_slice(lines, 164, 193), ref_loader.REFERENCE_APP, "exec"), ns)
    rng = np.random.default_rng(164)
    prices = np.round(100.0 * np.cumprod(1 + 0.04 * rng.standard_normal(40)), 2)
    prices[5] = 0.0                                          # exercises the prev_price != 0 guard
    cases = {
        "protective_put": [["خرید دارایی", 0, 0, 1.0], ["خرید پوت", 95.0, 2.5, 1.0]],
        "covered_call": [["خرید دارایی", 0, 0, 1.0], ["فروش کال", 110.0, 3.0, 1.0]],
        "collar": [["خرید دارایی", 0, 0, 1.0], ["خرید پوت", 90.0, 2.0, 1.0], ["فروش کال", 115.0, 1.5, 1.0]],
        "short_all": [["فروش دارایی", 0, 0, 0.5], ["فروش فیوچرز", 0, 0, 0.25], ["فروش پوت", 100.0, 4.0, 2.0],
                      ["خرید کال", 105.0, 1.0, 0.75]],
        "unknown": [["چیز دیگر", 1, 1, 1]],
    }
    out = [{"name": k, "legs": legs,
            "returns": ns["calc_options_series"]([tuple(l) for l in legs], pd.Series(prices)).to_numpy().tolist()}
           for k, legs in cases.items()]
    with open(os.path.join(GOLDEN, "option_overlay.json"), "w", encoding="utf-8") as fh:
        json.dump({"note": "calc_options_series (app.py:182-193) executed verbatim by oracle/make_golden.py",
                   "prices": prices.tolist(), "cases": out}, fh, ensure_ascii=False)


def main():
    assert ref_loader.available(), "reference not present; golden vectors are generated in the dev container"
    os.makedirs(GOLDEN, exist_ok=True)
    data = ref_loader.REFERENCE_DATA
    fns = ref_loader.load_reference_functions()
    meta = {"numpy": np.__version__, "pandas": pd.__version__, "generator": "oracle/make_golden.py"}

    # ---------------- C1: BTC + ETH weekly, 10k portfolios --------------------------------
    c1_files = [os.path.join(data, "BTC_USD 7 Years Weekly.csv"),
                os.path.join(data, "ETH_USD 7 Years Weekly.csv")]
    R, names, mu, sigma = build_returns(c1_files, names=["BTC", "ETH"], rule="W", annual_factor=52)
    P = 10_000
    c1 = {"returns_matrix": R, "mu": mu, "sigma": sigma}
    for tag, rf in (("rf3", 3.0), ("rf003", 0.03)):
        res = run_reference_methods(R, names, 52, rf, P, 42, [0.0, 0.0], [1.0, 1.0],
                                    ["Monte Carlo", "VaR", "CVaR", "Equal Weight"])
        assert np.allclose(res["_mu"], mu, rtol=1e-13) and np.allclose(res["_sigma"], sigma, rtol=1e-13)
        c1["mu"], c1["sigma"] = res["_mu"], res["_sigma"]    # the reference lines' own mu / Sigma (679-680)
        mc = res["Monte Carlo"]
        c1["weights"] = mc["weights"]                       # same seed -> same draws for every method
        c1["risks"] = mc["risks"]
        c1["returns"] = mc["returns"]
        c1[f"sharpes_{tag}"] = mc["metrics"]
        c1[f"opt_sharpe_{tag}"] = mc["opt_idx"]
        assert np.array_equal(res["VaR"]["weights"], mc["weights"])
        c1["neg_var95"] = res["VaR"]["metrics"]
        c1["neg_cvar95"] = res["CVaR"]["metrics"]
        c1["opt_var"] = res["VaR"]["opt_idx"]
        c1["opt_cvar"] = res["CVaR"]["opt_idx"]
        ew = res["Equal Weight"]
        c1[f"ew_{tag}"] = np.array([ew["risks"][0], ew["returns"][0], ew["metrics"][0]])
    c1["opt_target30"] = int(np.argmin(np.abs(c1["risks"] - 0.30)))   # a9: spec, not reference code
    np.savez_compressed(os.path.join(GOLDEN, "c1_btc_eth.npz"), **c1)

    # bounded run: min/max weights exercise the rejection + skip semantics (app.py:700-707)
    res = run_reference_methods(R, names, 52, 0.03, 2000, 123, [0.50, 0.0], [0.505, 1.0], ["Monte Carlo"])
    mc = res["Monte Carlo"]
    np.savez_compressed(os.path.join(GOLDEN, "c1_bounded.npz"), mu=mu, sigma=sigma,
                        min_weights=np.array([0.50, 0.0]), max_weights=np.array([0.505, 1.0]),
                        weights=mc["weights"], risks=mc["risks"], returns=mc["returns"],
                        sharpes=mc["metrics"], opt_idx=mc["opt_idx"], n_requested=2000, seed=123)
    # a bound no draw can satisfy -> every portfolio skipped, empty arrays (app.py:706-707)
    res = run_reference_methods(R, names, 52, 0.03, 5, 1, [0.6, 0.6], [1.0, 1.0], [])
    # (running 'Monte Carlo' here would raise at argmax of an empty array -- app.py:747; recorded as such)

    # ---------------- efficient_frontier (dead-code twin) verbatim -------------------------
    ef = {}
    np.random.seed(7)
    results, weights = fns["efficient_frontier"](mu, sigma, points=5)
    ef["seed7_results"] = results
    ef["seed7_weights"] = weights
    np.random.seed(11)
    results, weights = fns["efficient_frontier"](pd.Series(mu), pd.DataFrame(sigma), points=200)
    ef["seed11_results"] = results
    ef["seed11_weights"] = weights
    np.random.seed(5)
    results, weights = fns["efficient_frontier"](mu, sigma, points=50,
                                                 min_weights=np.array([0.49, 0.0]),
                                                 max_weights=np.array([0.5, 1.0]))
    ef["seed5_bounded_results"] = results       # exhausts 100 tries sometimes -> keeps last draw (277)
    ef["seed5_bounded_weights"] = weights
    ef["mu"], ef["sigma"] = mu, sigma
    np.savez_compressed(os.path.join(GOLDEN, "efficient_frontier.npz"), **ef)

    # ---------------- var / cvar verbatim on assorted vectors ------------------------------
    rng = np.random.default_rng(2024)
    cases = []
    for n in (1, 2, 3, 20, 21, 100, 365):
        for kind in ("normal", "ties"):
            x = rng.standard_normal(n) * 0.05
            if kind == "ties":
                x = np.round(x, 2)
            for alpha in (0.95, 0.99, 0.5):
                cases.append({"x": x.tolist(), "alpha": alpha,
                              "var": float(fns["var"](x, alpha)), "cvar": float(fns["cvar"](x, alpha))})
    # historical portfolio series of C1 for a few weights (pandas Series operand, as app.py:710-713)
    Rdf = pd.DataFrame(R, columns=names)
    for w in ([0.5, 0.5], [1.0, 0.0], [0.1348708120084909, 0.8651291879915091]):
        s = Rdf @ np.array(w)
        cases.append({"x": s.to_numpy().tolist(), "alpha": 0.95,
                      "var": float(fns["var"](s, 0.95)), "cvar": float(fns["cvar"](s, 0.95))})
    with open(os.path.join(GOLDEN, "var_cvar.json"), "w") as fh:
        json.dump({"meta": meta, "cases": cases}, fh)

    # ---------------- C2 real-data policy: 14 non-weekly files, 24 common dates ------------
    files = sorted(f for f in glob.glob(os.path.join(data, "*.csv")) if "Weekly" not in f)
    names14 = [os.path.basename(f).split(" ")[0].split("_")[0] for f in files]
    R2, names14, mu2, sigma2 = build_returns(files, names=names14, rule=None, annual_factor=252)
    res = run_reference_methods(R2, names14, 252, 0.03, 300, 99, [0.0] * 14, [1.0] * 14,
                                ["Monte Carlo", "VaR", "CVaR"])
    mc = res["Monte Carlo"]
    np.savez_compressed(os.path.join(GOLDEN, "c2_14assets.npz"), returns_matrix=R2, mu=mu2, sigma=sigma2,
                        names=np.array(names14), weights=mc["weights"], risks=mc["risks"],
                        returns=mc["returns"], sharpes=mc["metrics"], opt_sharpe=mc["opt_idx"],
                        neg_var95=res["VaR"]["metrics"], neg_cvar95=res["CVaR"]["metrics"],
                        opt_var=res["VaR"]["opt_idx"], opt_cvar=res["CVaR"]["opt_idx"])

    # ---------------- per-asset stats (f2) verbatim: max_drawdown holds the only cumprod ----
    def ref_stats(col, rf, A):
        ser = pd.Series(col)
        return {"sharpe": float(fns["sharpe_ratio"](ser, rf, A)), "sortino": float(fns["sortino_ratio"](ser, rf, A)),
                "volatility_ann": float(fns["annual_volatility"](ser, A)), "total_return_ann": float(fns["annual_return"](ser, A)),
                "mean_ann": float(np.mean(ser) * A), "mean_month": float(np.mean(ser)), "std_month": float(np.std(ser, ddof=1)),
                "min_month": float(np.min(ser)), "max_month": float(np.max(ser)),
                "max_drawdown": float(fns["max_drawdown"](ser)), "var_95": float(fns["var"](ser, 0.95)),
                "cvar_95": float(fns["cvar"](ser, 0.95))}
    stats = {"c1": {"risk_free": 0.03, "ann_factor": 52, "assets": [ref_stats(R[1:, j], 0.03, 52) for j in range(2)]},
             "c2": {"risk_free": 0.0, "ann_factor": 252, "assets": [ref_stats(R2[1:, j], 0.0, 252) for j in range(R2.shape[1])]}}
    with open(os.path.join(GOLDEN, "asset_stats.json"), "w") as fh:
        json.dump({"meta": meta, "note": "reference functions app.py:231-263 on the returns columns without the leading "
                   "fillna(0) row (calc_asset_stats uses .dropna(), app.py:288)", **stats}, fh)

    make_option_overlay()

    with open(os.path.join(GOLDEN, "META.json"), "w") as fh:
        json.dump({**meta, "c1_shape": list(R.shape), "c2_shape": list(R2.shape),
                   "c1_mu": mu.tolist(), "c1_sigma": sigma.tolist()}, fh, indent=1)
    print("golden vectors written to", GOLDEN)
    for f in sorted(os.listdir(GOLDEN)):
        print(f"  {f:32s} {os.path.getsize(os.path.join(GOLDEN, f)):>9d} B")


if __name__ == "__main__":
    main()
