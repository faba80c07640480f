"""CPU tier, row f4 (dev container only: needs /root/reference): the reference's OWN lines around the adapter's result.

app.py:671-677 (the `simulation_methods` table) and app.py:719-783 (array materialisation, scatter, capital allocation line,
`opt_idx = config['opt_crit'](...)`, allocation columns, pie) are exec'ed verbatim with recording stubs for `st` / `go` / `px`,
fed from `MethodResult.arrays()` -- the substitution INTEGRATION.md section 2 shows.  The arrays here are the golden vectors the
reference loop itself produced (no GPU in this tier); the GPU tier (test_app_adapter_gpu.py) checks that the adapter returns
those arrays.  What this proves: every name the app's plotting code reads is one the adapter provides, and the app's own
`opt_crit` lambdas pick the adapter's `opt_idx` from `all_metrics`."""
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.needs_reference
APP = "/root/reference/app.py"


class _Rec:
    """Recording stand-in for streamlit / plotly modules and the objects they return."""

    def __init__(self, log, name="st"):
        self._log, self._name = log, name
        self.session_state = {"investment_amount": 1000.0}

    def __getattr__(self, attr):
        def call(*a, **k):
            self._log.append((f"{self._name}.{attr}", a, k))
            if attr == "columns":
                return [_Rec(self._log, "col") for _ in range(a[0])]
            return _Rec(self._log, f"{self._name}.{attr}()")
        return call

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _slice(lines, a, b):
    return textwrap.dedent("".join(lines[a - 1:b]))


def test_app_lines_run_on_the_adapter_result(c1):
    from mcportfolio.app_adapter import MethodResult, METHOD_METRIC
    with open(APP, encoding="utf-8") as fh:
        lines = fh.readlines()
    log = []
    ns = {"np": np, "st": _Rec(log), "go": _Rec(log, "go"), "px": _Rec(log, "px")}
    exec(compile(_slice(lines, 72, 82), "app.py:72-82", "exec"), ns)                 # format_money / format_percent
    exec(compile(_slice(lines, 671, 677), "app.py:671-677", "exec"), ns)
    methods = ns["simulation_methods"]
    assert {m: c["metric"] for m, c in methods.items()} == METHOD_METRIC             # the adapter's table is the app's
    W = c1["weights"]
    golden = {"Monte Carlo": (c1["sharpes_rf3"], int(c1["opt_sharpe_rf3"])), "MPT": (c1["sharpes_rf3"], int(c1["opt_sharpe_rf3"])),
              "VaR": (c1["neg_var95"], int(c1["opt_var"])), "CVaR": (c1["neg_cvar95"], int(c1["opt_cvar"]))}
    tail = compile(_slice(lines, 719, 783), "app.py:719-783", "exec")
    for method, (metrics, opt) in golden.items():
        res = MethodResult(method, METHOD_METRIC[method], c1["risks"], c1["returns"], W, metrics, opt, W[opt])
        ns.update(method=method, config=methods[method], user_rf=3.0, asset_names=["BTC", "ETH"])
        ns["all_risks"], ns["all_returns"], ns["all_weights"], ns["all_metrics"] = res.arrays()
        del log[:]
        exec(tail, ns)
        assert ns["opt_idx"] == res.opt_idx                                          # app.py:747 on the adapter's all_metrics
        assert np.array_equal(ns["weights"], res.opt_weights)                        # app.py:765
        calls = [c[0] for c in log]
        assert calls.count("st.plotly_chart") == 2 and "px.pie" in calls and calls.count("go.Scatter") == (3 if method == "MPT" else 2)
        if method == "MPT":                                                          # app.py:738-744
            cal_x, cal_y = res.capital_allocation_line(3.0)
            assert np.array_equal(ns["cal_x"], cal_x) and np.array_equal(ns["cal_y"], cal_y)
        pie = [c for c in log if c[0] == "px.pie"][0][2]
        assert np.allclose(pie["values"], res.opt_weights * 1000.0) and pie["names"] == ["BTC", "ETH"]
