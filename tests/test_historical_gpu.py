"""GPU tier: per-portfolio historical VaR / CVaR (app.py:710-713) and the 'VaR' / 'CVaR' /
'Equal Weight' method picks (app.py:673-676) against golden vectors from the reference's lines."""
import numpy as np
import pytest

from conftest import synthetic_inputs
from oracle import reference_np as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 1e-4)])
def test_c1_historical_var_cvar_golden(mcp, c1, dtype, tol):
    """C1: T=365 weekly rows, the 10k reference draws; -var / -cvar arrays and both picks."""
    out = mcp.historical_var_cvar(c1["returns_matrix"], c1["weights"], 0.95, dtype=dtype)
    assert np.allclose(-out["var"], c1["neg_var95"], rtol=tol, atol=tol * 0.1)
    assert np.allclose(-out["cvar"], c1["neg_cvar95"], rtol=tol, atol=tol * 0.1)
    assert out["best_var"]["index"] == int(c1["opt_var"]) == 4593
    assert out["best_cvar"]["index"] == int(c1["opt_cvar"]) == 4593
    assert np.isclose(out["best_var"]["value"], -0.13301122778135846, rtol=tol)
    assert np.isclose(out["best_cvar"]["value"], -0.1954488675070788, rtol=tol)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 1e-4)])
def test_c2_historical_golden(mcp, c2, dtype, tol):
    out = mcp.historical_var_cvar(c2["returns_matrix"], c2["weights"], 0.95, dtype=dtype)
    assert np.allclose(-out["var"], c2["neg_var95"], rtol=tol, atol=tol * 0.01)
    assert np.allclose(-out["cvar"], c2["neg_cvar95"], rtol=tol, atol=tol * 0.01)
    if dtype == "float64":
        assert out["best_var"]["index"] == int(c2["opt_var"])
        assert out["best_cvar"]["index"] == int(c2["opt_cvar"])


@pytest.mark.parametrize("T", [1, 2, 20, 21, 33, 64, 100, 365, 1000, 2048])
@pytest.mark.parametrize("alpha", [0.95, 0.99, 0.5])
def test_historical_shapes_and_alphas(mcp, T, alpha):
    n = 7
    rng = np.random.default_rng(T)
    R = rng.standard_normal((T, n)) * 0.04
    R[::5] = np.round(R[::5], 2)                       # ties in the series
    W = rng.dirichlet(np.ones(n), size=257)
    out = mcp.historical_var_cvar(R, W, alpha, dtype="float64")
    v, c = ref.historical_var_cvar(R, W, alpha)
    assert np.allclose(out["var"], v, rtol=1e-10, atol=1e-15)
    assert np.allclose(out["cvar"], c, rtol=1e-10, atol=1e-15)
    assert out["best_var"]["index"] == ref.select_method(-v, "VaR")
    assert out["best_cvar"]["index"] == ref.select_method(-c, "CVaR")


def test_historical_device_tensor_and_errors(mcp):
    import torch
    rng = np.random.default_rng(0)
    R = rng.standard_normal((120, 16)) * 0.03
    W = rng.dirichlet(np.ones(16), size=5000)
    out = mcp.historical_var_cvar(R, torch.from_numpy(W).cuda(), 0.95, dtype="float64")
    v, c = ref.historical_var_cvar(R, W, 0.95)
    assert np.allclose(out["var"].cpu().numpy(), v, rtol=1e-10) and np.allclose(out["cvar"].cpu().numpy(), c, rtol=1e-10)
    with pytest.raises(ValueError):
        mcp.historical_var_cvar(R, W[:, :3])
    with pytest.raises(mcp.McpError, match="2048"):
        mcp.historical_var_cvar(rng.standard_normal((3000, 2)), np.ones((4, 2)) / 2)


def test_methods_like_the_app(mcp, c1):
    """The five methods of app.py:671-677 on C1's weekly returns."""
    R = c1["returns_matrix"]
    for method in ("Monte Carlo", "MPT", "VaR", "CVaR"):
        out = mcp.simulate_method(R, method, 2500, annual_factor=52, risk_free=3.0, seed=1, dtype="float64")
        W = out["weights"]
        assert W.shape == (2500, 2)
        ret, risk, sharpe = ref.portfolio_metrics(W, c1["mu"], c1["sigma"], 3.0)
        assert np.allclose(out["risks"], risk, rtol=1e-9)
        if method in ("Monte Carlo", "MPT"):
            want = sharpe
        else:
            v, c = ref.historical_var_cvar(R, W, 0.95)
            want = -v if method == "VaR" else -c
        assert np.allclose(out["metrics"], want, rtol=1e-9, atol=1e-14)
        assert out["opt_idx"] == ref.select_method(want, method)
    ew = mcp.simulate_method(R, "Equal Weight", annual_factor=52, risk_free=3.0, dtype="float64")
    assert np.allclose([ew["risks"][0], ew["returns"][0], ew["metrics"][0]], c1["ew_rf3"], rtol=1e-9)
    with pytest.raises(IndexError):
        mcp.simulate_method(R, "Equal Weight", annual_factor=52, min_weights=np.array([0.6, 0.0]))


def test_asset_stats_kernel(mcp, c1, c2):
    """f2: per-asset statistics (app.py:231-263, 286-335) against the reference functions' outputs."""
    import json, os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "asset_stats.json")) as fh:
        g = json.load(fh)
    for tag, R in (("c1", c1["returns_matrix"]), ("c2", c2["returns_matrix"])):
        spec = g[tag]
        got = mcp.asset_stats(R[1:], annual_factor=spec["ann_factor"], risk_free=spec["risk_free"])
        assert len(got) == R.shape[1]
        for j, want in enumerate(spec["assets"]):
            for k, v in want.items():
                assert got[j][k] == pytest.approx(v, rel=1e-10, abs=1e-14), (tag, j, k)
    # odd lengths, a single period, all-positive series (no downside -> 0.0001, app.py:243)
    rng = np.random.default_rng(3)
    for T in (1, 2, 31, 33, 1000):
        R = np.abs(rng.standard_normal((T, 3))) * 0.01
        R[:, 1] *= -1
        got = mcp.asset_stats(R, annual_factor=12, risk_free=0.0)
        for j in range(3):
            want = ref.asset_stats(R[:, j], 0.0, 12)
            for k in ("total_return_ann", "max_drawdown", "mean_ann", "min_month", "max_month", "var_95", "cvar_95"):
                assert got[j][k] == pytest.approx(want[k], rel=1e-10, abs=1e-14), (T, j, k)
            if T > 2:
                assert got[j]["sharpe"] == pytest.approx(want["sharpe"], rel=1e-9)
                assert got[j]["sortino"] == pytest.approx(want["sortino"], rel=1e-9)


@pytest.mark.parametrize("T", [33, 100, 200, 365, 384, 500])
@pytest.mark.parametrize("alpha", [0.95, 0.99])
def test_fast_fp32_kernel_is_bit_identical_to_the_plain_one(mcp, T, alpha):
    """hist_var_fast (4 portfolios per warp, sorting network, one threshold reduction + a few warp-min pops / warp-max
    removals) against hist_var_kernel<float> (MCP_HIST_FAST=0) on the same inputs, and both against the FP64 oracle; ties
    included.  MCP_HIST_ROW forces the threshold row (0: pops only), so both directions of the finish run at every rank;
    MCP_HIST_REFINE=0 skips the interpolated threshold, MCP_HIST_OCC picks the other compiled block shapes."""
    import os
    rng = np.random.default_rng(T)
    n, P = 16, 1003                                        # P % 4 != 0: the last group is partial
    R = np.round(rng.standard_normal((T, n)) * 0.04, 3)    # rounded returns: repeated series values (ties at the rank)
    W = rng.dirichlet(np.ones(n), size=P)
    W[5] = 0; W[5, 2] = 1.0                                # a one-asset portfolio: the series IS a (tied) column of R
    W[6] = 0                                               # an all-zero row: every series value ties
    R[0::32, 3] = -0.3 - 0.01 * np.arange(len(R[0::32]))   # asset 3's worst periods all belong to lane 0 ...
    W[7] = 0; W[7, 3] = 1.0                                # ... so this portfolio's whole tail sits in one lane
    W[8] = 0; W[8, 3] = 0.7; W[8, 2] = 0.3
    fast = mcp.historical_var_cvar(R, W, alpha, dtype="float32")
    os.environ["MCP_HIST_FAST"] = "0"
    try:
        plain = mcp.historical_var_cvar(R, W, alpha, dtype="float32")
    finally:
        del os.environ["MCP_HIST_FAST"]
    assert np.array_equal(fast["var"], plain["var"]) and np.array_equal(fast["cvar"], plain["cvar"])
    assert fast["best_var"] == plain["best_var"] and fast["best_cvar"] == plain["best_cvar"]
    for knob, value in (("MCP_HIST_ROW", "0"), ("MCP_HIST_ROW", "1"), ("MCP_HIST_ROW", "2"), ("MCP_HIST_REFINE", "0"),
                        ("MCP_HIST_OCC", "2"), ("MCP_HIST_OCC", "3")):
        os.environ[knob] = value
        try:
            forced = mcp.historical_var_cvar(R, W, alpha, dtype="float32")
        finally:
            del os.environ[knob]
        assert np.array_equal(forced["var"], plain["var"]) and np.array_equal(forced["cvar"], plain["cvar"]), (knob, value)
        assert forced["best_var"] == plain["best_var"] and forced["best_cvar"] == plain["best_cvar"], (knob, value)
    v, c = ref.historical_var_cvar(R, W, alpha)
    assert np.allclose(fast["var"], v, rtol=1e-4, atol=1e-7) and np.allclose(fast["cvar"], c, rtol=1e-4, atol=1e-7)
    assert fast["best_var"]["index"] == int(np.argmax(fast["var"])) and fast["best_cvar"]["index"] == int(np.argmax(fast["cvar"]))


def test_fast_fp32_kernel_continuous_returns_many_groups(mcp):
    """Unrounded returns (no ties), enough portfolios for every warp to loop: the fast and the plain kernel agree bit for bit
    and sit within FP32 rounding of the FP64 oracle."""
    import os
    rng = np.random.default_rng(11)
    T, n, P = 365, 16, 300_001
    R = rng.standard_t(4, size=(T, n)) * 0.03              # heavy tails, as weekly crypto returns have
    W = rng.dirichlet(np.ones(n), size=P)
    fast = mcp.historical_var_cvar(R, W, 0.95, dtype="float32")
    os.environ["MCP_HIST_FAST"] = "0"
    try:
        plain = mcp.historical_var_cvar(R, W, 0.95, dtype="float32")
    finally:
        del os.environ["MCP_HIST_FAST"]
    assert np.array_equal(fast["var"], plain["var"]) and np.array_equal(fast["cvar"], plain["cvar"])
    assert fast["best_var"] == plain["best_var"] and fast["best_cvar"] == plain["best_cvar"]
    v, c = ref.historical_var_cvar(R[:, :], W[:20000], 0.95)
    assert np.allclose(fast["var"][:20000], v, rtol=1e-4, atol=1e-7) and np.allclose(fast["cvar"][:20000], c, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("n", [1, 3, 7, 16, 17, 24, 40])
def test_fast_fp32_kernel_shapes_fuzz(mcp, n):
    """Seeded sweep over period counts, universe sizes (n > 16: the weights of a group no longer fit a lane's two prefetch
    registers), ranks and portfolio counts: hist_var_fast == hist_var_kernel<float> bit for bit, both within FP32 rounding of
    the FP64 oracle (app.py:258-263 on app.py:710's series)."""
    import os
    rng = np.random.default_rng(1000 + n)
    for case in range(6):
        T = int(rng.integers(97, 513))
        P = int(rng.integers(1, 700))
        alpha = float(rng.choice([0.9, 0.95, 0.975, 0.99, 0.999]))
        hidx = (T - 1) * ((1 - alpha) * 100 / 100.0)
        if hidx - np.floor(hidx) > 1 - 1e-6:               # gamma = 1 - 4e-15 is 1.0f in FP32: VaR's upper neighbour enters the tail mean
            T += 1                                         # there, numpy's own FP64 lerp sits on the same knife edge
        R = rng.standard_t(3, size=(T, n)) * 0.03
        if case % 2:
            R = np.round(R, 2)                             # ties
        W = rng.dirichlet(np.ones(n), size=P)
        if case == 5:
            W = -W                                         # short book: the tail is the other side of every series
        fast = mcp.historical_var_cvar(R, W, alpha, dtype="float32")
        os.environ["MCP_HIST_FAST"] = "0"
        try:
            plain = mcp.historical_var_cvar(R, W, alpha, dtype="float32")
        finally:
            del os.environ["MCP_HIST_FAST"]
        tag = (n, T, P, alpha, case)
        assert np.array_equal(fast["var"], plain["var"]) and np.array_equal(fast["cvar"], plain["cvar"]), tag
        assert fast["best_var"] == plain["best_var"] and fast["best_cvar"] == plain["best_cvar"], tag
        v, c = ref.historical_var_cvar(R, W, alpha)
        assert np.allclose(fast["var"], v, rtol=1e-4, atol=2e-7) and np.allclose(fast["cvar"], c, rtol=1e-4, atol=2e-7), tag
