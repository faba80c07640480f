"""CPU tier: bench.py's expected-results gate (tests/golden/c3c4c5_expected.json) and its compact summary contract."""
import argparse
import copy
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(want):
    return {"selected": dict(want["c3"]), "paths": {"stats": {a: list(v) for a, v in want["c4"]["stats"].items()}},
            "envelope": {"target_risk": {"index": want["c5"]["target_risk_index"]}, "filled_bins": want["c5"]["filled_bins"]}}


def test_expected_file_and_gate():
    with open(os.path.join(ROOT, "tests", "golden", "c3c4c5_expected.json")) as fh:
        want = json.load(fh)
    assert set(want["c4"]["stats"]) == {"0.95", "0.99"}
    for v, c in want["c4"]["stats"].values():
        assert c < v                                   # CVaR = mean of the tail below the VaR quantile (app.py:258-263)
    args = argparse.Namespace(portfolios=bench.P_TOTAL, paths=bench.M_PATHS, envelope_portfolios=bench.P_LARGE)
    good = _line(want)
    assert bench.check_expected(good, args, 1) == []
    # a kernel variant that rounds the contraction differently moves VaR by ~1e-7: inside the tolerance
    near = copy.deepcopy(good)
    near["paths"]["stats"]["0.95"][0] += 3e-7
    assert bench.check_expected(near, args, 8) == []
    for mutate in (lambda l: l["selected"].__setitem__("max_sharpe_index", 1),
                   lambda l: l["selected"].__setitem__("target_risk_index", 2),
                   lambda l: l["paths"]["stats"]["0.99"].__setitem__(1, 0.3),
                   lambda l: l["envelope"].__setitem__("filled_bins", 433),
                   lambda l: l["envelope"]["target_risk"].__setitem__("index", 5)):
        bad = copy.deepcopy(good)
        mutate(bad)
        assert len(bench.check_expected(bad, args, 1)) == 1
    # reduced sizes are not the golden job: nothing to compare
    small = argparse.Namespace(portfolios=10 ** 6, paths=1000, envelope_portfolios=0)
    bad = copy.deepcopy(good)
    bad["selected"]["max_sharpe_index"] = 1
    bad["envelope"] = None
    assert bench.check_expected(bad, small, 1) == []


def test_ncu_figures_come_from_the_committed_file():
    f = bench.ncu_figures("small_sweep_packed<16, 4, 0, 10>") or bench.ncu_figures("small_sweep_packed<16, 4, 0>")
    assert f is not None and 0 < f["fp32_pipe_busy_pct"] < 100 and f["source"].startswith("profiles/")
    assert isinstance(f["digest_matches_build"], bool)
    assert bench.ncu_figures("no such kernel") is None


def test_expected_c3_picks_are_what_the_generator_restatement_gives():
    """The two golden indices of the 1e10-portfolio job, re-derived on the CPU: the numpy restatement of the generator draws the
    weights of those global indices, the FP64 oracle evaluates them -- the Sharpe ratio / risk the GPU reported (FP32, 1e-4) -- and
    inside a 4e5-index neighbourhood of each no other portfolio beats them (the whole range cannot be swept on a CPU)."""
    import numpy as np
    from conftest import synthetic_inputs
    from oracle import philox_np, reference_np as ref
    with open(os.path.join(ROOT, "tests", "golden", "c3c4c5_expected.json")) as fh:
        want = json.load(fh)["c3"]
    mu, sigma = synthetic_inputs(bench.N_ASSETS)
    assert np.array_equal(mu, bench.synthetic_inputs(bench.N_ASSETS)[0])
    half = 200_000
    i = want["max_sharpe_index"]
    W, _ = philox_np.dirichlet_weights(i - half, 2 * half, bench.N_ASSETS, seed=bench.SEED)
    e = ref.evaluate(W, mu, sigma, bench.RISK_FREE, bench.RISK_TARGET)
    assert int(np.argmax(e["sharpes"])) == half and np.isclose(e["sharpes"][half], want["max_sharpe"], rtol=1e-4)
    j = want["target_risk_index"]
    W, _ = philox_np.dirichlet_weights(j - half, 2 * half, bench.N_ASSETS, seed=bench.SEED)
    e = ref.evaluate(W, mu, sigma, bench.RISK_FREE, bench.RISK_TARGET)
    assert int(np.argmin(np.abs(e["risks"] - bench.RISK_TARGET))) == half and np.isclose(e["risks"][half], want["target_risk"], rtol=1e-4)
    # C5 (256 assets): the nearest-to-30 % pick is the riskiest portfolio of the job; same check on a 4e4-index neighbourhood
    with open(os.path.join(ROOT, "tests", "golden", "c3c4c5_expected.json")) as fh:
        c5 = json.load(fh)["c5"]
    mu, sigma = synthetic_inputs(bench.N_LARGE)
    k, half = c5["target_risk_index"], 20_000
    W, _ = philox_np.dirichlet_weights(k - half, 2 * half, bench.N_LARGE, seed=bench.SEED)
    e = ref.evaluate(W, mu, sigma, bench.RISK_FREE, bench.RISK_TARGET)
    assert int(np.argmin(np.abs(e["risks"] - bench.RISK_TARGET))) == half and np.isclose(e["risks"][half], c5["target_risk"], rtol=1e-4)
