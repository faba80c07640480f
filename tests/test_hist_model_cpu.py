"""CPU tier: the selection ALGORITHM of hist_var_fast (threshold set, interpolated second threshold, single steps in
either direction, tie loop, per-lane ascending tail sums) restated lane by lane in oracle/hist_select_model.py, against a
plain sort of the same FP32 series and against the FP64 reference definitions (app.py:258-263).  The kernel itself is
compared with the plain kernel and the oracle in tests/test_historical_gpu.py; this file pins the set logic where no GPU
is needed: every threshold row, arbitrary `shrink` tables (any threshold <= tau must be exact), ties, constant series,
tails that sit in one lane."""
import numpy as np
import pytest

from oracle import hist_select_model as hm
from oracle import reference_np as ref


def _check(series, alpha, row=None, shrink=None):
    x = np.asarray(series, dtype=np.float32)
    var, cvar, total = hm.fast_select(x, alpha, row=row, shrink=shrink)
    v_lo, v_hi, var_s, n_tail, mean_tail = hm.sorted_reference(x, alpha)
    assert var == var_s, (var, var_s)                              # same order statistics, same FP32 lerp: bit-equal
    assert total == n_tail, (total, n_tail)                         # the taken values ARE {x <= VaR}
    assert cvar == pytest.approx(mean_tail, rel=2e-6, abs=1e-9)
    # and the reference's own definitions on the FP32 series (FP64 lerp): within FP32 rounding
    k_lo, k_hi, gamma = hm.order_statistics(len(x), alpha)
    if 1e-6 < gamma < 1 - 1e-6 or k_hi == k_lo:
        x64 = x.astype(np.float64)
        assert float(var) == pytest.approx(ref.var(x64, alpha), rel=1e-6, abs=1e-9)
        assert float(cvar) == pytest.approx(ref.cvar(x64, alpha), rel=1e-5, abs=1e-9)


@pytest.mark.parametrize("T", [97, 128, 200, 365, 384, 500, 512])
@pytest.mark.parametrize("alpha", [0.9, 0.95, 0.99])
def test_default_row_matches_a_plain_sort(T, alpha):
    rng = np.random.default_rng(T * 1000 + int(alpha * 100))
    for case in range(12):
        x = rng.standard_t(3, size=T) * 0.03
        if case % 3 == 1:
            x = np.round(x, 2)                                      # ties, also at the rank
        if case % 3 == 2:
            x = np.round(x, 1)                                      # a handful of distinct values
        _check(x, alpha)


@pytest.mark.parametrize("row", [0, 1, 2])
def test_every_row_and_any_shrink_table_is_exact(row):
    """The threshold row and the interpolation table only decide how many single steps follow, never the result."""
    rng = np.random.default_rng(row)
    for case in range(40):
        T = int(rng.integers(97, 513))
        alpha = float(rng.choice([0.9, 0.95, 0.975, 0.99]))
        x = rng.standard_normal(T) * 0.02
        if case % 2:
            x = np.round(x, 2)
        table = rng.uniform(0, 1, size=68).astype(np.float32) if case % 4 < 2 else None
        if case % 8 == 7:
            table = np.zeros(68, dtype=np.float32)                  # theta = the minimum of the lanes' row-th values
        _check(x, alpha, row=row, shrink=table)


def test_degenerate_series():
    T = 365
    _check(np.zeros(T), 0.95)                                       # every value ties
    _check(np.full(T, -0.25), 0.99)
    x = np.linspace(0.1, 0.2, T)
    x[0::32] = -1.0 - 0.01 * np.arange(len(x[0::32]))               # the whole tail belongs to lane 0
    _check(x, 0.95)
    _check(x, 0.95, row=1)
    _check(-np.abs(np.random.default_rng(5).standard_cauchy(T)), 0.95)      # huge negative outliers
    x = np.random.default_rng(6).standard_normal(T)
    x[7] = x[40] = x[300] = np.sort(x)[18]                          # ties exactly at k_lo
    _check(x, 0.95)


def test_order_statistics_follow_numpy():
    """k_lo / k_hi / gamma as np.percentile's 'linear' method defines them (the C side repeats this arithmetic)."""
    rng = np.random.default_rng(9)
    for _ in range(200):
        T = int(rng.integers(1, 600))
        alpha = float(rng.choice([0.0, 0.5, 0.9, 0.95, 0.99, 1.0]))
        x = rng.standard_normal(T)
        k_lo, k_hi, gamma = hm.order_statistics(T, alpha)
        xs = np.sort(x)
        want = np.percentile(x, (1 - alpha) * 100)
        got = xs[k_lo] + (xs[k_hi] - xs[k_lo]) * gamma
        assert got == pytest.approx(want, rel=1e-12, abs=1e-15)


def test_the_random_cases_take_every_route():
    """The sweep above is only worth something if it walks up, walks down, uses the refined threshold and runs the tie loop."""
    rng = np.random.default_rng(77)
    seen = {"up": 0, "down": 0, "refined": 0, "tie_steps": 0, "exact": 0}
    sizes = []
    for case in range(300):
        x = rng.standard_normal(365).astype(np.float32) * 0.03
        if case % 3 == 0:
            x = np.round(x, 2)
        tr = {}
        hm.fast_select(x, 0.95, trace=tr)
        sizes.append(tr["first_set"])
        seen["up"] += tr["up"] > 0
        seen["down"] += tr["down"] > 0
        seen["refined"] += tr["refined"]
        seen["tie_steps"] += tr["tie_steps"] > 0
        seen["exact"] += tr["up"] == 0 and tr["down"] == 0
    assert min(seen[k] for k in ("up", "down", "refined", "tie_steps")) >= 10, seen
    assert 15 < np.mean(sizes) < 27                                 # the first set's size: around the 19 values alpha = 0.95 needs at T = 365
