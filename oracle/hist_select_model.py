"""Lane-by-lane numpy model of the selection inside ``hist_var_fast`` (monte-carlo-portfolio_b200/csrc/mcp_historical.cu).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): the product never imports this file.  The kernel finds the two order
statistics ``np.percentile`` interpolates between (app.py:258-259) and the set ``x <= VaR`` of the tail mean (app.py:261-263)
without sorting the series: 32 lanes own the periods ``lane, lane + 32, ...``, sort their own values, and

  1. one threshold ``tau`` = the minimum over lanes of the lanes' (row+1)-th smallest value defines a set S = every value
     ``<= tau`` among the lanes' first ``row`` values -- S is exactly the |S| smallest values of the series;
  2. a second threshold ``theta <= tau``, interpolated between the minimum of the lanes' row-th values and ``tau``, is
     counted too and replaces ``tau`` when S is too big and theta's set is nearer to the rank;
  3. single steps finish: the lowest lane holding the smallest untaken value takes it (set too small), or the lowest lane
     holding the largest taken value gives it back (set too big), until ``k_lo + 1`` values are taken;
  4. VaR is numpy's lerp between the largest taken and the smallest untaken value (FP32), further values are taken while
     the smallest untaken one is ``<= VaR`` (ties), and the tail mean adds each lane's taken values in ascending order,
     then the lanes' sums in a butterfly.

This model repeats those steps on one series with FP32 arithmetic, so that ``tests/test_hist_model_cpu.py`` can check the
ALGORITHM (which values are taken, for any threshold row, any ``shrink`` table, ties, tails concentrated in one lane)
against a plain sort -- on the CPU, where the GPU tier is not available.  The kernel itself is compared with the plain
kernel and with the FP64 oracle in ``tests/test_historical_gpu.py``.
"""
from __future__ import annotations

import numpy as np

F = np.float32
LANES = 32


def order_statistics(n_periods: int, alpha: float):
    """k_lo, k_hi, gamma of ``np.percentile(x, (1-alpha)*100)`` as hist_exec (mcp_historical.cu) derives them."""
    percent = (1 - alpha) * 100
    hidx = (n_periods - 1) * (percent / 100.0)
    if hidx >= n_periods - 1:
        return n_periods - 1, n_periods - 1, 0.0
    if hidx < 0:
        return 0, 0, 0.0
    k_lo = int(np.floor(hidx))
    return k_lo, k_lo + 1, hidx - np.floor(hidx)


def shrink_table(k_lo: int, row: int, size: int = 68):
    """The host-side table of hist_launch_fast: where between its two thresholds the refined one goes."""
    need, c_l = float(k_lo + 1), (7.0 if row == 2 else 1.0)
    return np.array([min(1.0, max(0.0, (need - c_l) / (c - c_l))) if c > c_l else 1.0 for c in range(size)], dtype=F)


def default_row(k_lo: int) -> int:
    return 2 if k_lo + 1 >= 12 else 1 if k_lo + 1 >= 5 else 0


def _butterfly_sum(v):
    v = np.array(v, dtype=F)
    for m in (16, 8, 4, 2, 1):
        v = (v + v[np.arange(LANES) ^ m]).astype(F)
    return v[0]


def fast_select(series, alpha: float, row: int | None = None, shrink=None, trace: dict | None = None):
    """(VaR, CVaR, taken) of one FP32 series the way hist_var_fast computes them; ``taken`` = how many values entered the tail mean.
    ``trace`` (optional dict) receives which route the selection took: 'refined', 'up' / 'down' step counts, 'tie_steps'."""
    x = np.asarray(series, dtype=F)
    T = x.shape[0]
    k_lo, k_hi, gamma = order_statistics(T, alpha)
    gamma = F(gamma)
    vpl = -(-T // LANES)
    row = default_row(k_lo) if row is None else row
    if row >= vpl:
        raise ValueError("the threshold row must exist in every lane")
    shrink = shrink_table(k_lo, row) if shrink is None else np.asarray(shrink, dtype=F)
    pad = np.full(vpl * LANES, np.inf, dtype=F)
    pad[:T] = x
    y = np.sort(pad.reshape(vpl, LANES).T, axis=1)                    # y[lane] = the lane's values, ascending (padding last)
    parked = np.concatenate([np.full((LANES, 2), -np.inf, F), y, np.full((LANES, 2), np.inf, F)], axis=1)
    PAD = 2
    need = k_lo + 1
    cnt = np.zeros(LANES, dtype=np.int64)
    c = 0
    if row > 0:
        tau = y[:, row].min()
        tau_l = y[:, row - 1].min()
        cnt1 = (y[:, :row] <= tau).sum(axis=1)
        c1 = int(cnt1.sum())
        with np.errstate(invalid="ignore"):
            theta = F(np.float64(tau - tau_l) * np.float64(shrink[c1]) + np.float64(tau_l))      # one rounding: fmaf
        theta = tau if not (theta <= tau) else theta                      # fminf(theta, tau) (a NaN theta gives tau)
        cnt2 = (y[:, :row] <= theta).sum(axis=1)
        c2 = int(cnt2.sum())
        use2 = c1 > need and abs(c2 - need) < c1 - need
        cnt, c = (cnt2, c2) if use2 else (cnt1, c1)
        cnt = cnt.astype(np.int64)
        if trace is not None:
            trace["refined"] = bool(use2)
            trace["first_set"] = c1
    if trace is not None:
        trace["up"], trace["down"] = max(0, need - c), max(0, c - need)
    if c <= k_lo:
        for _ in range(need - c):                                        # pops from below: the lowest lane holding the minimum
            h = parked[np.arange(LANES), PAD + cnt]
            cnt[int(np.argmin(h))] += 1                                  # argmin = first (lowest) lane
    else:
        for _ in range(c - need):                                        # removals from above: the lowest lane holding the maximum
            top = parked[np.arange(LANES), PAD + cnt - 1]
            cnt[int(np.argmax(top))] -= 1
    v_lo = parked[np.arange(LANES), PAD + cnt - 1].max()
    peek = parked[np.arange(LANES), PAD + cnt].min()
    v_hi = peek if k_hi > k_lo else v_lo
    diff = F(v_hi - v_lo)
    var = F(v_lo + F(diff * gamma))
    if gamma >= F(0.5):
        var = F(v_hi - F(diff * F(F(1) - gamma)))
    total = need
    while total < T and peek <= var:                                     # ties at the rank
        h = parked[np.arange(LANES), PAD + cnt]
        cnt[int(np.argmin(h))] += 1
        total += 1
        peek = parked[np.arange(LANES), PAD + cnt].min()
    if trace is not None:
        trace["tie_steps"] = total - need
    acc = np.zeros(LANES, dtype=F)
    for lane in range(LANES):                                            # each lane: its taken values in ascending order
        s = F(0)
        for v in range(int(cnt[lane])):
            s = F(s + y[lane, v])
        acc[lane] = s
    cvar = F(_butterfly_sum(acc) / F(total))
    return var, cvar, total


def sorted_reference(series, alpha: float):
    """The same quantities from a plain sort of the FP32 series: (v_lo, v_hi, VaR in FP32, number of values <= VaR, their FP64 mean)."""
    x = np.asarray(series, dtype=F)
    k_lo, k_hi, gamma = order_statistics(x.shape[0], alpha)
    gamma = F(gamma)
    xs = np.sort(x)
    v_lo, v_hi = xs[k_lo], xs[k_hi]
    diff = F(v_hi - v_lo)
    var = F(v_lo + F(diff * gamma))
    if gamma >= F(0.5):
        var = F(v_hi - F(diff * F(F(1) - gamma)))
    tail = x[x <= var]
    return v_lo, v_hi, var, int(tail.shape[0]), float(np.mean(tail.astype(np.float64)))
