"""CPU oracle for the Monte Carlo portfolio hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference
legs of ``bench.py`` may import it, and there only as the checker or as the
timed CPU baseline -- never as the thing shipped.  The product path
(``monte-carlo-portfolio_b200/``) must not import this package.

Parity status
-------------
The reference (``/root/reference/app.py``) ships no tests, golden vectors or
fixtures for this path, so nothing in the reference's own test-suite pins the
oracle.  What pins it instead:

* ``oracle/ref_loader.py`` ``exec``s the reference's own source lines
  (``app.py:231-284``: ``var``, ``cvar``, ``efficient_frontier`` ...) in this
  container and ``oracle/make_golden.py`` records their outputs under
  ``tests/golden/`` -- the restatement in ``reference_np.py`` is checked
  against those (``tests/test_oracle.py``).
* The pieces the reference does not implement at all (30 %-risk pick,
  correlated-path simulator, VaR/CVaR over simulated paths, frontier
  envelope, the Philox generator) are **parity unpinned**: the oracle there
  restates the north-star text with the reference's conventions
  (``app.py:258-263`` quantiles, ``app.py:253`` arithmetic compounding,
  ``app.py:709`` volatility).

* ``oracle/hist_select_model.py`` is not a restatement of the reference but of the KERNEL: the lane-by-lane selection of
  ``hist_var_fast`` in numpy, checked against a plain sort (``tests/test_hist_model_cpu.py``) so that the algorithm's set
  logic is pinned on the CPU tier as well.
"""
