"""GPU tier, round 2: tcgen05 path kernel, fused paths + VaR/CVaR call, wide universes, device mu / Sigma, the FP64 recheck
of the historical picks, pageable staging, stream ordering, the in-library communicator."""
import os

import numpy as np
import pytest

from conftest import synthetic_inputs
from oracle import paths_np, philox_np, reference_np as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# ---- tcgen05 path kernel ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,steps,M", [(16, 252, 2000), (16, 12, 1000), (5, 40, 777), (32, 20, 500), (21, 33, 300), (1, 7, 130)])
def test_tc_paths_match_the_generator_restatement_and_the_simt_kernel(mcp, n, steps, M):
    """The tensor-core contraction (TF32-split operands, FP32 accumulate) against the FP64 oracle on the SAME Philox normals:
    1e-4 relative on the terminal value (north star, FP32), and the SIMT kernel as a second witness."""
    mu, sigma = synthetic_inputs(n)
    w = np.random.default_rng(n).dirichlet(np.ones(n))
    first, seed = 5_000_000_000, 77
    a = mcp.simulate_paths(mu, sigma, w, M, steps, seed=seed, first_index=first, return_terminal=True)
    with _env(MCP_PATHS_TC="0"):
        b = mcp.simulate_paths(mu, sigma, w, M, steps, seed=seed, first_index=first, return_terminal=True)
    Z = philox_np.normals(first, M, steps, n, seed, "float32")
    want = paths_np.terminal_returns(mu, sigma, w, Z)
    assert np.allclose(a["terminal"] + 1.0, want + 1.0, rtol=1e-4)
    assert np.allclose(b["terminal"] + 1.0, want + 1.0, rtol=1e-4)
    assert np.allclose(a["terminal"] + 1.0, b["terminal"] + 1.0, rtol=2e-5)
    x = a["terminal"].astype(np.float64)
    for alpha, (v, c) in a["stats"].items():
        assert v == ref.var(x, alpha) and np.isclose(c, ref.cvar(x, alpha), rtol=1e-12)


@pytest.mark.parametrize("cfg", [dict(MCP_PATHS_TC_STAGES="2"), dict(MCP_PATHS_TC_WG="3", MCP_PATHS_TC_PPT="2"), dict(MCP_PATHS_TC_WG="7"),
                                 dict(MCP_PATHS_TC_MMA="1"), dict(MCP_PATHS_TC_WG="2", MCP_PATHS_TC_MMA="2")])
def test_tc_paths_variants_are_bit_identical(cfg):
    """Tile count, stages, paths per thread and issue warps are scheduling choices: a path's arithmetic does not depend on them.
    (The knobs are read once per process, so each variant runs in a child process.)"""
    import json
    import subprocess
    import sys
    code = ("import sys, json, numpy as np; sys.path[:0]=[%r, %r]; import mcportfolio as m; from conftest import synthetic_inputs;"
            "mu, s = synthetic_inputs(16); o = m.simulate_paths(mu, s, np.full(16, 1/16), 70001, 30, seed=3, first_index=12345, return_terminal=True);"
            "print(json.dumps({'sum': float(o['terminal'].astype('float64').sum()), 'h': int(np.bitwise_xor.reduce(o['terminal'].view('uint32')))}))")
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.join(os.path.dirname(here), "monte-carlo-portfolio_b200")
    outs = []
    for env in ({}, cfg):
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, "-c", code % (here, pkg)], capture_output=True, text=True, env=e, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1]


def test_paths_stats_single_call_equals_two_step(mcp):
    """mcp_paths_stats (kernel-filled first histogram, device-side interpolation, one copy back) against mcp_paths followed by
    mcp_quantiles on the same terminal values, and against numpy."""
    import torch
    mu, sigma = synthetic_inputs(16)
    w = np.full(16, 1 / 16)
    for M in (1, 2, 129, 300_001):
        o = mcp.simulate_paths(mu, sigma, w, M, 20, seed=5, return_terminal=True, alphas=(0.95, 0.99, 0.5, 0.0, 1.0))
        st2 = mcp.quantile_stats(o["terminal_device"], (0.95, 0.99, 0.5, 0.0, 1.0))
        assert o["stats"] == st2
        x = o["terminal"].astype(np.float64)
        for alpha, (v, c) in o["stats"].items():
            assert v == ref.var(x, alpha) and np.isclose(c, ref.cvar(x, alpha), rtol=1e-12, atol=1e-300)
    # without terminal values wanted: library scratch, same statistics
    a = mcp.simulate_paths(mu, sigma, w, 300_001, 20, seed=5, return_terminal=False)
    b = mcp.simulate_paths(mu, sigma, w, 300_001, 20, seed=5, return_terminal=True)
    assert a["terminal"] is None and a["stats"] == b["stats"]
    # FP64 takes the same route (no kernel-filled histogram there)
    d = mcp.simulate_paths(mu, sigma, w, 5000, 20, seed=5, return_terminal=True, dtype="float64")
    xd = d["terminal"]
    for alpha, (v, c) in d["stats"].items():
        assert v == ref.var(xd, alpha) and np.isclose(c, ref.cvar(xd, alpha), rtol=1e-12)
    del torch


def test_philox7_is_a_different_stream_with_the_same_distribution(mcp):
    mu, sigma = synthetic_inputs(16)
    w = np.full(16, 1 / 16)
    a = mcp.simulate_paths(mu, sigma, w, 400_000, 30, seed=9, return_terminal=True)["terminal"].astype(np.float64)
    b = mcp.simulate_paths(mu, sigma, w, 400_000, 30, seed=9, return_terminal=True, philox_rounds=7)["terminal"].astype(np.float64)
    assert not np.array_equal(a, b)
    se = a.std() / np.sqrt(len(a))
    assert abs(a.mean() - b.mean()) < 6 * se and abs(a.std() / b.std() - 1) < 0.01
    for q in (0.01, 0.05, 0.5, 0.95):
        assert abs(np.quantile(a, q) - np.quantile(b, q)) < 0.004
    with pytest.raises(mcp.McpError, match="philox_rounds"):
        mcp.simulate_paths(mu, sigma, w, 10, 3, philox_rounds=5)
    # the portfolio sweep takes the option too (7 rounds: other weights, same uniform-on-the-simplex law)
    r10 = mcp.simulate_portfolios(mu, sigma, 200_000, seed=3)
    r7 = mcp.simulate_portfolios(mu, sigma, 200_000, seed=3, philox_rounds=7)
    assert not np.array_equal(r10.weights, r7.weights)
    assert np.allclose(r7.weights.sum(1), 1, atol=1e-5) and abs(r7.weights.mean() - 1 / 16) < 1e-6
    assert np.allclose(r7.weights.var(0), 15 / (16 * 16 * 17), rtol=0.03)
    own = ref.evaluate(np.asarray(r7.weights, dtype=np.float64), mu, sigma, 0.0, 0.30)
    assert np.allclose(r7.risks, own["risks"], rtol=1e-4) and r7.max_sharpe["index"] == int(np.argmax(r7.sharpes))
    # value by value against the 7-round restatement: every sweep kernel family and the path kernels
    W7, _ = philox_np.dirichlet_weights(0, 3000, 16, seed=3, rounds=7)
    assert np.allclose(r7.weights[:3000], W7, atol=2e-6)
    sel = mcp.simulate_portfolios(mu, sigma, 200_000, seed=3, philox_rounds=7, return_arrays=False)     # packed, selection only
    assert sel.max_sharpe["global_index"] == r7.max_sharpe["index"] and sel.target_risk["global_index"] == r7.target_risk["index"]
    for n, dt in ((16, "float64"), (70, "float32"), (40, "float64"), (300, "float32")):
        mu_n, s_n = synthetic_inputs(n)
        lo = np.zeros(n) if n == 70 else None                      # N = 70 with (vacuous) bounds: the tiled SIMT kernel, else tcgen05
        rr = mcp.simulate_portfolios(mu_n, s_n, 500, seed=11, philox_rounds=7, dtype=dt, min_weights=lo)
        Wn, _ = philox_np.dirichlet_weights(0, 500, n, seed=11, dtype=dt, rounds=7)
        assert np.allclose(rr.weights, Wn, atol=2e-6 if dt == "float32" else 1e-12), (n, dt)
    Z7 = philox_np.normals(77, 300, 9, 16, 5, "float32", rounds=7)
    p7 = mcp.simulate_paths(mu, sigma, w, 300, 9, seed=5, first_index=77, return_terminal=True, philox_rounds=7)
    assert np.allclose(p7["terminal"] + 1.0, paths_np.terminal_returns(mu, sigma, w, Z7) + 1.0, rtol=1e-4)


# ---- wide universes (32 < N) -----------------------------------------------------------------------------------------

@pytest.mark.parametrize("dtype,tol", [("float64", 1e-6), ("float32", 1e-4)])
@pytest.mark.parametrize("n,steps", [(33, 9), (64, 30), (100, 12), (256, 20)])
def test_wide_paths_supplied_normals_parity(mcp, n, steps, dtype, tol):
    mu, sigma = synthetic_inputs(n, seed=4)
    rng = np.random.default_rng(n)
    M = 300
    Z = rng.standard_normal((M, steps, n))
    w = rng.dirichlet(np.ones(n))
    out = mcp.simulate_paths(mu, sigma, w, M, steps, normals=Z, dtype=dtype)
    Zc = Z.astype(np.float32).astype(np.float64) if dtype == "float32" else Z
    want = paths_np.terminal_returns(mu, sigma, w, Zc)
    assert np.allclose(out["terminal"] + 1.0, want + 1.0, rtol=tol)


@pytest.mark.parametrize("dtype,tol", [("float32", 2e-4), ("float64", 1e-9)])
def test_wide_paths_rng_matches_generator_restatement(mcp, dtype, tol):
    n, steps, M = 70, 6, 400
    mu, sigma = synthetic_inputs(n, seed=5)
    w = np.full(n, 1 / n)
    out = mcp.simulate_paths(mu, sigma, w, M, steps, seed=21, first_index=2**33 + 5, dtype=dtype)
    Z = philox_np.normals(2**33 + 5, M, steps, n, 21, dtype)
    want = paths_np.terminal_returns(mu, sigma, w, Z)
    assert np.allclose(out["terminal"] + 1, want + 1, rtol=tol)


# ---- mu / Sigma on the device (f2) -----------------------------------------------------------------------------------

def test_moments_kernel_matches_the_reference_lines(mcp, c1, c2):
    """app.py:679-680 on the device against the golden mu / Sigma the reference's own lines produced (1e-12), plus pandas."""
    import pandas as pd
    for g, A in ((c1, 52), (c2, 252)):
        mu, sigma = mcp.estimate_moments(g["returns_matrix"], A)
        assert np.allclose(mu, g["mu"], rtol=1e-12, atol=1e-15) and np.allclose(sigma, g["sigma"], rtol=1e-12, atol=1e-16)
        assert np.array_equal(sigma, sigma.T)
    rng = np.random.default_rng(0)
    R = rng.standard_normal((365, 37)) * 0.03
    R[0] = 0.0
    mu, sigma = mcp.estimate_moments(R, 12)
    df = pd.DataFrame(R)
    assert np.allclose(mu, df.mean().to_numpy() * 12, rtol=1e-12) and np.allclose(sigma, df.cov().to_numpy() * 12, rtol=1e-11, atol=1e-18)
    mu1, sigma1 = mcp.estimate_moments(R[:1], 12)                 # one period: pandas' covariance is NaN
    assert np.allclose(mu1, R[0] * 12) and np.isnan(sigma1).all()
    mu256, s256 = mcp.estimate_moments(rng.standard_normal((50, 256)), 4)
    assert s256.shape == (256, 256) and np.isfinite(s256).all()


# ---- historical picks: negated metric, FP64 recheck ------------------------------------------------------------------

def test_historical_negate_and_recheck(mcp):
    import torch
    rng = np.random.default_rng(3)
    T, n = 365, 16
    R = rng.standard_normal((T, n)) * 0.04
    base = rng.dirichlet(np.ones(n), size=40).astype(np.float32)
    # near-duplicates: every base row perturbed by a few float32 ulps in two weights -> FP32 VaR ties / flips, FP64 separates them
    rows = []
    for b in base:
        for k in range(50):
            r = b.copy()
            i, j = rng.integers(0, n, 2)
            r[i] = np.nextafter(r[i], 2, dtype=np.float32) if k % 2 else r[i]
            r[j] = np.nextafter(r[j], -1, dtype=np.float32) if k % 3 else r[j]
            rows.append(r)
    W = np.stack(rows)
    series = W.astype(np.float64) @ R.T
    var64 = np.array([ref.var(s, 0.95) for s in series])
    cvar64 = np.array([ref.cvar(s, 0.95) for s in series])
    for src in ("host", "device"):
        arg = torch.from_numpy(W).cuda() if src == "device" else W
        hv = mcp.historical_var_cvar(R, arg, 0.95, negate=True, recheck=True)
        plain = mcp.historical_var_cvar(R, arg, 0.95, negate=False, recheck=False)
        v, pv = (np.asarray(hv["var"].cpu()) if src == "device" else hv["var"]), (np.asarray(plain["var"].cpu()) if src == "device" else plain["var"])
        c, pc = (np.asarray(hv["cvar"].cpu()) if src == "device" else hv["cvar"]), (np.asarray(plain["cvar"].cpu()) if src == "device" else plain["cvar"])
        assert np.array_equal(v, -pv) and np.array_equal(c, -pc)                       # app.py:717: the metric is the negated value
        assert np.allclose(pv, var64, rtol=1e-4, atol=1e-7)
        assert hv["best_var"]["index"] == int(np.argmin(-var64))                       # app.py:673, 747 on the FP64 values
        assert hv["best_cvar"]["index"] == int(np.argmin(-cvar64))
        assert np.isclose(hv["best_var"]["value"], var64.max(), rtol=1e-12) and np.isclose(hv["best_cvar"]["value"], cvar64.max(), rtol=1e-12)


# ---- HOST space with pageable memory; stream ordering; argument combinations --------------------------

def test_pageable_host_buffers_are_staged(mcp):
    mu, sigma = synthetic_inputs(16)
    P = 1_500_000                                   # 96 MB of weights: several 16 MB staging chunks
    ref_run = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=2)            # pooled page-locked results
    out = {"weights": np.empty((P, 16), np.float32), "returns": np.empty(P, np.float32), "risks": np.empty(P, np.float32),
           "sharpes": np.empty(P, np.float32), "accepted": np.empty(P, np.uint8)}                # pageable
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=2, out=out)
    for k in ("weights", "returns", "risks", "sharpes"):
        assert np.array_equal(getattr(r, k), getattr(ref_run, k)), k
    assert r.max_sharpe["index"] == ref_run.max_sharpe["index"] and out["accepted"].all()
    # pageable supplied weights in, pageable arrays out
    W = np.array(ref_run.weights, dtype=np.float32, copy=True)
    s = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03, out=out)
    assert np.allclose(s.risks, ref_run.risks, rtol=2e-6) and s.max_sharpe["index"] == ref_run.max_sharpe["index"]


def test_pinned_pool_recycles_blocks(mcp):
    from mcportfolio import api
    a = mcp.pinned_empty((1 << 20,), np.float32)
    addr = a.ctypes.data
    a[:] = 1
    del a
    import gc
    gc.collect()
    b = mcp.pinned_empty((1 << 20,), np.float32)
    assert b.ctypes.data == addr                    # same block, no new cudaHostAlloc
    assert addr not in api._pinned_pool.free.get(4 << 20, [])   # ... and it left the idle list of its size class
    del b


def test_device_calls_are_ordered_behind_torch_default_stream(mcp):
    """a big dtype conversion queued on torch's default stream right before the call must be finished
    when libmcp's kernels read its output."""
    import torch
    mu, sigma = synthetic_inputs(16)
    P = 6_000_000
    g = torch.Generator(device="cuda").manual_seed(1)
    W64 = torch.rand((P, 16), dtype=torch.float64, device="cuda", generator=g)
    W64 /= W64.sum(1, keepdim=True)
    for _ in range(3):
        W = W64 * 1.0                                # fresh producer kernel on the default stream every round
        r = mcp.simulate_portfolios(mu, sigma, P, weights=W, risk_free=0.03, dtype="float32", return_arrays="device")
        sample = torch.cat([torch.arange(0, 2000), torch.arange(P - 2000, P)]).cuda()
        want = ref.evaluate(W64[sample].cpu().numpy(), mu, sigma, 0.03, 0.30)
        assert np.allclose(r.risks[sample].cpu().numpy(), want["risks"], rtol=1e-4)
        assert np.allclose(r.returns[sample].cpu().numpy(), want["returns"], rtol=1e-4)
        assert torch.isfinite(r.sharpes).all()


def test_host_weights_with_device_output_are_uploaded(mcp):
    """numpy weights + return_arrays='device' used to hand host pointers to a DEVICE-space call."""
    mu, sigma = synthetic_inputs(16)
    W = np.random.default_rng(0).dirichlet(np.ones(16), size=5000)
    r = mcp.simulate_portfolios(mu, sigma, 5000, weights=W, risk_free=0.03, return_arrays="device")
    want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    assert r.risks.is_cuda and np.allclose(r.risks.cpu().numpy(), want["risks"], rtol=1e-4)
    assert r.max_sharpe["index"] == want["max_sharpe"]["index"]
    m = mcp.simulate_portfolios(mu, sigma, 5000, weights=W, risk_free=0.03, return_arrays="device-metrics")
    assert np.allclose(m.returns.cpu().numpy(), want["returns"], rtol=1e-4)


def test_recheck_overflow_falls_back_to_a_full_fp64_pass(mcp):
    """More FP32 near-ties than the screen holds (here: 20000 identical rows): the pick must still be the FP64 argmax's first
    occurrence, and the call says that it took the slow route."""
    mu, sigma = synthetic_inputs(16)
    rng = np.random.default_rng(5)
    head = rng.dirichlet(np.ones(16) * 50, size=100)
    best = int(np.argmax(ref.evaluate(head, mu, sigma, 0.03, 0.30)["sharpes"]))
    W = np.tile(head[best], (20_000, 1))                           # 19 900 copies of the best row: more exact ties than RC_CAP = 8192
    W[:100] = head
    want = ref.evaluate(W, mu, sigma, 0.03, 0.30)
    assert want["max_sharpe"]["index"] == best
    for src in ("host", "device"):
        if src == "device":
            import torch
            arg = torch.from_numpy(W).cuda()
        else:
            arg = W
        r = mcp.simulate_portfolios(mu, sigma, len(W), weights=arg, risk_free=0.03, dtype="float32", return_arrays=False)
        assert r.max_sharpe["global_index"] == want["max_sharpe"]["index"]
        assert r.target_risk["global_index"] == want["target_risk"]["index"]
        assert r.extra.get("recheck_overflow") is True


# ---- the in-library communicator (one rank here; 2 ranks: test_multi_gpu.py / bench) ----------------------------------

def test_comm_world_of_one_merges_to_the_same_results(mcp):
    from mcportfolio import multi
    multi.close_all()                                # a devices=[...] call of an earlier test left engine 0 in a wider communicator
    eng = mcp.Engine(0)                              # a private handle: the shared engine's state stays untouched
    try:
        eng.comm_init(mcp.comm_unique_id(), 0, 1)
        assert eng.comm_info() == (0, 1)
        assert np.array_equal(eng.allgather(np.arange(5.0)), np.arange(5.0)[None])
        assert np.array_equal(eng.allreduce(np.array([3.0, -1.0]), "min"), [3.0, -1.0])
        assert np.array_equal(eng.allreduce(np.array([7, 9], dtype=np.uint64), "sum"), [7, 9])
        with pytest.raises(mcp.McpError, match="already has a communicator"):
            eng.comm_init(mcp.comm_unique_id(), 0, 1)
    finally:
        eng.close()
    # through the public API on the shared engine
    from mcportfolio import api
    mu, sigma = synthetic_inputs(16)
    shared = mcp.get_engine(0)
    if shared.comm_info()[1] == 0:
        shared.comm_init(mcp.comm_unique_id(), 0, 1)
    a = mcp.simulate_portfolios(mu, sigma, 100_001, risk_free=0.03, seed=8, return_arrays=False, n_bins=16, risk_range=(0.1, 0.5))
    b = mcp.simulate_portfolios(mu, sigma, 100_001, risk_free=0.03, seed=8, return_arrays=False, n_bins=16, risk_range=(0.1, 0.5), comm_merge=True)
    for pick in ("max_sharpe", "target_risk"):
        assert getattr(a, pick)["global_index"] == getattr(b, pick)["global_index"] and getattr(a, pick)["sharpe"] == getattr(b, pick)["sharpe"]
        assert np.array_equal(getattr(a, pick)["weights"], getattr(b, pick)["weights"])
    assert b.extra["n_accepted_global"] == 100_001 and a.risk_range == b.risk_range
    assert np.array_equal(a.extra["envelope"]["best_index"], b.extra["envelope"]["best_index"])
    assert np.array_equal(a.extra["envelope"]["best_return"], b.extra["envelope"]["best_return"])
    w = a.max_sharpe["weights"]
    pa = mcp.simulate_paths(mu, sigma, w, 50_000, 16, seed=8, return_terminal=False)
    pb = mcp.simulate_paths(mu, sigma, w, 50_000, 16, seed=8, return_terminal=False, comm_merge=True, n_total=50_000)
    assert pa["stats"] == pb["stats"]
    x = np.random.default_rng(1).standard_normal(10_001).astype(np.float32)
    assert mcp.quantile_stats(x, (0.95, 0.99)) == mcp.quantile_stats(x, (0.95, 0.99), allreduce="comm", n_total=len(x))
    shared.comm_destroy()
    with pytest.raises(mcp.McpError, match="communicator"):
        mcp.simulate_portfolios(mu, sigma, 10, comm_merge=True)
    del api


# ---- bounds rejection with the quadratic forms on the tensor cores (32 < N <= 256) ------------------------------------------

@pytest.mark.parametrize("n,lo,hi,tries", [(64, None, 0.07, 3), (64, 1e-4, None, 5), (256, None, 0.022, 4), (100, 2e-5, 0.06, 100),
                                           (64, None, 0.045, 2)])
def test_bounded_sweep_on_tensor_cores_equals_the_simt_sweep(mcp, n, lo, hi, tries):
    """app.py:700-707 at N > 32: the tcgen05 kernel accepts what is safely inside the bounds at attempt 0 and hands the rest to
    the tiled SIMT kernel.  Accept / skip decisions, attempts and therefore the weights are the SIMT-only sweep's; the metrics
    carry the tensor-core kernel's FP16-split arithmetic (1e-4 against FP64).  Rejected rows are redrawn in further tensor-core
    rounds (attempt + 1); the last case rejects almost everything, so most rows end skipped after two attempts."""
    mu, sigma = synthetic_inputs(n)
    P, seed = 70_001, 11
    kw = dict(min_weights=None if lo is None else np.full(n, lo), max_weights=None if hi is None else np.full(n, hi), seed=seed,
              max_tries=tries, risk_free=0.03)
    a = mcp.simulate_portfolios(mu, sigma, P, **kw)
    with _env(MCP_LARGE_TC_BOUNDS="0"):
        b = mcp.simulate_portfolios(mu, sigma, P, **kw)
    assert 0 < b.n_accepted and (b.n_accepted < P or tries == 100) and a.n_accepted == b.n_accepted     # 100 tries: everything lands
    assert np.array_equal(a.accepted, b.accepted)
    assert np.allclose(a.weights, b.weights, rtol=0, atol=3e-7)
    want = ref.evaluate(np.asarray(a.weights, dtype=np.float64), mu, sigma, 0.03, 0.30)
    assert np.allclose(a.risks, want["risks"], rtol=1e-4) and np.allclose(a.returns, want["returns"], rtol=1e-4)
    assert np.allclose(a.sharpes, want["sharpes"], rtol=1e-4, atol=1e-4)
    if lo is not None:
        assert a.weights.min() >= np.float32(lo) * (1 - 1e-6)
    if hi is not None:
        assert a.weights.max() <= np.float32(hi) * (1 + 1e-6)
    assert a.max_sharpe["index"] == int(np.argmax(a.sharpes)) and a.target_risk["index"] == int(np.argmin(np.abs(a.risks - np.float32(0.30))))
    # value by value against the generator restatement with the reference's rejection loop (a sample)
    W, valid = philox_np.dirichlet_weights(0, 3000, n, seed, "float32", kw["min_weights"], kw["max_weights"], max_tries=tries)
    acc = a.accepted[:3000].astype(bool)
    edge = np.full(3000, np.inf)
    if hi is not None:
        edge = np.minimum(edge, np.abs(W - hi).min(1))
    if lo is not None:
        edge = np.minimum(edge, np.abs(W - lo).min(1))
    differ = acc != valid
    assert differ.sum() <= 3 and np.all(edge[differ] < 1e-6)
    both = acc & valid
    assert np.allclose(a.weights[np.cumsum(acc)[both] - 1], W[both], atol=1e-6)
    # selections only (no arrays), and device-resident arrays, take the same route
    s = mcp.simulate_portfolios(mu, sigma, P, return_arrays=False, **kw)
    assert s.n_accepted == a.n_accepted and s.max_sharpe["global_index"] == a.max_sharpe["global_index"]
    assert s.target_risk["global_index"] == a.target_risk["global_index"]
    d = mcp.simulate_portfolios(mu, sigma, P, return_arrays="device", **kw)
    assert d.n_accepted == a.n_accepted and np.array_equal(d.sharpes.cpu().numpy(), a.sharpes)


def test_bounded_tensor_core_sweep_sharding_and_host_pipeline(mcp):
    """The bounded route inside the HOST-space chunk pipeline (two streams, several chunks) and across index shards."""
    n = 64
    mu, sigma = synthetic_inputs(n)
    hi = np.full(n, 0.08)
    P = 700_001                                           # 183 MB of weights: several pipeline chunks
    whole = mcp.simulate_portfolios(mu, sigma, P, max_weights=hi, seed=2, max_tries=4, risk_free=0.03)
    assert 0 < whole.n_accepted < P
    sel = mcp.simulate_portfolios(mu, sigma, P, max_weights=hi, seed=2, max_tries=4, risk_free=0.03, return_arrays=False)
    assert sel.n_accepted == whole.n_accepted and sel.max_sharpe["global_index"] == whole.max_sharpe["global_index"]
    cuts = [0, 40_000, 300_001, P]
    parts = [mcp.simulate_portfolios(mu, sigma, b - a, max_weights=hi, seed=2, max_tries=4, risk_free=0.03, first_index=a, return_arrays=False)
             for a, b in zip(cuts[:-1], cuts[1:])]
    assert sum(p.n_accepted for p in parts) == whole.n_accepted
    best = max(parts, key=lambda r: (r.max_sharpe["key"], -r.max_sharpe["global_index"]))
    assert best.max_sharpe["global_index"] == whole.max_sharpe["global_index"]
    with _env(MCP_LARGE_TC_BOUNDS="0"):
        simt = mcp.simulate_portfolios(mu, sigma, P, max_weights=hi, seed=2, max_tries=4, risk_free=0.03, return_arrays=False)
    assert simt.n_accepted == whole.n_accepted


# ---- tcgen05 path kernel for 32 < N <= 128 (normals drawn and stored 16 at a time) and, in blocks of 64 assets, up to 256 -----

@pytest.mark.parametrize("n,steps,M", [(33, 9, 300), (64, 30, 700), (100, 12, 1000), (128, 20, 500), (70, 252, 260),
                                       (129, 8, 300), (200, 15, 400), (256, 20, 300), (192, 252, 140)])
def test_wide_tc_paths_match_the_generator_restatement_and_the_simt_kernel(mcp, n, steps, M):
    mu, sigma = synthetic_inputs(n, seed=6)
    w = np.random.default_rng(n).dirichlet(np.ones(n))
    first, seed = 3_000_000_000, 41
    a = mcp.simulate_paths(mu, sigma, w, M, steps, seed=seed, first_index=first, return_terminal=True)
    with _env(MCP_PATHS_TC="0"):
        b = mcp.simulate_paths(mu, sigma, w, M, steps, seed=seed, first_index=first, return_terminal=True)
    Z = philox_np.normals(first, M, steps, n, seed, "float32")
    want = paths_np.terminal_returns(mu, sigma, w, Z)
    assert np.allclose(a["terminal"] + 1.0, want + 1.0, rtol=1e-4)            # north star, FP32
    assert np.allclose(b["terminal"] + 1.0, want + 1.0, rtol=1e-4)
    assert np.allclose(a["terminal"] + 1.0, b["terminal"] + 1.0, rtol=3e-5)
    with _env(MCP_PATHS_TC_WIDE16="0"):                                       # the one-stage TF32-split kernel: a third witness
        t32 = mcp.simulate_paths(mu, sigma, w, M, steps, seed=seed, first_index=first, return_terminal=True)
    assert np.allclose(t32["terminal"] + 1.0, want + 1.0, rtol=1e-4) and np.allclose(a["terminal"] + 1.0, t32["terminal"] + 1.0, rtol=3e-5)
    x = a["terminal"].astype(np.float64)
    for alpha, (v, c) in a["stats"].items():                                  # the kernel-filled first histogram feeds the exact select
        assert v == ref.var(x, alpha) and np.isclose(c, ref.cvar(x, alpha), rtol=1e-12)
    # shards of the same seed reproduce the same paths bit for bit; 7 rounds is another stream
    k = min(130, M - 129)
    part = mcp.simulate_paths(mu, sigma, w, k, steps, seed=seed, first_index=first + 129, return_terminal=True)["terminal"]
    assert np.array_equal(part, a["terminal"][129:129 + k])
    if steps <= 30:
        Z7 = philox_np.normals(first, 200, steps, n, seed, "float32", rounds=7)
        p7 = mcp.simulate_paths(mu, sigma, w, 200, steps, seed=seed, first_index=first, return_terminal=True, philox_rounds=7)
        assert np.allclose(p7["terminal"] + 1.0, paths_np.terminal_returns(mu, sigma, w, Z7) + 1.0, rtol=1e-4)
