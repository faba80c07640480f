"""GPU tier: the N > 1 paths.  Two ranks on ONE GPU over gloo (host-side merge code against the single-rank result; runs on the
one-GPU test box), and -- when the box has two GPUs -- two NCCL ranks through libmcp's communicator and the single-process
`devices=[...]` mode."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import synthetic_inputs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run_ranks(backend, nproc, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / f"{backend}.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "_dist_worker.py"), backend, str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    with open(out) as fh:
        return json.load(fh)


def _check(d):
    assert d["ranks_agree"]
    s, one = d["sharded"], d["single"]
    for k in ("picks", "n_acc", "stats", "env_idx", "env_ret", "env_pick", "bounded"):
        assert s[k] == one[k], k


def test_two_ranks_on_one_gpu_match_single_rank(tmp_path):
    """World size 2 on cuda:0 (gloo): sharded picks, counts, VaR / CVaR and envelope bins identical to the one-rank run."""
    import mcportfolio
    mcportfolio.build()
    d = _run_ranks("gloo", 2, tmp_path)
    assert d["world"] == 2
    _check(d)


def _two_gpus():
    import torch
    return torch.cuda.device_count() >= 2


def test_two_nccl_ranks_merge_inside_libmcp(tmp_path):
    if not _two_gpus():
        pytest.skip("needs two GPUs")
    d = _run_ranks("nccl", 2, tmp_path)
    assert d["world"] == 2 and d["comm"] == [0, 2]          # the engine joined libmcp's own communicator
    _check(d)


def test_single_process_devices_mode(tmp_path):
    """devices=[0, 1]: one process, a handle + host thread per GPU, NCCL communicator inside libmcp."""
    if not _two_gpus():
        pytest.skip("needs two GPUs")
    import mcportfolio as mcp
    mcp.build()
    mu, sigma = synthetic_inputs(16)
    P = 2_000_003
    a = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=4, return_arrays=False)
    b = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=4, return_arrays=False, devices=[0, 1])
    for pick in ("max_sharpe", "target_risk"):
        assert getattr(a, pick)["global_index"] == getattr(b, pick)["global_index"] and getattr(a, pick)["sharpe"] == getattr(b, pick)["sharpe"]
        assert np.array_equal(getattr(a, pick)["weights"], getattr(b, pick)["weights"])
    assert b.n_accepted == P and a.risk_range == b.risk_range and b.extra["devices"] == [0, 1]
    # arrays: every device fills its slice of one allocation
    fa = mcp.simulate_portfolios(mu, sigma, 300_001, risk_free=0.03, seed=6)
    fb = mcp.simulate_portfolios(mu, sigma, 300_001, risk_free=0.03, seed=6, devices=[0, 1])
    for k in ("weights", "returns", "risks", "sharpes"):
        assert np.array_equal(getattr(fa, k), getattr(fb, k)), k
    assert fa.max_sharpe["index"] == fb.max_sharpe["index"] and fa.target_risk["index"] == fb.target_risk["index"]
    # bounds: skipped portfolios, compacted arrays, positions
    lo, hi = np.full(16, 0.01), np.full(16, 0.2)
    ba = mcp.simulate_portfolios(mu, sigma, 40_000, risk_free=0.03, seed=5, min_weights=lo, max_weights=hi)
    bb = mcp.simulate_portfolios(mu, sigma, 40_000, risk_free=0.03, seed=5, min_weights=lo, max_weights=hi, devices=[0, 1])
    assert ba.n_accepted == bb.n_accepted < 40_000 and np.array_equal(ba.weights, bb.weights) and np.array_equal(ba.sharpes, bb.sharpes)
    assert ba.max_sharpe["index"] == bb.max_sharpe["index"] == int(np.argmax(bb.sharpes))
    # supplied weights
    W = np.random.default_rng(0).dirichlet(np.ones(16), size=100_001)
    sa = mcp.simulate_portfolios(mu, sigma, len(W), weights=W, risk_free=0.03)
    sb = mcp.simulate_portfolios(mu, sigma, len(W), weights=W, risk_free=0.03, devices=[0, 1])
    assert sa.max_sharpe["index"] == sb.max_sharpe["index"] and np.array_equal(sa.risks, sb.risks)
    # paths
    w = a.max_sharpe["weights"]
    pa = mcp.simulate_paths(mu, sigma, w, 500_003, 32, seed=4, return_terminal=False)
    pb = mcp.simulate_paths(mu, sigma, w, 500_003, 32, seed=4, devices=[0, 1])
    assert pa["stats"] == pb["stats"]
    # envelope
    mu2, sigma2 = synthetic_inputs(64)
    ea = mcp.frontier_envelope(mu2, sigma2, 200_000, 32, risk_free=0.03, seed=4)
    eb = mcp.frontier_envelope(mu2, sigma2, 200_000, 32, risk_free=0.03, seed=4, devices=[0, 1])
    assert np.array_equal(ea.extra["envelope"]["best_index"], eb.extra["envelope"]["best_index"])
    assert np.array_equal(ea.extra["envelope"]["best_return"], eb.extra["envelope"]["best_return"])
    assert ea.target_risk["global_index"] == eb.target_risk["global_index"]
