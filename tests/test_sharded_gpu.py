"""GPU tier: the sharded entry points over a real NCCL process group (world size 1 on the test
box; the 2/4/8-rank runs are bench.py under torchrun -- profiles/r1c_bench_n*.json), and the C2
real-data configuration."""
import socket

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


@pytest.fixture(scope="module")
def nccl_group():
    import torch
    import torch.distributed as dist
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    yield dist
    dist.destroy_process_group()


def test_sharded_entry_points_match_single_process(mcp, nccl_group):
    from mcportfolio import dist as mdist
    mu, sigma = synthetic_inputs(16)
    P = 3_000_001
    a = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=4, return_arrays=False)
    b = mdist.simulate_portfolios_sharded(mu, sigma, P, risk_free=0.03, seed=4, return_arrays=False)
    for pick in ("max_sharpe", "target_risk"):
        ra, rb = getattr(a, pick), getattr(b, pick)
        assert ra["global_index"] == rb["global_index"] and ra["sharpe"] == rb["sharpe"]
        assert np.array_equal(ra["weights"], rb["weights"])
    assert b.extra["n_accepted_global"] == P
    # paths: the all-reduce callback path (NCCL sum of the radix histograms) gives the same exact VaR
    w = a.max_sharpe["weights"]
    pa = mcp.simulate_paths(mu, sigma, w, 200_000, 32, seed=4)
    import torch.distributed as dist
    eng = mcp.get_engine(0)
    pb = mcp.simulate_paths(mu, sigma, w, 200_000, 32, seed=4, allreduce=mdist.make_allreduce(eng.device), n_total=200_000)
    assert pa["stats"] == pb["stats"]
    x = pa["terminal"].astype(np.float64)
    for alpha, (v, c) in pb["stats"].items():
        assert v == ref.var(x, alpha) and np.isclose(c, ref.cvar(x, alpha), rtol=1e-12)
    # envelope
    mu2, sigma2 = synthetic_inputs(64)
    ea = mcp.frontier_envelope(mu2, sigma2, 100_000, 32, risk_free=0.03, seed=4)
    eb = mdist.frontier_envelope_sharded(mu2, sigma2, 100_000, 32, risk_free=0.03, seed=4)
    assert np.array_equal(ea.extra["envelope"]["best_index"], eb.extra["envelope"]["best_index"])
    assert np.array_equal(ea.extra["envelope"]["best_return"], eb.extra["envelope"]["best_return"])
    assert ea.target_risk["global_index"] == eb.target_risk["global_index"]


def test_c2_real_data_1e6_portfolios(mcp, c2):
    """Config C2 (policy of SURVEY 8(d): the 14 non-weekly files, 24 common dates, N = 14):
    1e6 in-kernel portfolios against host numpy on the same mu / Sigma."""
    mu, sigma = c2["mu"], c2["sigma"]
    N, P = 14, 1_000_000
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=0)
    assert r.weights.shape == (P, N) and r.n_accepted == P
    own = ref.evaluate(np.asarray(r.weights, dtype=np.float64), mu, sigma, 0.03, 0.30)
    assert np.allclose(r.risks, own["risks"], rtol=1e-4) and np.allclose(r.sharpes, own["sharpes"], rtol=1e-4, atol=1e-4)
    assert r.max_sharpe["index"] == int(np.argmax(r.sharpes))
    assert r.target_risk["index"] == int(np.argmin(np.abs(r.risks - np.float32(0.30))))
    Wn = np.random.RandomState(0).dirichlet(np.ones(N), size=P)            # the reference's sampler (app.py:702)
    _, risk_n, sharpe_n = ref.portfolio_metrics(Wn, mu, sigma, 0.03)
    for q in (0.001, 0.01, 0.25, 0.5, 0.75, 0.99, 0.999):
        assert np.isclose(np.quantile(r.risks, q), np.quantile(risk_n, q), rtol=1e-2)
        assert np.isclose(np.quantile(r.sharpes, q), np.quantile(sharpe_n, q), rtol=2e-2, atol=2e-3)
    assert np.isclose(r.risks.min(), risk_n.min(), rtol=0.05) and np.isclose(r.sharpes.max(), sharpe_n.max(), rtol=0.05)
