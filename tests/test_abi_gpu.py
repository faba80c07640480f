"""GPU tier: the raw C ABI (include/mcp.h) as a foreign binder would drive it -- plain ctypes, no Python
host layer in between.  Error behaviour first: every bad call returns a negative class, leaves a message in
mcp_last_error, never throws / exits, and the handle stays usable; then one good call per entry point against
the oracle so that the ABI itself (struct layout, ownership, HOST and DEVICE space) is what is pinned."""
import ctypes as C

import numpy as np
import pytest

from conftest import synthetic_inputs
from oracle import reference_np as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    import mcportfolio
    from mcportfolio import _lib
    mcportfolio.build()
    L = _lib.lib()
    h = C.c_void_p()
    assert L.mcp_create(0, C.byref(h)) == 0
    yield L, h, _lib
    assert L.mcp_destroy(h) == 0


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def make_params(_lib, n, P, **kw):
    p = _lib.PortfolioParams()
    p.n_assets, p.dtype, p.n_portfolios, p.first_index, p.seed = n, _lib.MCP_F32, P, 0, 0
    p.risk_free, p.risk_target, p.max_tries, p.keep_last, p.space = 0.03, 0.30, 100, 0, _lib.MCP_HOST
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def last_error(L, h):
    return L.mcp_last_error(h).decode()


def test_bad_arguments_return_error_classes_and_keep_the_handle_alive(abi):
    L, h, _lib = abi
    n = 16
    mu, sigma = synthetic_inputs(n, seed=0)
    out = _lib.PortfolioOut()
    good = make_params(_lib, n, 1000)

    assert L.mcp_portfolios(None, C.byref(good), ptr(mu), ptr(sigma), C.byref(out)) == _lib.MCP_ERR_INVALID
    assert L.mcp_portfolios(h, None, ptr(mu), ptr(sigma), C.byref(out)) == _lib.MCP_ERR_INVALID
    assert "NULL" in last_error(L, h)
    assert L.mcp_portfolios(h, C.byref(good), None, ptr(sigma), C.byref(out)) == _lib.MCP_ERR_INVALID
    assert L.mcp_portfolios(h, C.byref(good), ptr(mu), ptr(sigma), None) == _lib.MCP_ERR_INVALID
    for field, value, word in (("n_assets", 0, "n_assets"), ("n_assets", 5000, "n_assets"), ("dtype", 7, "dtype"),
                               ("space", 3, "space"), ("max_tries", 0, "max_tries"), ("n_portfolios", 1 << 40, "2^39"),
                               ("n_bins", -1, "n_bins"), ("n_bins", 1 << 20, "n_bins")):
        bad = make_params(_lib, n, 1000, **{field: value})
        assert L.mcp_portfolios(h, C.byref(bad), ptr(mu), ptr(sigma), C.byref(out)) == _lib.MCP_ERR_INVALID, field
        assert word in last_error(L, h), (field, last_error(L, h))
    # envelope requested without bin buffers / with an empty range
    bad = make_params(_lib, n, 1000, n_bins=8, risk_lo=0.1, risk_hi=0.5)
    assert L.mcp_portfolios(h, C.byref(bad), ptr(mu), ptr(sigma), C.byref(out)) == _lib.MCP_ERR_INVALID
    rets, idx = np.empty(8), np.empty(8, dtype=np.uint64)
    out.bin_best_return, out.bin_best_index = ptr(rets), ptr(idx)
    bad = make_params(_lib, n, 1000, n_bins=8, risk_lo=0.5, risk_hi=0.5)
    assert L.mcp_portfolios(h, C.byref(bad), ptr(mu), ptr(sigma), C.byref(out)) == _lib.MCP_ERR_INVALID
    out.bin_best_return = out.bin_best_index = None
    # non-finite inputs are rejected by name
    mu_nan = mu.copy(); mu_nan[3] = np.nan
    assert L.mcp_portfolios(h, C.byref(good), ptr(mu_nan), ptr(sigma), C.byref(out)) == _lib.MCP_ERR_INVALID
    assert "mean_returns[3]" in last_error(L, h)
    s_inf = sigma.copy(); s_inf[2, 5] = np.inf
    assert L.mcp_portfolios(h, C.byref(good), ptr(mu), ptr(s_inf), C.byref(out)) == _lib.MCP_ERR_INVALID
    assert "cov_matrix[2,5]" in last_error(L, h)
    lo = np.zeros(n); lo[1] = np.nan
    bad = make_params(_lib, n, 1000, min_weights=ptr(lo))
    assert L.mcp_portfolios(h, C.byref(bad), ptr(mu), ptr(sigma), C.byref(out)) == _lib.MCP_ERR_INVALID

    # paths: argument classes and the numeric class for a covariance that is not positive definite
    pp = _lib.PathParams()
    pp.n_assets, pp.dtype, pp.n_paths, pp.n_steps, pp.space, pp.dt = n, _lib.MCP_F32, 100, 10, _lib.MCP_HOST, 1 / 252
    w = np.full(n, 1 / n)
    term = np.empty(100, dtype=np.float32)
    ms = C.c_double()
    assert L.mcp_paths(h, C.byref(pp), ptr(mu), ptr(sigma), None, ptr(term), C.byref(ms)) == _lib.MCP_ERR_INVALID
    for field, value in (("n_assets", 1025), ("n_steps", 0), ("dt", 0.0), ("dt", float("nan")), ("dtype", 2)):
        old = getattr(pp, field)
        setattr(pp, field, value)
        assert L.mcp_paths(h, C.byref(pp), ptr(mu), ptr(sigma), ptr(w), ptr(term), C.byref(ms)) == _lib.MCP_ERR_INVALID, field
        setattr(pp, field, old)
    not_pd = sigma.copy(); not_pd[0, 0] = -1.0
    assert L.mcp_paths(h, C.byref(pp), ptr(mu), ptr(not_pd), ptr(w), ptr(term), C.byref(ms)) == _lib.MCP_ERR_NUMERIC
    assert "positive definite" in last_error(L, h)

    # envelope over arrays
    assert L.mcp_envelope_arrays(h, 0, None, None, 10, 0, 0.1, 0.5, 8, ptr(rets), ptr(idx)) == _lib.MCP_ERR_INVALID
    assert L.mcp_envelope_arrays(h, 0, None, None, 0, 0, 0.1, 0.5, 0, ptr(rets), ptr(idx)) == _lib.MCP_ERR_INVALID
    assert L.mcp_envelope_arrays(h, 0, None, None, 0, 0, 0.1, 0.5, 8, ptr(rets), ptr(idx)) == 0      # empty input: empty bins
    assert np.all(np.isneginf(rets)) and np.all(idx == _lib.MCP_NO_INDEX)

    # the handle survived all of it: a good call right after
    assert L.mcp_portfolios(h, C.byref(good), ptr(mu), ptr(sigma), C.byref(out)) == 0
    assert out.n_accepted == 1000 and out.max_sharpe.index < 1000 and np.isfinite(out.max_sharpe.sharpe)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_supplied_weights_through_the_raw_abi_host_space(abi, dtype):
    """Caller-owned host buffers in, caller-owned host buffers out, selection records with weights."""
    L, h, _lib = abi
    n, P = 7, 5000
    mu, sigma = synthetic_inputs(n, seed=3)
    W = np.random.RandomState(0).dirichlet(np.ones(n), size=P)
    np_t = np.float32 if dtype == "float32" else np.float64
    w_in = np.ascontiguousarray(W, dtype=np_t)
    p = make_params(_lib, n, P, dtype=_lib.MCP_F32 if dtype == "float32" else _lib.MCP_F64, weights_in=ptr(w_in),
                    risk_free=0.01, risk_target=0.25)
    risks, rets, sharpes = (np.full(P, np.nan, dtype=np_t) for _ in range(3))
    acc = np.zeros(P, dtype=np.uint8)
    ws, wt = np.empty(n), np.empty(n)
    out = _lib.PortfolioOut()
    out.risks, out.returns, out.sharpes, out.accepted = ptr(risks), ptr(rets), ptr(sharpes), ptr(acc)
    out.max_sharpe.weights, out.target_risk.weights = ptr(ws), ptr(wt)
    assert L.mcp_portfolios(h, C.byref(p), ptr(mu), ptr(sigma), C.byref(out)) == 0, last_error(L, h)
    want = ref.evaluate(w_in.astype(np.float64), mu, sigma, 0.01, 0.25)
    tol = 1e-4 if dtype == "float32" else 1e-6                 # north-star tolerances
    assert out.n_accepted == P and acc.all()
    assert np.allclose(risks, want["risks"], rtol=tol) and np.allclose(rets, want["returns"], rtol=tol)
    assert np.allclose(sharpes, want["sharpes"], rtol=tol, atol=tol)
    i, j = out.max_sharpe.index, out.target_risk.index
    assert risks[i] == np_t(out.max_sharpe.risk) and sharpes[i] == np_t(out.max_sharpe.sharpe)
    assert sharpes[i] >= sharpes.max() and abs(risks[j] - np_t(0.25)) <= np.abs(risks - np_t(0.25)).min()
    if dtype == "float64":
        assert i == want["max_sharpe"]["index"] and j == want["target_risk"]["index"]
    assert np.allclose(ws, w_in[i], atol=1e-7) and np.allclose(wt, w_in[j], atol=1e-7)
    assert out.risk_min == risks.min() and out.risk_max == risks.max()


def test_empty_range_and_quantile_edge_cases_through_the_raw_abi(abi):
    L, h, _lib = abi
    n = 4
    mu, sigma = synthetic_inputs(n, seed=1)
    out = _lib.PortfolioOut()
    p = make_params(_lib, n, 0)
    assert L.mcp_portfolios(h, C.byref(p), ptr(mu), ptr(sigma), C.byref(out)) == 0
    assert out.n_accepted == 0 and out.max_sharpe.index == _lib.MCP_NO_INDEX and np.isnan(out.risk_min)
    pp = _lib.PathParams()
    pp.n_assets, pp.dtype, pp.n_paths, pp.n_steps, pp.space, pp.dt = n, _lib.MCP_F32, 0, 10, _lib.MCP_HOST, 1 / 252
    ms = C.c_double(-1)
    term = np.empty(1, dtype=np.float32)
    assert L.mcp_paths(h, C.byref(pp), ptr(mu), ptr(sigma), ptr(np.full(n, .25)), ptr(term), C.byref(ms)) == 0 and ms.value == 0
    assert L.mcp_abi_version() >= 1
    info = _lib.DeviceInfo()
    assert L.mcp_device_info(h, C.byref(info)) == 0 and info.cc_major == 10 and info.sm_count > 0


def _multi_job(L, _lib, handles, n_dev, mu, sigma, P, want_arrays):
    n = len(mu)
    arr = (C.c_void_p * n_dev)(*handles)
    p = make_params(_lib, n, P, seed=5)
    out = _lib.PortfolioOut()
    ws, wt = np.empty(n), np.empty(n)
    out.max_sharpe.weights, out.target_risk.weights = ptr(ws), ptr(wt)
    bufs = {}
    if want_arrays:
        bufs = {"weights": np.empty((P, n), np.float32), "returns": np.empty(P, np.float32), "risks": np.empty(P, np.float32),
                "sharpes": np.empty(P, np.float32), "accepted": np.zeros(P, np.uint8)}
        for k, v in bufs.items():
            setattr(out, k, ptr(v))
    ms = np.zeros(n_dev)
    rc = L.mcp_portfolios_multi(arr, n_dev, C.byref(p), ptr(mu), ptr(sigma), C.byref(out), ptr(ms))
    assert rc == 0, last_error(L, handles[0])
    return out, ws, wt, bufs, ms


def test_multi_entry_points_with_one_handle_equal_the_plain_calls(abi):
    """mcp_portfolios_multi / mcp_paths_stats_multi with n = 1 (what a one-GPU box can run): same results as the plain entry points."""
    L, h, _lib = abi
    n = 16
    mu, sigma = synthetic_inputs(n, seed=0)
    P = 200_001
    out, ws, wt, bufs, ms = _multi_job(L, _lib, [h], 1, mu, sigma, P, True)
    p = make_params(_lib, n, P, seed=5)
    ref_out = _lib.PortfolioOut()
    assert L.mcp_portfolios(h, C.byref(p), ptr(mu), ptr(sigma), C.byref(ref_out)) == 0
    assert out.max_sharpe.index == ref_out.max_sharpe.index and out.target_risk.index == ref_out.target_risk.index
    assert out.n_accepted == P and bufs["accepted"].all() and ms[0] > 0
    assert int(np.argmax(bufs["sharpes"])) == out.max_sharpe.index and np.allclose(bufs["weights"].sum(1), 1, atol=1e-5)
    pp = _lib.PathParams()
    pp.n_assets, pp.dtype, pp.n_paths, pp.first_index, pp.seed, pp.n_steps, pp.space, pp.dt = n, _lib.MCP_F32, 50_001, 0, 3, 20, _lib.MCP_DEVICE, 1 / 252
    w = np.full(n, 1 / n)
    st1, st2 = _lib.PathStats(), _lib.PathStats()
    for st in (st1, st2):
        st.n_alphas = 2
        st.alphas[0], st.alphas[1] = 0.95, 0.99
    arr = (C.c_void_p * 1)(h)
    assert L.mcp_paths_stats_multi(arr, 1, C.byref(pp), ptr(mu), ptr(sigma), ptr(w), C.byref(st1)) == 0, last_error(L, h)
    assert L.mcp_paths_stats(h, C.byref(pp), ptr(mu), ptr(sigma), ptr(w), None, C.byref(st2)) == 0
    assert list(st1.var)[:2] == list(st2.var)[:2] and list(st1.cvar)[:2] == list(st2.cvar)[:2]
    # error behaviour: NULL group, a group without a communicator
    assert L.mcp_portfolios_multi(None, 1, C.byref(p), ptr(mu), ptr(sigma), C.byref(ref_out), None) == _lib.MCP_ERR_INVALID
    two = (C.c_void_p * 2)(h, h)
    assert L.mcp_portfolios_multi(two, 2, C.byref(p), ptr(mu), ptr(sigma), C.byref(ref_out), None) == _lib.MCP_ERR_COMM
    assert "communicator" in last_error(L, h)


def test_two_communicators_merge_through_the_raw_abi(abi):
    """Two handles on two GPUs, mcp_comm_init_all, one mcp_portfolios_multi call: the picks and arrays of the one-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    L, h, _lib = abi
    n = 16
    mu, sigma = synthetic_inputs(n, seed=0)
    P = 400_003
    hs = [C.c_void_p(), C.c_void_p()]
    for d, hd in enumerate(hs):
        assert L.mcp_create(d, C.byref(hd)) == 0
    try:
        arr = (C.c_void_p * 2)(*hs)
        assert L.mcp_comm_init_all(arr, 2) == 0, last_error(L, hs[0])
        rank, world = C.c_int(), C.c_int()
        assert L.mcp_comm_info(hs[1], C.byref(rank), C.byref(world)) == 0 and (rank.value, world.value) == (1, 2)
        one, ws1, wt1, b1, _ = _multi_job(L, _lib, [h], 1, mu, sigma, P, True)
        two, ws2, wt2, b2, ms = _multi_job(L, _lib, hs, 2, mu, sigma, P, True)
        assert two.max_sharpe.index == one.max_sharpe.index and two.target_risk.index == one.target_risk.index
        assert two.max_sharpe.sharpe == one.max_sharpe.sharpe and np.array_equal(ws1, ws2) and np.array_equal(wt1, wt2)
        assert two.n_accepted == one.n_accepted == P and two.n_accepted_global == P and (ms > 0).all()
        for k in b1:
            assert np.array_equal(b1[k], b2[k]), k
        for hd in hs:
            assert L.mcp_comm_destroy(hd) == 0
    finally:
        for hd in hs:
            L.mcp_destroy(hd)
