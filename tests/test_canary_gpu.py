"""GPU tier: out-of-bounds guard for the device-space write-back of every sweep kernel family (compute-sanitizer is not
available on this pool): outputs are interior slices of one arena whose gaps hold a sentinel; after the call every gap
must be intact and every output element must have been written."""
import ctypes as C

import numpy as np
import pytest

from conftest import synthetic_inputs

pytestmark = pytest.mark.gpu
SENT = -7.25e30          # exactly representable in FP32 / FP64, never a metric or a weight


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("n", [2, 16, 20, 33, 100, 255, 256, 300])
@pytest.mark.parametrize("P", [1, 127, 129, 1001])
@pytest.mark.parametrize("supplied", [False, True])
def test_device_outputs_stay_inside_their_buffers(n, P, dtype, supplied):
    import torch
    import mcportfolio as mcp
    from mcportfolio import api
    from mcportfolio._lib import MCP_DEVICE, MCP_F32, MCP_F64, PortfolioOut, PortfolioParams, check, lib
    mcp.build()
    eng = api.get_engine(None)
    tdt = torch.float32 if dtype == "float32" else torch.float64
    es = 4 if dtype == "float32" else 8
    mu, sigma = synthetic_inputs(n, seed=3)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    gap = 64                                                   # elements of sentinel around every output
    sizes = {"weights": P * n, "returns": P, "risks": P, "sharpes": P}
    total = sum(sizes.values()) + gap * (len(sizes) + 1)
    arena = torch.full((total,), SENT, dtype=tdt, device="cuda")
    acc_arena = torch.full((P + 2 * gap,), 77, dtype=torch.uint8, device="cuda")
    offs, o = {}, gap
    for k, sz in sizes.items():
        offs[k] = o
        o += sz + gap
    w_in = None
    if supplied:
        w_in = torch.from_numpy(np.random.RandomState(P).dirichlet(np.ones(n), size=P)).to("cuda", tdt).contiguous()
    p = PortfolioParams()
    p.n_assets, p.dtype = n, MCP_F32 if dtype == "float32" else MCP_F64
    p.n_portfolios, p.first_index, p.seed = P, 2**33 + 5, 9
    p.risk_free, p.risk_target, p.max_tries, p.keep_last, p.space = 0.03, 0.3, 100, 0, MCP_DEVICE
    p.weights_in = w_in.data_ptr() if supplied else None
    res = PortfolioOut()
    base = arena.data_ptr()
    res.weights, res.returns = base + offs["weights"] * es, base + offs["returns"] * es
    res.risks, res.sharpes = base + offs["risks"] * es, base + offs["sharpes"] * es
    res.accepted = acc_arena.data_ptr() + gap
    ws, wt = np.empty(n), np.empty(n)
    res.max_sharpe.weights, res.target_risk.weights = ws.ctypes.data, wt.ctypes.data
    eng.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
    check(eng.handle, lib().mcp_portfolios(eng.handle, C.byref(p), mu.ctypes.data, sigma.ctypes.data, C.byref(res)))
    torch.cuda.synchronize()
    host = arena.cpu().numpy()
    inside = np.zeros(total, dtype=bool)
    for k, sz in sizes.items():
        inside[offs[k]:offs[k] + sz] = True
    assert np.all(host[~inside] == SENT), f"write outside the output buffers at {np.nonzero((host != SENT) & ~inside)[0][:8]}"
    assert not np.any(host[inside] == SENT), "an output element was never written"
    a = acc_arena.cpu().numpy()
    assert np.all(a[:gap] == 77) and np.all(a[gap + P:] == 77) and np.all(a[gap:gap + P] == 1)
    W = host[offs["weights"]:offs["weights"] + P * n].reshape(P, n)
    assert np.allclose(W.sum(1), 1.0, atol=1e-5) and int(res.n_accepted) == P


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("T,P", [(33, 5), (100, 1003), (365, 1), (365, 4097), (500, 130)])
def test_historical_outputs_stay_inside_their_buffers(T, P, dtype):
    """hist_var_fast handles portfolios in groups of four: the partial last group must not write past P."""
    import torch
    import mcportfolio as mcp
    from mcportfolio import api, _lib
    from mcportfolio._lib import MCP_DEVICE, MCP_F32, MCP_F64, check, lib
    mcp.build()
    eng = api.get_engine(None)
    tdt = torch.float32 if dtype == "float32" else torch.float64
    es = 4 if dtype == "float32" else 8
    n, gap = 16, 64
    rng = np.random.default_rng(T + P)
    R = np.ascontiguousarray(rng.standard_normal((T, n)) * 0.03)
    W = torch.from_numpy(rng.dirichlet(np.ones(n), size=P)).to("cuda", tdt).contiguous()
    arena = torch.full((2 * P + 3 * gap,), SENT, dtype=tdt, device="cuda")
    p = _lib.HistParams()
    p.n_assets, p.n_periods, p.dtype, p.space = n, T, MCP_F32 if dtype == "float32" else MCP_F64, MCP_DEVICE
    p.n_portfolios, p.first_index, p.alpha, p.weights_in = P, 0, 0.95, W.data_ptr()
    out = _lib.HistOut()
    out.var, out.cvar = arena.data_ptr() + gap * es, arena.data_ptr() + (2 * gap + P) * es
    eng.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
    check(eng.handle, lib().mcp_historical_var(eng.handle, C.byref(p), R.ctypes.data, C.byref(out)))
    torch.cuda.synchronize()
    host = arena.cpu().numpy()
    inside = np.zeros(host.size, dtype=bool)
    inside[gap:gap + P] = True
    inside[2 * gap + P:2 * gap + 2 * P] = True
    assert np.all(host[~inside] == SENT) and not np.any(host[inside] == SENT)
    assert np.all(host[2 * gap + P:2 * gap + 2 * P] <= host[gap:gap + P] + 1e-12)        # CVaR <= VaR


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("n,M", [(16, 1), (16, 511), (16, 513), (5, 1000), (32, 257)])
def test_path_terminals_stay_inside_their_buffer(n, M, dtype):
    """path_kernel_packed handles two paths per thread: an odd / ragged path count must not write past M."""
    import torch
    import mcportfolio as mcp
    from mcportfolio import api
    from mcportfolio._lib import MCP_DEVICE, MCP_F32, MCP_F64, PathParams, check, lib
    mcp.build()
    eng = api.get_engine(None)
    tdt = torch.float32 if dtype == "float32" else torch.float64
    es = 4 if dtype == "float32" else 8
    gap = 64
    mu, sigma = synthetic_inputs(n, seed=1)
    mu, sigma = np.ascontiguousarray(mu, dtype=np.float64), np.ascontiguousarray(sigma, dtype=np.float64)
    w = np.full(n, 1.0 / n)
    arena = torch.full((M + 2 * gap,), SENT, dtype=tdt, device="cuda")
    p = PathParams()
    p.n_assets, p.dtype, p.n_paths, p.first_index = n, MCP_F32 if dtype == "float32" else MCP_F64, M, 2**35
    p.seed, p.n_steps, p.space, p.dt, p.normals_in = 4, 7, MCP_DEVICE, 1.0 / 252, None
    ms = C.c_double()
    eng.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
    check(eng.handle, lib().mcp_paths(eng.handle, C.byref(p), mu.ctypes.data, sigma.ctypes.data, w.ctypes.data,
                                      arena.data_ptr() + gap * es, C.byref(ms)))
    torch.cuda.synchronize()
    host = arena.cpu().numpy()
    assert np.all(host[:gap] == SENT) and np.all(host[gap + M:] == SENT) and not np.any(host[gap:gap + M] == SENT)
    assert np.all(np.abs(host[gap:gap + M]) < 1.0)                                      # 7 daily steps: returns of a few percent
