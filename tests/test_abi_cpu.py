"""CPU tier: the C-ABI library loads without a GPU and exports every symbol include/mcp.h
declares; host-side argument validation; the radix-select state machine (pure host code)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


def test_library_exports_every_declared_symbol(mcp):
    header = open(os.path.join(ROOT, "include", "mcp.h")).read()
    declared = set(re.findall(r"\b(mcp_[a-z0-9_]+)\s*\(", header))
    declared -= {"mcp_allreduce_fn"}
    from mcportfolio import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = mcp.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.mcp_abi_version() == _lib.MCP_ABI_VERSION == 2


def test_struct_layouts_match_header(mcp):
    """sizeof checks against the C layout rules (x86-64 SysV) for the POD structs."""
    from mcportfolio import _lib
    assert C.sizeof(_lib.PortfolioParams) == 8 + 24 + 16 + 16 + 8 + 8 + 16 + 8 + 16
    assert C.sizeof(_lib.Selection) == 48
    assert C.sizeof(_lib.PortfolioOut) == 7 * 8 + 8 + 16 + 2 * 48 + 8 + 8 + 8
    assert C.sizeof(_lib.PathParams) == 8 + 24 + 8 + 8 + 8 + 8
    assert C.sizeof(_lib.PathStats) == 8 + 8 + 3 * 8 * _lib.MCP_MAX_ALPHAS + 16
    assert C.sizeof(_lib.HistParams) == 16 + 16 + 8 + 8 + 8
    assert C.sizeof(_lib.SelectState) == 16 + 16 * 8 * 2 + 16 * 4 + 16 * 8


def test_no_gpu_fails_loudly(mcp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mcp.McpError, match="no CUDA device|CPU fallback"):
        mcp.Engine(0)


def test_argument_validation_happens_on_host(mcp):
    mu = np.array([0.1, 0.2])
    sig = np.eye(2)
    with pytest.raises(ValueError, match="cov_matrix must have shape"):
        mcp.simulate_portfolios(mu, np.eye(3), 10)
    with pytest.raises(ValueError, match="dtype"):
        mcp.simulate_portfolios(mu, sig, 10, dtype="float16")
    with pytest.raises(ValueError, match="weights must have shape"):
        mcp.simulate_portfolios(mu, sig, 10, weights=np.ones((10, 3)))
    with pytest.raises(ValueError, match="rows but n_portfolios"):
        mcp.simulate_portfolios(mu, sig, 10, weights=np.ones((4, 2)))
    with pytest.raises(ValueError, match="non-finite"):
        mcp.simulate_portfolios(np.array([np.nan, 0.1]), sig, 10)
    with pytest.raises(ValueError, match="n_portfolios"):
        mcp.simulate_portfolios(mu, sig, -1)
    with pytest.raises(ValueError, match="weights must have shape"):
        mcp.simulate_paths(mu, sig, np.ones(3), 10)


def _hist_np(keys, prefix, bits_done, bits, key_bits):
    """numpy stand-in for one device histogram pass (what select_hist_kernel computes)."""
    shift = key_bits - bits_done - bits
    hi = keys >> np.uint64(shift + bits) if shift + bits < 64 else np.zeros_like(keys)
    sel = keys[hi == np.uint64(prefix)] if bits_done else keys
    digit = (sel >> np.uint64(shift)) & np.uint64((1 << bits) - 1)
    return np.bincount(digit.astype(np.int64), minlength=1 << bits).astype(np.uint64)


def f32_keys(x):
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    neg = (b >> np.uint64(31)).astype(bool)
    return np.where(neg, b ^ np.uint64(0xFFFFFFFF), b ^ np.uint64(0x80000000))


def run_select(mcp, keys_per_rank, ranks, key_bits=32):
    """Drive the real host state machine with numpy histograms summed over 'ranks'."""
    from mcportfolio import _lib
    L = mcp.lib()
    st = _lib.SelectState()
    r = np.asarray(ranks, dtype=np.uint64)
    assert L.mcp_select_init(C.byref(st), key_bits, r.ctypes.data, len(r)) == 0
    while True:
        bits = L.mcp_select_pass_bits(C.byref(st))
        if bits == 0:
            break
        hist = np.zeros((st.n_slots, 1 << bits), dtype=np.uint64)
        for keys in keys_per_rank:                      # the all-reduce, spelled out
            for s in range(st.n_slots):
                hist[s] += _hist_np(keys, st.slot_prefix[s], st.bits_done, bits, key_bits)
        assert L.mcp_select_advance(C.byref(st), hist.ctypes.data) == 0
    return [L.mcp_key_to_value(st.prefix[t], 0 if key_bits == 32 else 1) for t in range(len(r))]


def test_select_state_machine_finds_exact_order_statistics(mcp):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.standard_normal(5000), np.round(rng.standard_normal(3000), 1), [-0.0, 0.0, 1e-30, -1e-30]])
    x = x.astype(np.float32)
    xs = np.sort(x)
    ranks = [0, 1, 17, 4001, len(x) - 2, len(x) - 1]
    got = run_select(mcp, [f32_keys(x)], ranks)
    assert got == [float(xs[r]) for r in ranks]
    # sharded over 3 uneven "ranks": same answer (histograms add)
    parts = np.split(x, [100, 5000])
    got2 = run_select(mcp, [f32_keys(p) for p in parts], ranks)
    assert got2 == got


def test_select_rank_out_of_population_is_an_error(mcp):
    from mcportfolio import _lib
    L = mcp.lib()
    st = _lib.SelectState()
    r = np.array([10], dtype=np.uint64)
    assert L.mcp_select_init(C.byref(st), 32, r.ctypes.data, 1) == 0
    hist = np.zeros((1, 2048), dtype=np.uint64)
    hist[0, 3] = 5                                     # only 5 elements, rank 10 requested
    assert L.mcp_select_advance(C.byref(st), hist.ctypes.data) != 0
    assert L.mcp_select_init(C.byref(st), 16, r.ctypes.data, 1) != 0


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` prints ONE JSON line with the contract's keys (CPU only)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
