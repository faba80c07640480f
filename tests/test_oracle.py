"""The oracle against the golden vectors generated from the reference's own lines
(oracle/make_golden.py) and the known-answer vectors of SURVEY.md section 4.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import paths_np, philox_np, reference_np as ref


# ---- SURVEY section 4 known-answer vectors (computed from app.py lines under legacy seeds) ----

def test_c1_mu_sigma_known_answer(c1):
    assert np.allclose(c1["mu"], [0.5616761212195296, 0.5740125994716275], rtol=1e-14)
    assert np.allclose(c1["sigma"], [[0.43009738895980193, 0.4498662775736327],
                                     [0.4498662775736327, 0.7434297067371608]], rtol=1e-14)
    mu, sigma = ref.estimate_mu_sigma(c1["returns_matrix"], 52)
    assert np.allclose(mu, c1["mu"], rtol=1e-12) and np.allclose(sigma, c1["sigma"], rtol=1e-12)
    assert np.all(c1["returns_matrix"][0] == 0.0)          # fillna(0) row is kept (app.py:666)


def test_c1_legacy_dirichlet_stream(c1):
    np.random.seed(42)
    W, valid = ref.draw_weights_reference(2, 10_000, [0, 0], [1, 1])
    assert valid.all() and np.array_equal(W, c1["weights"])
    assert np.allclose(W[0], [0.13487081200849088, 0.8651291879915091], rtol=1e-15)


def test_c1_metrics_and_picks(c1):
    for tag, rf, want in (("rf3", 3.0, 1451), ("rf003", 0.03, 4593)):
        out = ref.evaluate(c1["weights"], c1["mu"], c1["sigma"], rf, 0.30)
        assert np.allclose(out["risks"], c1["risks"], rtol=1e-13)
        assert np.allclose(out["returns"], c1["returns"], rtol=1e-13)
        assert np.allclose(out["sharpes"], c1[f"sharpes_{tag}"], rtol=1e-12)
        assert out["max_sharpe"]["index"] == want == int(c1[f"opt_sharpe_{tag}"])
        assert out["target_risk"]["index"] == 4593 == int(c1["opt_target30"])
    assert np.isclose(c1["sharpes_rf3"][1451], -2.813664661896989, rtol=1e-13)
    assert np.isclose(c1["risks"].min(), 0.6558193284015847) and np.isclose(c1["risks"].max(), 0.8622163471565484)


def test_c1_historical_var_cvar_methods(c1):
    v, c = ref.historical_var_cvar(c1["returns_matrix"], c1["weights"], 0.95)
    assert np.allclose(-v, c1["neg_var95"], rtol=1e-12, atol=1e-15)
    assert np.allclose(-c, c1["neg_cvar95"], rtol=1e-12, atol=1e-15)
    assert ref.select_method(-v, "VaR") == int(c1["opt_var"]) == 4593
    assert ref.select_method(-c, "CVaR") == int(c1["opt_cvar"]) == 4593
    assert np.isclose(v[4593], -0.13301122778135846) and np.isclose(c[4593], -0.1954488675070788)


def test_equal_weight_method(c1):
    for tag, rf in (("rf3", 3.0), ("rf003", 0.03)):
        ret, risk, sharpe = ref.portfolio_metrics(np.array([[0.5, 0.5]]), c1["mu"], c1["sigma"], rf)
        assert np.allclose([risk[0], ret[0], sharpe[0]], c1[f"ew_{tag}"], rtol=1e-13)
    assert ref.select_method(np.array([1.0]), "Equal Weight") == 0
    with pytest.raises(IndexError):
        ref.select_method(np.array([]), "Equal Weight")


def test_bounded_rejection_skip_semantics():
    g = load_golden("c1_bounded.npz")
    np.random.seed(int(g["seed"]))
    risk, ret, W, metrics = ref.inline_loop(g["mu"], g["sigma"], int(g["n_requested"]), 0.03,
                                            g["min_weights"], g["max_weights"])
    assert W.shape == g["weights"].shape and W.shape[0] < int(g["n_requested"])
    assert np.array_equal(W, g["weights"])
    assert np.allclose(risk, g["risks"], rtol=1e-13) and np.allclose(metrics, g["sharpes"], rtol=1e-12)
    assert int(np.argmax(metrics)) == int(g["opt_idx"])


def test_efficient_frontier_twin():
    g = load_golden("efficient_frontier.npz")
    for seed, kw, tag in ((7, dict(points=5), "seed7"), (11, dict(points=200), "seed11"),
                          (5, dict(points=50, min_weights=np.array([0.49, 0.0]),
                                   max_weights=np.array([0.5, 1.0])), "seed5_bounded")):
        np.random.seed(seed)
        res, W = ref.efficient_frontier(g["mu"], g["sigma"], **kw)
        assert np.array_equal(W, g[f"{tag}_weights"])
        assert np.allclose(res, g[f"{tag}_results"], rtol=1e-13)
    assert np.allclose(g["seed7_results"][0], [0.845491685764601, 0.7666897874447456, 0.6667317572369921,
                                               0.6606946177962404, 0.7664003944524161], rtol=1e-14)


def test_var_cvar_against_reference_outputs():
    with open(os.path.join(GOLDEN, "var_cvar.json")) as fh:
        cases = json.load(fh)["cases"]
    assert len(cases) > 40
    for c in cases:
        x = np.array(c["x"])
        assert ref.var(x, c["alpha"]) == pytest.approx(c["var"], rel=1e-15, abs=0)
        assert ref.cvar(x, c["alpha"]) == pytest.approx(c["cvar"], rel=1e-14, abs=0)
        assert ref.percentile_linear(x, ref.lower_tail_percent(c["alpha"])) == pytest.approx(c["var"], rel=1e-14, abs=1e-18)


def test_quantile_fraction_constants():
    assert ref.lower_tail_percent(0.95) == 5.000000000000004
    assert ref.lower_tail_percent(0.99) == 1.0000000000000009


def test_c2_policy_fixture(c2):
    assert c2["returns_matrix"].shape == (24, 14)
    out = ref.evaluate(c2["weights"], c2["mu"], c2["sigma"], 0.03, 0.30)
    assert np.allclose(out["risks"], c2["risks"], rtol=1e-12)
    assert out["max_sharpe"]["index"] == int(c2["opt_sharpe"])
    v, c = ref.historical_var_cvar(c2["returns_matrix"], c2["weights"], 0.95)
    assert np.allclose(-v, c2["neg_var95"], rtol=1e-11, atol=1e-15)
    assert np.allclose(-c, c2["neg_cvar95"], rtol=1e-11, atol=1e-15)
    assert ref.select_method(-v, "VaR") == int(c2["opt_var"])
    assert ref.select_method(-c, "CVaR") == int(c2["opt_cvar"])


@pytest.mark.needs_reference
def test_restatement_vs_reference_lines_live():
    """Runs the reference's own functions (exec of app.py:231-284) beside the restatement."""
    from oracle import ref_loader
    fns = ref_loader.load_reference_functions()
    rng = np.random.default_rng(1)
    for n in (5, 64, 365):
        x = rng.standard_normal(n) * 0.1
        for a in (0.9, 0.95, 0.99):
            assert fns["var"](x, a) == ref.var(x, a)
            assert fns["cvar"](x, a) == ref.cvar(x, a)
    mu = rng.uniform(0.05, 0.5, 6)
    A = rng.standard_normal((6, 6))
    sigma = A @ A.T / 6
    np.random.seed(3)
    r1, w1 = fns["efficient_frontier"](mu, sigma, points=40)
    np.random.seed(3)
    r2, w2 = ref.efficient_frontier(mu, sigma, points=40)
    assert np.array_equal(w1, w2) and np.allclose(r1, r2, rtol=1e-13)


# ---- Philox4x32-10: Random123 known-answer vectors -------------------------------------------

def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox_np.philox4x32(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_philox_dirichlet_is_flat():
    """Normalised base-2 exponentials = flat Dirichlet: mean 1/N, var (N-1)/(N^2 (N+1))."""
    N, P = 16, 200_000
    W, valid = philox_np.dirichlet_weights(0, P, N, seed=0)
    assert valid.all() and np.allclose(W.sum(1), 1.0)
    assert np.allclose(W.mean(0), 1 / N, atol=4 * np.sqrt((N - 1) / (N * N * (N + 1)) / P))
    assert np.allclose(W.var(0), (N - 1) / (N * N * (N + 1)), rtol=0.03)
    # sharding invariance: a sub-range reproduces the same rows
    W2, _ = philox_np.dirichlet_weights(1000, 50, N, seed=0)
    assert np.array_equal(W2, W[1000:1050])


def test_philox_normals_moments():
    Z = philox_np.normals(0, 20_000, 3, 16, seed=1)
    assert abs(Z.mean()) < 4 / np.sqrt(Z.size) and abs(Z.var() - 1) < 0.01
    c = np.corrcoef(Z[:, 0, :], rowvar=False)
    assert np.abs(c - np.eye(16)).max() < 0.04


# ---- path simulator / envelope (north-star spec; parity unpinned) ----------------------------

def test_paths_oracle_moments(synth16):
    mu, sigma = synth16
    rng = np.random.default_rng(0)
    Z = rng.standard_normal((4000, 252, 16))
    w = np.full(16, 1 / 16)
    x = paths_np.terminal_returns(mu, sigma, w, Z)
    # E[prod(1 + mu dt + ...)] = (1 + mu dt)^252 per asset (independent steps)
    want = w @ ((1 + mu / 252) ** 252) - 1
    assert abs(x.mean() - want) < 4 * x.std() / np.sqrt(len(x))
    st = paths_np.risk_stats(x)
    assert st[0.99][0] < st[0.95][0] < 0 or st[0.95][0] < x.mean()
    assert st[0.95][1] <= st[0.95][0] and st[0.99][1] <= st[0.99][0]


def test_envelope_oracle():
    risks = np.array([0.1, 0.15, 0.2, 0.2, 0.3, 0.05, 0.35])
    rets = np.array([1.0, 2.0, 3.0, 3.0, 0.5, 9.0, 9.0])
    best, idx = paths_np.envelope(risks, rets, 2, 0.1, 0.3)
    assert list(idx) == [1, 2] and list(best) == [2.0, 3.0]


def test_asset_stats_against_reference_functions(c1, c2):
    with open(os.path.join(GOLDEN, "asset_stats.json")) as fh:
        g = json.load(fh)
    for tag, R in (("c1", c1["returns_matrix"]), ("c2", c2["returns_matrix"])):
        spec = g[tag]
        for j, want in enumerate(spec["assets"]):
            got = ref.asset_stats(R[1:, j], spec["risk_free"], spec["ann_factor"])
            for k, v in want.items():
                assert got[k] == pytest.approx(v, rel=1e-12, abs=1e-15), (tag, j, k)


def test_fp32_uniform_fields_are_24_bit_slices_of_the_stream():
    """float32 streams: field i = bits [24 i, 24 i + 23) of the concatenated Philox blocks (word 0 of block 0 lowest),
    the spec `philox_fields` in mcp_device.cuh implements with funnel shifts; float64 keeps one word per uniform."""
    idx = np.array([5, 2**40 + 7, 2**63 + 11], dtype=np.uint64)
    sub = np.array([0, 3, 99], dtype=np.uint64)
    for nf in (1, 3, 4, 5, 16, 24, 33, 256):
        f = philox_np._fields24(idx, sub, nf, philox_np.STREAM_WEIGHTS, 99)
        n_words = 4 * ((3 * ((nf + 3) // 4) + 3) // 4)
        w = philox_np._raw_outputs(idx, sub, n_words, philox_np.STREAM_WEIGHTS, 99)
        assert f.shape == (3, nf) and f.dtype == np.uint32 and f.max() < 2**23
        for r in range(3):
            big = 0
            for j, x in enumerate(w[r]):
                big |= int(x) << (32 * j)
            assert [int(v) for v in f[r]] == [(big >> (24 * i)) & 0x7FFFFF for i in range(nf)]
    # 16 uniforms take 3 blocks, not 4; the exponentials are -log2(1 - field 2^-23)
    e32 = philox_np.exponentials(idx, 0, 16, 7, "float32")
    f = philox_np._fields24(idx, np.zeros(3, dtype=np.uint64), 16, philox_np.STREAM_WEIGHTS, 7)
    assert np.array_equal(e32, -np.log2(1.0 - f.astype(np.float64) * 2.0 ** -23))
    e64 = philox_np.exponentials(idx, 0, 16, 7, "float64")
    x = philox_np._raw_outputs(idx, np.zeros(3, dtype=np.uint64), 16, philox_np.STREAM_WEIGHTS, 7)
    assert np.array_equal(e64, -np.log2(1.0 - x.astype(np.float64) * 2.0 ** -32))
