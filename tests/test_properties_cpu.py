"""CPU tier: property tests (hypothesis) on the host logic -- shard partitioning, the radix-select
state machine against sorting, key transforms, envelope merge -- and on the oracle itself."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from test_abi_cpu import _hist_np, f32_keys, run_select


@pytest.fixture(scope="module")
def mcp():
    import mcportfolio
    mcportfolio.build()
    return mcportfolio


@settings(max_examples=200, deadline=None)
@given(total=st.integers(0, 10**12), world=st.integers(1, 16))
def test_shard_range_is_a_partition(total, world):
    from mcportfolio.dist import shard_range
    blocks = [shard_range(total, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][0] + blocks[-1][1] == total
    assert all(a[0] + a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


@settings(max_examples=60, deadline=None)
@given(data=st.data())
def test_select_state_machine_equals_sorting(mcp, data):
    n = data.draw(st.integers(1, 400))
    vals = data.draw(st.lists(st.floats(allow_nan=False, allow_infinity=False, width=32), min_size=n, max_size=n))
    x = np.array(vals, dtype=np.float32)
    ranks = sorted(set(data.draw(st.lists(st.integers(0, n - 1), min_size=1, max_size=6))))
    parts = data.draw(st.integers(1, 3))
    cuts = sorted(data.draw(st.lists(st.integers(0, n), min_size=parts - 1, max_size=parts - 1)))
    shards = np.split(x, cuts)
    got = run_select(mcp, [f32_keys(s) for s in shards], ranks)
    xs = np.sort(x)
    want = [float(xs[r]) for r in ranks]
    # -0.0 and +0.0 are distinct keys (ordered -0 < +0) but equal values
    assert all(g == w for g, w in zip(got, want))


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(allow_nan=False, width=32), min_size=2, max_size=50))
def test_f32_key_transform_preserves_order(mcp, vals):
    x = np.array(vals, dtype=np.float32)
    k = f32_keys(x)
    order_v = np.argsort(x, kind="stable")
    assert np.all(np.diff(k[order_v].astype(np.int64)) >= 0) or np.any((x == 0) & np.signbit(x))
    L = mcp.lib()
    back = np.array([L.mcp_key_to_value(int(kk), 0) for kk in k], dtype=np.float32)
    assert np.array_equal(back.view(np.uint32), x.view(np.uint32))


@settings(max_examples=50, deadline=None)
@given(data=st.data())
def test_merge_envelopes_is_order_independent(data):
    from mcportfolio.dist import merge_envelopes
    K = data.draw(st.integers(1, 12))
    world = data.draw(st.integers(1, 4))
    rets, idxs = [], []
    used = set()
    for _ in range(world):
        r = np.array(data.draw(st.lists(st.sampled_from([-np.inf, 0.1, 0.2, 0.3]), min_size=K, max_size=K)))
        i = np.full(K, -1, dtype=np.int64)
        for b in range(K):
            if np.isfinite(r[b]):
                j = data.draw(st.integers(0, 10**6).filter(lambda v: v not in used))
                used.add(j)
                i[b] = j
        rets.append(r); idxs.append(i)
    a = merge_envelopes(rets, idxs)
    b = merge_envelopes(rets[::-1], idxs[::-1])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    stack = np.stack(rets)
    assert np.array_equal(a[0], stack.max(0))


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 500), st.sampled_from([0.5, 0.9, 0.95, 0.99]), st.integers(0, 2**31 - 1))
def test_oracle_percentile_restatement_matches_numpy(n, alpha, seed):
    from oracle import reference_np as ref
    x = np.random.default_rng(seed).standard_normal(n)
    q = ref.lower_tail_percent(alpha)
    assert ref.percentile_linear(x, q) == pytest.approx(np.percentile(x, q), rel=1e-14, abs=1e-300)
    v, c = ref.var(x, alpha), ref.cvar(x, alpha)
    assert c <= v + 1e-15 and v >= x.min() and v <= x.max()


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 40), st.integers(0, 2**31 - 1))
def test_oracle_metrics_properties(n, seed):
    """Scale invariances of the reference's formulas (app.py:708-711): w.mu is linear, risk is
    1-homogeneous, Sharpe with rf=0 is scale-free; risk^2 equals the explicit double sum."""
    from oracle import reference_np as ref
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    sigma = A @ A.T / n
    mu = rng.uniform(-0.2, 0.6, n)
    W = rng.dirichlet(np.ones(n), size=7)
    ret, risk, sharpe = ref.portfolio_metrics(W, mu, sigma, 0.0)
    ret2, risk2, sharpe2 = ref.portfolio_metrics(2.5 * W, mu, sigma, 0.0)
    assert np.allclose(ret2, 2.5 * ret) and np.allclose(risk2, 2.5 * risk) and np.allclose(sharpe2, sharpe)
    explicit = np.array([sum(w[i] * sigma[i, j] * w[j] for i in range(n) for j in range(n)) for w in W])
    assert np.allclose(risk ** 2, explicit, rtol=1e-10)
