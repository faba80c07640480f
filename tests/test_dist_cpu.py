"""CPU tier, world_size 2 over gloo: the host logic of the N>1 path -- index-range sharding,
the selection-record merge (all_gather + deterministic pick) and the distributed radix select
(histogram all-reduce between passes driving libmcp's host state machine)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from conftest import PKG_DIR, ROOT


def test_shard_range_partitions_exactly():
    from mcportfolio.dist import shard_range
    for total in (0, 1, 7, 10**10, 10**10 + 3):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks[:-1], blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_merge_records_first_occurrence():
    from mcportfolio.dist import merge_records, pack_record, unpack_record
    mk = lambda k, i: {"key": k, "global_index": i, "index": i, "ret": 0.1, "risk": 0.2, "sharpe": k, "weights": np.array([0.5, 0.5])}
    recs = [mk(1.0, 50), None, mk(2.0, 99), mk(2.0, 7), mk(float("nan"), 1)]
    assert merge_records(recs, True)["global_index"] == 7
    assert merge_records(recs, False)["global_index"] == 50
    assert merge_records([None, None], True) is None
    i, b = pack_record(recs[2], 2)
    back = unpack_record(i, b)
    assert back["global_index"] == 99 and back["key"] == 2.0 and np.array_equal(back["weights"], [0.5, 0.5])
    assert unpack_record(*pack_record(None, 2)) is None
    big = mk(1.0, 2**40 + 12345)                     # 64-bit indices survive the trip
    assert unpack_record(*pack_record(big, 2))["global_index"] == 2**40 + 12345


def test_flat_records_carry_int64_indices_bit_for_bit():
    """The per-step all_gather ships records as ONE float64 vector; the int64 global index is a bit pattern inside it
    (patterns that read as NaN / denormal / -0.0 as a double must come back unchanged)."""
    import torch
    from mcportfolio.dist import _pack_flat, _unpack_flat
    mk = lambda i: {"key": 1.5, "global_index": i, "index": i, "ret": 0.1, "risk": 0.2, "sharpe": 1.5, "weights": np.array([0.25, 0.75])}
    for idx in (0, 1, 2**40 + 12345, 2**53 + 1, 2**63 - 1, 0x7FF8000000000001, 0x7FF0000000000000, 0x0000000000000001):
        flat = _pack_flat(mk(idx), 2)
        assert flat.dtype == np.float64 and flat.shape == (8,)
        trip = torch.from_numpy(flat.copy()).clone().numpy()          # what a gather + copy back does: no arithmetic
        back = _unpack_flat(trip)
        assert back["global_index"] == idx and back["key"] == 1.5 and np.array_equal(back["weights"], [0.25, 0.75])
    assert _unpack_flat(_pack_flat(None, 2)) is None
    acc = np.array([123456789012], dtype=np.int64).view(np.float64)
    assert int(np.ascontiguousarray(acc).view(np.int64)[0]) == 123456789012


def test_merge_envelopes():
    from mcportfolio.dist import merge_envelopes
    a_r, a_i = np.array([1.0, -np.inf, 3.0, 2.0]), np.array([5, -1, 7, 9])
    b_r, b_i = np.array([1.0, 4.0, 2.5, -np.inf]), np.array([2, 11, 8, -1])
    best, idx = merge_envelopes([a_r, b_r], [a_i, b_i])
    assert list(idx) == [2, 11, 7, 9] and list(best) == [1.0, 4.0, 3.0, 2.0]
    best2, idx2 = merge_envelopes([b_r, a_r], [b_i, a_i])                 # order-independent
    assert list(idx2) == list(idx) and list(best2) == list(best)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    for p in (ROOT, PKG_DIR, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from mcportfolio import _lib
    from mcportfolio.dist import all_gather_records, merge_records, shard_range
    from test_abi_cpu import _hist_np, f32_keys
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # ---- selection merge: each rank contributes its local best; ties -> lowest global index
        local = {"key": 1.5, "global_index": 1000 - rank, "index": 0, "ret": 0.3, "risk": 0.2,
                 "sharpe": 1.5, "weights": np.full(4, 0.25) * (rank + 1)}
        merged = merge_records(all_gather_records(local, 4), True)
        # ---- distributed exact select: this rank holds one shard of a common vector
        x = np.random.default_rng(42).standard_normal(20_001).astype(np.float32)
        first, count = shard_range(len(x), rank, world)
        keys = f32_keys(x[first:first + count])
        L = _lib.lib()
        st = _lib.SelectState()
        ranks = np.array([0, 1000, 1001, len(x) - 1], dtype=np.uint64)
        assert L.mcp_select_init(C.byref(st), 32, ranks.ctypes.data, len(ranks)) == 0
        while L.mcp_select_pass_bits(C.byref(st)):
            bits = L.mcp_select_pass_bits(C.byref(st))
            hist = np.stack([_hist_np(keys, st.slot_prefix[s], st.bits_done, bits, 32) for s in range(st.n_slots)])
            t = torch.from_numpy(hist.astype(np.int64))
            dist.all_reduce(t)                                     # what NCCL does on the GPU box
            summed = np.ascontiguousarray(t.numpy().astype(np.uint64))
            assert L.mcp_select_advance(C.byref(st), summed.ctypes.data) == 0
        vals = [L.mcp_key_to_value(st.prefix[i], 0) for i in range(len(ranks))]
        q.put((rank, merged["global_index"], merged["weights"].tolist(), vals))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_selection_and_select_merge():
    import torch.multiprocessing as mp
    import mcportfolio
    mcportfolio.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.random.default_rng(42).standard_normal(20_001).astype(np.float32)
    xs = np.sort(x)
    want = [float(xs[i]) for i in (0, 1000, 1001, len(x) - 1)]
    for rank, gidx, w, vals in out:
        assert gidx == 999                       # equal keys: lowest global index (rank 1's) wins
        assert w == [0.5] * 4
        assert vals == want
