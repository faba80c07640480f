"""Multi-GPU sharding of the path: one process per GPU, ``torch.distributed`` for the plumbing.

Portfolios and paths are independent units: rank r owns the contiguous global index block
``shard_range(total, r, world)`` and, because the Philox counter is the GLOBAL index, the
union over ranks is the same set of portfolios / paths for any world size.  No data-path
collective exists; only results cross NVLink (NCCL), all latency-bound:

* selection:  all_gather of one fixed-size record per rank and criterion, then the same
  deterministic merge on every rank (best key, lowest global index on ties = numpy's
  first-occurrence rule, app.py:672);
* VaR / CVaR: per radix pass an all-reduce(sum) of the <= 16 x 2048 histogram counts, then an
  all-reduce of the FP64 tail sums (``make_allreduce`` is the callback libmcp calls).

Per-portfolio arrays are never moved between GPUs.

Over NCCL the merges run INSIDE libmcp (``init_comm``: the library's own communicator, created from an id that rank 0
broadcasts through torch.distributed; entry points are then called with ``comm_merge``): the collectives are issued on
the handle's stream between the library's kernels and a sharded call costs one host wait.  The Python merges below
(``merge_records``, ``all_gather_flat`` ...) remain for the gloo backend -- the CPU tests of the host logic -- and as the
specification of what the library's merge does.
"""
from __future__ import annotations

import threading

import numpy as np

from . import api


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of rank `rank`: (first_index, count); blocks differ by at most 1."""
    total, rank, world = int(total), int(rank), int(world)
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


# ---- selection merge -----------------------------------------------------------------------

_REC_HEAD = 6     # valid, key, ret, risk, sharpe, (index travels separately as int64)


def pack_record(rec: dict | None, n_assets: int):
    """(int64 index, float64[_REC_HEAD - 1 + N]) -- fixed size so all_gather needs no negotiation."""
    body = np.full(_REC_HEAD - 1 + n_assets, np.nan)
    if rec is None:
        body[0] = 0.0
        return -1, body
    body[0] = 1.0
    body[1:5] = rec["key"], rec["ret"], rec["risk"], rec["sharpe"]
    body[5:] = rec["weights"]
    return int(rec["global_index"]), body


def unpack_record(index: int, body: np.ndarray) -> dict | None:
    if body[0] != 1.0:
        return None
    return {"index": int(index), "global_index": int(index), "key": float(body[1]), "ret": float(body[2]),
            "risk": float(body[3]), "sharpe": float(body[4]), "weights": np.array(body[5:], dtype=np.float64)}


def merge_records(records, larger_is_better: bool) -> dict | None:
    """Best key wins; ties go to the lowest global index (first occurrence, app.py:672)."""
    best = None
    for r in records:
        if r is None or np.isnan(r["key"]):
            continue
        if best is None:
            best = r
            continue
        better = r["key"] > best["key"] if larger_is_better else r["key"] < best["key"]
        if better or (r["key"] == best["key"] and r["global_index"] < best["global_index"]):
            best = r
    return best


def _pack_flat(rec: dict | None, n_assets: int) -> np.ndarray:
    """One record as float64[_REC_HEAD + N]: the int64 global index rides along bit for bit."""
    idx, body = pack_record(rec, n_assets)
    return np.concatenate([np.array([idx], dtype=np.int64).view(np.float64), body])


def _unpack_flat(flat: np.ndarray) -> dict | None:
    return unpack_record(int(flat[:1].view(np.int64)[0]), flat[1:])


def all_gather_flat(flat: np.ndarray, device=None, group=None) -> np.ndarray:
    """Every rank's float64 vector on every rank, shape (world, len): ONE all_gather and one copy back
    (each collective and each tiny D2H costs tens of microseconds: they are what a 20 ms step can lose)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cpu") if device is None else device
    t = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float64)).to(dev)
    out = torch.empty((world, t.numel()), dtype=torch.float64, device=dev)
    dist.all_gather(list(out.unbind(0)), t, group=group)
    return out.cpu().numpy()


def all_gather_records(rec: dict | None, n_assets: int, device=None, group=None):
    """Every rank's record, on every rank (one all_gather of (7 + N) doubles)."""
    return [_unpack_flat(row) for row in all_gather_flat(_pack_flat(rec, n_assets), device, group)]


def _with_global(rec):
    return None if rec is None else dict(rec, index=rec["global_index"])


_COMM_GROUPS = {}


def init_comm(group=None, device=None):
    """Give this rank's engine a libmcp communicator spanning `group` (idempotent).  Rank 0 creates the NCCL id, the 128
    bytes travel through torch.distributed, every rank calls mcp_comm_init.  Returns the engine."""
    import torch
    import torch.distributed as dist
    eng = api.get_engine(device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    key = (eng.device, id(group) if group is not None else 0, world)
    if _COMM_GROUPS.get(key) is eng and eng.comm_info() == (rank, world):
        return eng
    if eng.comm_info()[1]:
        eng.comm_destroy()
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group,
                               device=torch.device("cuda", eng.device) if dist.get_backend(group) == "nccl" else None)
    eng.comm_init(box[0], rank, world)
    _COMM_GROUPS[key] = eng
    return eng


def _use_comm(group, kw) -> bool:
    import torch.distributed as dist
    return dist.get_backend(group) == "nccl"


def simulate_portfolios_sharded(mean_returns, cov_matrix, n_portfolios, *, group=None, **kw):
    """`simulate_portfolios` over the whole job: this rank evaluates its block, selections are
    merged across ranks.  Arrays (if requested) stay local to the rank that produced them."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = shard_range(n_portfolios, rank, world)
    n = len(np.asarray(mean_returns))
    if _use_comm(group, kw):
        # the merge happens inside libmcp (one all-gather on the handle's stream, one host wait)
        init_comm(group, kw.get("device"))
        r = api.simulate_portfolios(mean_returns, cov_matrix, count, first_index=first, comm_merge=True, **kw)
        r.extra["shard"] = (first, count)
        return r
    r = api.simulate_portfolios(mean_returns, cov_matrix, count, first_index=first, **kw)
    dev = torch.device("cuda", api.get_engine(kw.get("device")).device) if dist.get_backend(group) == "nccl" else None
    L = _REC_HEAD + n
    flat = np.concatenate([_pack_flat(_with_global(r.max_sharpe), n), _pack_flat(_with_global(r.target_risk), n),
                           np.array([r.n_accepted], dtype=np.int64).view(np.float64)])
    G = all_gather_flat(flat, dev, group)
    r.max_sharpe = merge_records([_unpack_flat(row[:L]) for row in G], True)
    r.target_risk = merge_records([_unpack_flat(row[L:2 * L]) for row in G], False)
    r.extra["n_accepted_global"] = int(np.ascontiguousarray(G[:, 2 * L]).view(np.int64).sum())
    r.extra["shard"] = (first, count)
    return r


def frontier_envelope_sharded(mean_returns, cov_matrix, n_portfolios, n_bins=512, *, risk_range=None, group=None, **kw):
    """`frontier_envelope` over the whole job (C5).  One sweep of this rank's block with its (risk, return)
    pairs kept in HBM, the attained risk range all-reduced (min / max), then a bandwidth-bound binning pass
    (two sweeps -- range, then bins -- when the pairs do not fit or the backend is gloo); the per-rank bins
    (n_bins x (return, global index)) are all-gathered and merged (larger return, then lower
    index), so every rank ends with the envelope of the WHOLE job."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = shard_range(n_portfolios, rank, world)
    n = len(np.asarray(mean_returns))
    eng = api.get_engine(kw.get("device"))
    nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", eng.device) if nccl else torch.device("cpu")
    kw = dict(kw, return_arrays=False)
    if nccl:
        # picks and the risk range are merged inside libmcp; the bins through its host-buffer all-gather
        init_comm(group, kw.get("device"))
        single = risk_range is None and kw.get("weights") is None and api._metrics_fit(count, kw)
        if risk_range is None:                  # every rank must take the same route: all or none keep their metrics
            single = bool(eng.allreduce(np.array([0.0 if single else 1.0]), "max")[0] == 0.0)
        if single:
            r = api.simulate_portfolios(mean_returns, cov_matrix, count, first_index=first, comm_merge=True,
                                        **dict(kw, return_arrays="device-metrics"))
            if r.extra["n_accepted_global"] == 0:
                raise ValueError("no portfolio satisfied the bounds; the envelope is empty")
            risk_range = api._widen(r.risk_range)
            env = api.envelope_from_arrays(r.risks, r.returns, n_bins, risk_range, first_index=first, device=kw.get("device"))
            r.risks = r.returns = None
        else:
            if risk_range is None:
                probe = api.simulate_portfolios(mean_returns, cov_matrix, count, first_index=first, comm_merge=True, **kw)
                if probe.extra["n_accepted_global"] == 0:
                    raise ValueError("no portfolio satisfied the bounds; the envelope is empty")
                risk_range = api._widen(probe.risk_range)
            r = api.simulate_portfolios(mean_returns, cov_matrix, count, first_index=first, n_bins=n_bins, risk_range=risk_range,
                                        comm_merge=True, **kw)
            r.extra["risk_range_global"] = risk_range
            return r                             # n_bins > 0 with comm_merge: the library merged the bins as well
        K = int(n_bins)
        flat = np.concatenate([np.asarray(env["best_return"], dtype=np.float64), np.asarray(env["best_index"], dtype=np.int64).view(np.float64)])
        G = eng.allgather(flat)
        env["best_return"], env["best_index"] = merge_envelopes([row[:K] for row in G],
                                                                [np.ascontiguousarray(row[K:2 * K]).view(np.int64) for row in G])
        r.extra["envelope"] = env
        r.extra["risk_range_global"] = risk_range
        return r

    def global_range(r):
        lo, hi = r.risk_range if r.n_accepted else (float("inf"), float("-inf"))
        t = torch.tensor([-lo, hi], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return api._widen((-float(t[0]), float(t[1])))

    single = risk_range is None and nccl and kw.get("weights") is None and api._metrics_fit(count, kw)
    if risk_range is None:                      # every rank must take the same route: all or none keep their metrics
        flag = torch.tensor([1 if single else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        single = bool(flag.item())
    if single:
        # ONE sweep per rank with (risk, return) kept in HBM; the range is all-reduced, then each rank bins its arrays
        r = api.simulate_portfolios(mean_returns, cov_matrix, count, first_index=first, **dict(kw, return_arrays="device-metrics"))
        risk_range = global_range(r)
        r.extra["envelope"] = api.envelope_from_arrays(r.risks, r.returns, n_bins, risk_range, first_index=first, device=kw.get("device"))
        r.risks = r.returns = None
    else:
        if risk_range is None:
            probe = api.simulate_portfolios(mean_returns, cov_matrix, count, first_index=first, **kw)
            risk_range = global_range(probe)
        r = api.simulate_portfolios(mean_returns, cov_matrix, count, first_index=first, n_bins=n_bins,
                                    risk_range=risk_range, **kw)
    env = r.extra["envelope"]
    L = _REC_HEAD + n
    flat = np.concatenate([np.asarray(env["best_return"], dtype=np.float64), np.asarray(env["best_index"], dtype=np.int64).view(np.float64),
                           _pack_flat(_with_global(r.max_sharpe), n), _pack_flat(_with_global(r.target_risk), n)])
    G = all_gather_flat(flat, dev if nccl else None, group)
    K = int(n_bins)
    env["best_return"], env["best_index"] = merge_envelopes([row[:K] for row in G],
                                                            [np.ascontiguousarray(row[K:2 * K]).view(np.int64) for row in G])
    r.max_sharpe = merge_records([_unpack_flat(row[2 * K:2 * K + L]) for row in G], True)
    r.target_risk = merge_records([_unpack_flat(row[2 * K + L:2 * K + 2 * L]) for row in G], False)
    r.extra["risk_range_global"] = risk_range
    return r


def merge_envelopes(returns_list, index_list):
    """Per bin: the larger return wins, ties go to the lower global index; -1 marks an empty bin."""
    best = np.array(returns_list[0], dtype=np.float64, copy=True)
    idx = np.array(index_list[0], dtype=np.int64, copy=True)
    for r, i in zip(returns_list[1:], index_list[1:]):
        r = np.asarray(r, dtype=np.float64)
        i = np.asarray(i, dtype=np.int64)
        take = (i >= 0) & ((idx < 0) | (r > best) | ((r == best) & (i < idx)))
        best[take], idx[take] = r[take], i[take]
    return best, idx


# ---- VaR / CVaR merge ------------------------------------------------------------------------

class _DeviceBuffer:
    """Zero-copy view of a raw device pointer for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


def wrap_device_buffer(ptr: int, count: int, kind: int, device_index: int):
    import torch
    return torch.as_tensor(_DeviceBuffer(ptr, count, "<i8" if kind == 0 else "<f8"),
                           device=torch.device("cuda", device_index))


def make_allreduce(device_index: int, group=None, stream_ordered: bool = False):
    """The callback libmcp's mcp_quantiles calls between radix passes: NCCL sum, in place.

    stream_ordered=True: the all-reduce is only ENQUEUED (torch's NCCL work is ordered after, and the current
    stream then waits on, everything on the current stream -- which is the handle's stream), no host wait: use
    with Engine.set_allreduce_stream_ordered(True)."""
    import torch
    import torch.distributed as dist

    def allreduce(ptr, count, kind):
        t = wrap_device_buffer(ptr, count, kind, device_index)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if not stream_ordered:
            torch.cuda.current_stream(device_index).synchronize()

    return allreduce


def simulate_paths_sharded(mean_returns, cov_matrix, weights, n_paths, n_steps=252, *, group=None, **kw):
    """`simulate_paths` over the whole job: local paths, globally exact VaR / CVaR.  Over NCCL the radix passes,
    their histogram all-reduces and the digit selection are all stream-ordered (no host round trip per pass)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = shard_range(n_paths, rank, world)
    eng = api.get_engine(kw.get("device"))
    if dist.get_backend(group) == "nccl" and not kw.pop("legacy_callback", False):
        # libmcp's own communicator: path kernel (+ first histogram), select passes with their all-reduces, interpolation,
        # tail sums and their all-reduce all on the handle's stream; ONE host wait and one copy back
        init_comm(group, kw.get("device"))
        return api.simulate_paths(mean_returns, cov_matrix, weights, count, n_steps, first_index=first, comm_merge=True,
                                  n_total=n_paths, **kw)
    if not (world > 1 and dist.get_backend(group) == "nccl"):
        return api.simulate_paths(mean_returns, cov_matrix, weights, count, n_steps, first_index=first,
                                  allreduce=make_allreduce(eng.device, group) if world > 1 else None,
                                  n_total=n_paths, **kw)
    # stream-ordered route: libmcp and NCCL must share ONE real stream.  torch's default stream has handle 0, which
    # mcp_set_stream reads as "use the handle's own stream", so run the whole call on a side stream of ours.
    import torch
    side = _side_stream(eng.device)
    cur = torch.cuda.current_stream(eng.device)
    side.wait_stream(cur)
    eng.set_allreduce_stream_ordered(True)
    try:
        with torch.cuda.stream(side):
            out = api.simulate_paths(mean_returns, cov_matrix, weights, count, n_steps, first_index=first,
                                     allreduce=make_allreduce(eng.device, group, True), n_total=n_paths, **kw)
    finally:
        eng.set_allreduce_stream_ordered(False)
    cur.wait_stream(side)
    if out.get("terminal_device") is not None:
        out["terminal_device"].record_stream(cur)
    return out


_SIDE_STREAMS = {}


def _side_stream(device_index: int):
    import torch
    if device_index not in _SIDE_STREAMS:
        _SIDE_STREAMS[device_index] = torch.cuda.Stream(device_index)
    return _SIDE_STREAMS[device_index]


def emulate_sharded_quantiles(shards, alphas, device_index: int = 0):
    """Multi-rank VaR / CVaR merge on ONE GPU, for tests: one host thread and one libmcp handle
    per shard; the all-reduce is a thread barrier plus a device-side sum.  (No kernel ever
    waits on another kernel: the waiting happens on the host between launches.)"""
    import torch
    world = len(shards)
    barrier = threading.Barrier(world)
    slots = [None] * world
    results = [None] * world
    errors = []

    def worker(rank):
        try:
            eng = api.Engine(device_index)
            stream = torch.cuda.Stream(device_index)

            def allreduce(ptr, count, kind):
                slots[rank] = wrap_device_buffer(ptr, count, kind, device_index)
                torch.cuda.synchronize(device_index)
                barrier.wait()
                if rank == 0:
                    total = torch.stack(slots).sum(0)
                    for s in slots:
                        s.copy_(total)
                    torch.cuda.synchronize(device_index)
                barrier.wait()

            with torch.cuda.stream(stream):
                eng.set_stream(stream.cuda_stream)
                n_total = sum(int(s.numel()) for s in shards)
                alph = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
                import ctypes as C
                from . import _lib
                v = shards[rank].contiguous()
                code = _lib.MCP_F32 if v.dtype == torch.float32 else _lib.MCP_F64
                var_out, cvar_out = np.empty(alph.size), np.empty(alph.size)

                def _cb(ptr, count, kind, _u):
                    allreduce(ptr, count, kind)
                    return 0
                cb = _lib.ALLREDUCE_FN(_cb)
                _lib.check(eng.handle, _lib.lib().mcp_quantiles(
                    eng.handle, v.data_ptr(), _lib.MCP_DEVICE, code, v.numel(), n_total, alph.ctypes.data,
                    alph.size, var_out.ctypes.data, cvar_out.ctypes.data, cb, None))
                results[rank] = {float(a): (float(var_out[i]), float(cvar_out[i])) for i, a in enumerate(alph)}
            eng.close()
        except Exception as e:      # surface worker failures in the caller
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    assert all(r == results[0] for r in results), "ranks disagree on the merged quantiles"
    return results[0]
