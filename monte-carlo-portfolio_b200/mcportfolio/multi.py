"""One process, several GPUs: the mode the reference's caller needs.

The reference is ONE Streamlit process (``/root/reference/Procfile:1``) whose hot loop (``app.py:682-722``) runs in the
script thread; a drop-in called from there cannot be ``torchrun``.  ``simulate_portfolios(..., devices=[0, ..., 7])``
(and ``simulate_paths`` / ``frontier_envelope``) therefore drive every GPU of the box from this process: one libmcp
handle and one host thread per device (ctypes releases the GIL for the duration of a library call), one NCCL
communicator over the handles created inside libmcp (``mcp_comm_init_all``), the global index range sharded exactly as
the one-process-per-GPU route does (``dist.shard_range``; the Philox counter is the global index, so the portfolios /
paths are the same for any device count).  Picks, counts, histograms and tail sums are merged by the library over NVLink
(``comm_merge``); per-portfolio arrays never cross GPUs -- each device copies its rows into its slice of one result array.
"""
from __future__ import annotations

import concurrent.futures as cf
import ctypes as C
import threading

import numpy as np

from . import api
from ._lib import check, lib
from .dist import merge_envelopes, shard_range


class DeviceGroup:
    """Engines of several GPUs of this process, joined in one libmcp communicator, with a worker thread each."""

    def __init__(self, devices):
        self.devices = [int(d) for d in devices]
        if len(set(self.devices)) != len(self.devices):
            raise ValueError(f"devices must be distinct, got {devices}")
        self.engines = [api.get_engine(d) for d in self.devices]
        for e in self.engines:
            if e.comm_info()[1]:
                e.comm_destroy()                       # the engine leaves whatever communicator it was in
        handles = (C.c_void_p * len(self.engines))(*[e.handle for e in self.engines])
        check(self.engines[0].handle, lib().mcp_comm_init_all(handles, len(self.engines)))
        self.pool = cf.ThreadPoolExecutor(max_workers=len(self.engines), thread_name_prefix="mcp-gpu")
        self.lock = threading.Lock()                   # one job at a time: handles are not re-entrant

    @property
    def world(self):
        return len(self.engines)

    def run(self, fn):
        """fn(rank, engine) on every device concurrently; results in rank order.  Every rank must reach the same
        collectives: an exception on one rank is re-raised after all workers have returned."""
        with self.lock:
            futs = [self.pool.submit(fn, r, e) for r, e in enumerate(self.engines)]
            done = [f.exception() for f in futs]       # waits for all
            for e in done:
                if e is not None:
                    raise e
            return [f.result() for f in futs]

    def close(self):
        self.pool.shutdown(wait=True)
        for e in self.engines:
            if e.comm_info()[1]:
                e.comm_destroy()


_groups: dict[tuple, DeviceGroup] = {}
_groups_lock = threading.Lock()


def device_group(devices) -> DeviceGroup:
    key = tuple(int(d) for d in devices)
    with _groups_lock:
        g = _groups.get(key)
        if g is None:
            for k in [k for k in _groups if set(k) & set(key)]:      # a device belongs to one communicator at a time
                _groups.pop(k).close()
            g = _groups[key] = DeviceGroup(key)
        return g


def close_all():
    """Dissolve every device group of this process (their engines leave the groups' communicators)."""
    with _groups_lock:
        for k in list(_groups):
            _groups.pop(k).close()


def simulate_portfolios(mean_returns, cov_matrix, n_portfolios, *, devices, weights=None, return_arrays=True, first_index=0,
                        dtype="float32", **kw):
    """`api.simulate_portfolios` over several GPUs of this process; same result object.  Arrays (host only) are one
    allocation that every device fills its slice of."""
    if return_arrays not in (True, False):
        raise ValueError("devices=[...] returns host arrays or picks only (return_arrays True / False)")
    grp = device_group(devices)
    P, first_index = int(n_portfolios), int(first_index)
    n = len(np.asarray(mean_returns))
    _, npdt = api._dtype(dtype)
    if weights is not None:
        weights = np.asarray(weights)
        if weights.ndim != 2 or weights.shape != (P, n):
            raise ValueError(f"weights must have shape ({P}, {n}), got {weights.shape}")
    full = None
    if return_arrays:
        full = {"weights": api._result_empty((P, n), npdt), "returns": api._result_empty((P,), npdt), "risks": api._result_empty((P,), npdt),
                "sharpes": api._result_empty((P,), npdt), "accepted": api._result_empty((P,), np.uint8)}

    def work(rank, eng):
        lo, cnt = shard_range(P, rank, grp.world)
        out = None if full is None else {k: v[lo:lo + cnt] for k, v in full.items()}
        return api.simulate_portfolios(mean_returns, cov_matrix, cnt, weights=None if weights is None else weights[lo:lo + cnt],
                                       return_arrays=return_arrays, first_index=first_index + lo, dtype=dtype, device=eng.device,
                                       out=out, comm_merge=True, **kw)

    parts = grp.run(work)
    accepted_per_part = [p.n_accepted for p in parts]      # before parts[0] becomes the merged result
    r = parts[0]
    r.n_requested = P
    r.n_accepted = sum(accepted_per_part)
    r.extra["kernel_ms_per_device"] = [p.kernel_ms for p in parts]
    r.kernel_ms = max(r.extra["kernel_ms_per_device"])
    r.extra["devices"] = list(grp.devices)
    if full is not None:
        skipped = r.n_accepted < P
        if skipped:      # app.py:706-707: the reference's arrays simply do not contain skipped portfolios
            for name in ("weights", "returns", "risks", "sharpes"):
                setattr(r, name, np.concatenate([getattr(p, name) for p in parts]))
            r.accepted = full["accepted"]
        else:
            r.weights, r.returns, r.risks, r.sharpes, r.accepted = (full[k] for k in ("weights", "returns", "risks", "sharpes", "accepted"))
        # `index` = position in the returned arrays (app.py:747): the owning device knows the local position
        offsets = np.concatenate([[0], np.cumsum(accepted_per_part)])
        for pick in ("max_sharpe", "target_risk"):
            rec = getattr(r, pick)
            if rec is None:
                continue
            g = rec["global_index"] - first_index
            for rank, p in enumerate(parts):
                lo, cnt = shard_range(P, rank, grp.world)
                if lo <= g < lo + cnt:
                    rec = dict(rec, index=int(offsets[rank]) + int(getattr(p, pick)["index"]))
            setattr(r, pick, rec)
    return r


def simulate_paths(mean_returns, cov_matrix, weights, n_paths, n_steps=252, *, devices, first_index=0, **kw):
    """`api.simulate_paths` over several GPUs of this process: paths sharded by index, exact VaR / CVaR of the whole job
    (radix-select histograms and tail sums all-reduced inside libmcp)."""
    grp = device_group(devices)
    M, first_index = int(n_paths), int(first_index)

    def work(rank, eng):
        lo, cnt = shard_range(M, rank, grp.world)
        return api.simulate_paths(mean_returns, cov_matrix, weights, cnt, n_steps, first_index=first_index + lo, device=eng.device,
                                  comm_merge=True, n_total=M, return_terminal=False, **kw)

    parts = grp.run(work)
    assert all(p["stats"] == parts[0]["stats"] for p in parts), "devices disagree on the merged quantiles"
    out = dict(parts[0])
    out["kernel_ms"] = max(p["kernel_ms"] for p in parts)
    out["quantile_ms"] = max(p["quantile_ms"] for p in parts)
    out["devices"] = list(grp.devices)
    return out


def frontier_envelope(mean_returns, cov_matrix, n_portfolios, n_bins=512, *, devices, risk_range=None, first_index=0, **kw):
    """`api.frontier_envelope` over several GPUs of this process (C5): every device sweeps its index block once with its
    (risk, return) pairs kept in its own HBM, the attained risk range is merged, every device bins its arrays, the bins are
    merged (larger return, then lower index)."""
    grp = device_group(devices)
    P, first_index = int(n_portfolios), int(first_index)
    kw = {k: v for k, v in kw.items() if k != "return_arrays"}

    def sweep(rank, eng):
        lo, cnt = shard_range(P, rank, grp.world)
        return api.simulate_portfolios(mean_returns, cov_matrix, cnt, first_index=first_index + lo, device=eng.device,
                                       return_arrays="device-metrics", comm_merge=True, **kw)

    parts = grp.run(sweep)
    if sum(p.extra["n_accepted_global"] for p in parts[:1]) == 0:
        raise ValueError("no portfolio satisfied the bounds; the envelope is empty")
    if risk_range is None:
        risk_range = api._widen(parts[0].risk_range)          # merged in the library: the whole job's range

    def bins(rank, eng):
        lo, cnt = shard_range(P, rank, grp.world)
        if cnt == 0:
            return None
        return api.envelope_from_arrays(parts[rank].risks, parts[rank].returns, n_bins, risk_range, first_index=first_index + lo,
                                        device=eng.device)

    envs = [e for e in grp.run(bins) if e is not None]
    r = parts[0]
    best, idx = merge_envelopes([e["best_return"] for e in envs], [e["best_index"] for e in envs])
    r.extra["envelope"] = {"edges": envs[0]["edges"], "best_return": best, "best_index": idx}
    r.extra["risk_range_global"] = risk_range
    r.extra["devices"] = list(grp.devices)
    r.n_requested = P
    r.n_accepted = r.extra["n_accepted_global"]
    r.kernel_ms = max(p.kernel_ms for p in parts)
    r.risks = r.returns = None
    return r
