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
        self.handles = (C.c_void_p * len(self.engines))(*[e.handle for e in self.engines])        # rank r = devices[r]
        check(self.engines[0].handle, lib().mcp_comm_init_all(self.handles, len(self.engines)))
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


def simulate_portfolios(mean_returns, cov_matrix, n_portfolios, *, devices, **kw):
    """`api.simulate_portfolios` over several GPUs of this process; same result object.  ONE library call
    (`mcp_portfolios_multi`): libmcp cuts the index range, drives every handle from its own host thread and merges picks, counts,
    risk range and envelope bins over its communicator.  Arrays (host only) are one allocation; every device fills its slice."""
    grp = device_group(devices)
    with grp.lock:
        return api.simulate_portfolios(mean_returns, cov_matrix, n_portfolios, _group=grp, **kw)


def simulate_paths(mean_returns, cov_matrix, weights, n_paths, n_steps=252, *, devices, **kw):
    """`api.simulate_paths` over several GPUs of this process (`mcp_paths_stats_multi`): paths sharded by index, exact VaR / CVaR of
    the whole job (radix-select histograms and tail sums all-reduced inside libmcp)."""
    grp = device_group(devices)
    with grp.lock:
        out = api.simulate_paths(mean_returns, cov_matrix, weights, n_paths, n_steps, _group=grp, **kw)
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
