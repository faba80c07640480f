"""Python host side of the B200 Monte Carlo portfolio path.

Mirrors the reference's function-shaped interface for the path
(``/root/reference/app.py``): ``efficient_frontier`` keeps the exact signature and
return layout of app.py:265-284; ``simulate_portfolios`` is the inline loop
app.py:682-722 collapsed into one call (same array names / order: risks, returns,
weights, metrics) plus the two picks; ``simulate_paths`` and ``frontier_envelope``
are the north-star additions.  All arithmetic runs in libmcp.so (hand-written
sm_100a CUDA) through the C ABI of include/mcp.h -- this module only validates
arguments, owns buffers and shapes results.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import (MCP_DEVICE, MCP_F32, MCP_F64, MCP_HOST, MCP_NO_INDEX, McpError, PathParams,
                   PortfolioOut, PortfolioParams, check, lib)

_DTYPES = {"float32": (MCP_F32, np.float32), "float64": (MCP_F64, np.float64),
           np.float32: (MCP_F32, np.float32), np.float64: (MCP_F64, np.float64)}


def _dtype(dtype):
    try:
        key = dtype if isinstance(dtype, str) else np.dtype(dtype).type
        return _DTYPES[key]
    except (KeyError, TypeError):
        raise ValueError(f"dtype must be 'float32' or 'float64', got {dtype!r}") from None


class Engine:
    """One libmcp handle = one device.  Not re-entrant (one call at a time per engine)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().mcp_create(int(device), C.byref(self._h))
        if rc != 0:
            msg = lib().mcp_last_error(None)
            raise McpError(rc, msg.decode() if msg else "mcp_create failed")
        self.device = int(device)
        self._finalizer = weakref.finalize(self, lib().mcp_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def close(self):
        self._finalizer()

    def info(self):
        d = _lib.DeviceInfo()
        check(self._h, lib().mcp_device_info(self._h, C.byref(d)))
        return {"name": d.name.decode(), "sm_count": d.sm_count, "cc": (d.cc_major, d.cc_minor),
                "max_smem_per_block": d.max_smem_per_block, "total_mem": d.total_mem}

    def launch_count(self) -> int:
        return int(lib().mcp_launch_count(self._h))

    def copy_stream(self):
        """A (non-blocking) torch stream of this device for copies that overlap the engine's kernels."""
        import torch
        st = getattr(self, "_copy_stream", None)
        if st is None:
            st = self._copy_stream = torch.cuda.Stream(device=self.device)
        return st

    def last_kernel_ms(self) -> float:
        return float(lib().mcp_last_kernel_ms(self._h))

    def set_stream(self, cuda_stream_ptr: int | None):
        """Order this engine's work on a caller stream (e.g. torch.cuda.current_stream().cuda_stream).

        torch's default stream is the LEGACY default stream, whose handle is 0 -- which `mcp_set_stream` reads as
        "use the handle's own (non-blocking) stream", a stream that does not synchronise with the legacy one.  A 0
        handle is therefore passed on as `cudaStreamLegacy` (0x1): libmcp's kernels then queue behind whatever torch
        enqueued before the call (dtype conversions, a producer kernel of `weights`) and torch's later work behind them.
        `None` selects the handle's own stream (HOST-space calls that share nothing with torch)."""
        ptr = 0 if cuda_stream_ptr is None else (int(cuda_stream_ptr) or _CUDA_STREAM_LEGACY)
        check(self._h, lib().mcp_set_stream(self._h, C.c_void_p(ptr)))

    def synchronize(self):
        check(self._h, lib().mcp_synchronize(self._h))

    def set_allreduce_stream_ordered(self, on: bool) -> None:
        """Promise that all-reduce callbacks only enqueue on the handle's stream (see include/mcp.h)."""
        check(self._h, lib().mcp_set_allreduce_stream_ordered(self._h, 1 if on else 0))

    # ---- multi-GPU: the NCCL communicator inside libmcp (include/mcp.h, mcp_comm_*) ----
    def comm_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        """Join a communicator: every rank passes the same 128-byte id (`comm_unique_id()` of rank 0)."""
        if len(unique_id) != _lib.MCP_COMM_ID_BYTES:
            raise ValueError(f"the communicator id is {_lib.MCP_COMM_ID_BYTES} bytes")
        buf = C.create_string_buffer(bytes(unique_id), _lib.MCP_COMM_ID_BYTES)
        check(self._h, lib().mcp_comm_init(self._h, buf, int(rank), int(nranks)))

    def comm_destroy(self) -> None:
        check(self._h, lib().mcp_comm_destroy(self._h))

    def comm_info(self):
        """(rank, nranks); nranks = 0 when the engine has no communicator."""
        r, n = C.c_int(), C.c_int()
        check(self._h, lib().mcp_comm_info(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def allgather(self, values: np.ndarray) -> np.ndarray:
        """Every rank's (same-shape) array on every rank, stacked along a new first axis, through libmcp's NCCL."""
        v = np.ascontiguousarray(values)
        _, world = self.comm_info()
        out = np.empty((world,) + v.shape, dtype=v.dtype)
        check(self._h, lib().mcp_comm_allgather(self._h, v.ctypes.data, v.nbytes, out.ctypes.data))
        return out

    def allreduce(self, values: np.ndarray, op: str = "sum") -> np.ndarray:
        """Element-wise reduction over ranks of a float64 or uint64 array (op: 'sum', 'min', 'max')."""
        v = np.ascontiguousarray(values)
        kinds = {("float64", "sum"): _lib.MCP_REDUCE_F64_SUM, ("float64", "min"): _lib.MCP_REDUCE_F64_MIN,
                 ("float64", "max"): _lib.MCP_REDUCE_F64_MAX, ("uint64", "sum"): _lib.MCP_REDUCE_U64_SUM,
                 ("uint64", "max"): _lib.MCP_REDUCE_U64_MAX}
        kind = kinds.get((v.dtype.name, op))
        if kind is None:
            raise ValueError(f"allreduce supports float64 sum/min/max and uint64 sum/max, got {v.dtype.name} {op}")
        out = v.copy()
        check(self._h, lib().mcp_comm_allreduce(self._h, out.ctypes.data, out.size, kind))
        return out

    def measure_fma_peak(self, dtype="float32") -> float:
        code = 2 if dtype == "float32x2" else _dtype(dtype)[0]
        out = C.c_double()
        check(self._h, lib().mcp_measure_fma_peak(self._h, code, C.byref(out)))
        return out.value


_CUDA_STREAM_LEGACY = 0x1          # cudaStreamLegacy (driver_types.h): the legacy default stream as an explicit handle

_engines: dict[int, Engine] = {}


def comm_unique_id() -> bytes:
    """A fresh communicator id (ncclGetUniqueId): create it on one rank, ship it to the others, pass it to comm_init."""
    buf = C.create_string_buffer(_lib.MCP_COMM_ID_BYTES)
    rc = lib().mcp_comm_unique_id(buf)
    if rc != 0:
        msg = lib().mcp_last_error(None)
        raise McpError(rc, msg.decode() if msg else "mcp_comm_unique_id failed")
    return buf.raw


def get_engine(device: int | None = None) -> Engine:
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    eng = _engines.get(device)
    if eng is None:
        eng = _engines[device] = Engine(device)
    return eng


class _PinnedPool:
    """Caching allocator for page-locked result buffers.

    Pinning memory costs about as much as copying it (cudaHostAlloc of 77 MB: tens of ms), so the buffers behind the
    arrays a call returns are recycled: when the last numpy view of a block dies, the block goes back to the pool
    (size class = next multiple of 1 MiB) instead of being unpinned, and the next call of the same shape reuses it.
    Results therefore land in page-locked memory at PCIe rate without the caller passing `out=`.  The pool keeps at
    most `limit` bytes of idle blocks; beyond that blocks are freed."""

    GRAIN = 1 << 20

    def __init__(self, limit=int(os.environ.get("MCP_PINNED_POOL_BYTES", 4 << 30))):
        import threading
        self.free, self.idle, self.limit, self.lock = {}, 0, limit, threading.Lock()

    def take(self, nbytes):
        cap = max(self.GRAIN, -(-nbytes // self.GRAIN) * self.GRAIN)
        with self.lock:
            blocks = self.free.get(cap)
            if blocks:
                self.idle -= cap
                return blocks.pop(), cap
        p = C.c_void_p()
        rc = lib().mcp_host_alloc(cap, C.byref(p))
        if rc != 0:
            raise McpError(rc, f"mcp_host_alloc({cap}) failed")
        return p.value, cap

    def give(self, ptr, cap):
        with self.lock:
            if self.idle + cap <= self.limit:
                self.free.setdefault(cap, []).append(ptr)
                self.idle += cap
                return
        lib().mcp_host_free(C.c_void_p(ptr))

    def clear(self):
        with self.lock:
            blocks, self.free, self.idle = self.free, {}, 0
        for cap, ptrs in blocks.items():
            for ptr in ptrs:
                lib().mcp_host_free(C.c_void_p(ptr))


_pinned_pool = None


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array on page-locked host memory (the fast path for HOST-space buffers), from a caching pool: the block
    returns to the pool when the array and all its views are garbage."""
    global _pinned_pool
    if _pinned_pool is None:
        _pinned_pool = _PinnedPool()
    dt = np.dtype(dtype)
    count = int(np.prod(shape))
    ptr, cap = _pinned_pool.take(max(count * dt.itemsize, 1))
    buf = (C.c_char * cap).from_address(ptr)
    weakref.finalize(buf, _pinned_pool.give, ptr, cap)
    return np.frombuffer(buf, dtype=dt, count=count).reshape(shape)


_PINNED_MIN_BYTES = 1 << 16      # smaller results are not worth a pinned block


def _result_empty(shape, dtype):
    """Result array of a HOST-space call: page-locked (pooled) when it is big enough for the copy rate to matter."""
    dt = np.dtype(dtype)
    if int(np.prod(shape)) * dt.itemsize >= _PINNED_MIN_BYTES:
        return pinned_empty(shape, dt)
    return np.empty(shape, dtype=dt)


# ------------------------------------------------------------------------------------------
# argument handling shared by the entry points
# ------------------------------------------------------------------------------------------

def _mu_sigma(mean_returns, cov_matrix):
    """Accepts pd.Series / pd.DataFrame or ndarrays, like the app does (app.py:679-680)."""
    mu = np.ascontiguousarray(np.asarray(mean_returns, dtype=np.float64))
    sigma = np.ascontiguousarray(np.asarray(cov_matrix, dtype=np.float64))
    if mu.ndim != 1 or mu.size < 1:
        raise ValueError(f"mean_returns must be a non-empty vector, got shape {mu.shape}")
    n = mu.size
    if sigma.shape != (n, n):
        raise ValueError(f"cov_matrix must have shape ({n}, {n}), got {sigma.shape}")
    if not (np.all(np.isfinite(mu)) and np.all(np.isfinite(sigma))):
        raise ValueError("mean_returns / cov_matrix contain non-finite values")
    return mu, sigma, n


def _bounds(v, n, name):
    if v is None:
        return None
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)))
    if a.shape != (n,):
        raise ValueError(f"{name} must have {n} entries")
    return a


def _is_device_tensor(x) -> bool:
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _ptr(a):
    if a is None:
        return None
    if _is_device_tensor(a):
        return a.data_ptr()
    return a.ctypes.data


@dataclass
class PortfolioResult:
    """What one method's pass of app.py:682-722 leaves behind, plus the two picks."""
    risks: np.ndarray | None          # all_risks   (app.py:719)
    returns: np.ndarray | None        # all_returns (app.py:720)
    weights: np.ndarray | None        # all_weights (app.py:721)
    sharpes: np.ndarray | None        # all_metrics (app.py:722, metric = sharpe)
    max_sharpe: dict | None           # np.argmax(sharpe) pick (app.py:672, 747)
    target_risk: dict | None          # argmin |risk - target| pick (README.md:4)
    n_requested: int = 0
    n_accepted: int = 0
    risk_range: tuple = (float("nan"), float("nan"))
    kernel_ms: float = 0.0
    accepted: np.ndarray | None = None
    extra: dict = field(default_factory=dict)

    def __iter__(self):               # risks, returns, weights, metrics = simulate_portfolios(...)
        return iter((self.risks, self.returns, self.weights, self.sharpes))


def _selection_dict(sel, wbuf, first_index, positions, n_local=None):
    if sel.index == MCP_NO_INDEX:
        return None
    g = int(sel.index)
    d = {"index": g, "global_index": g, "weights": wbuf.copy(), "ret": sel.ret, "risk": sel.risk,
         "sharpe": sel.sharpe, "key": sel.key}
    local = n_local is None or first_index <= g < first_index + n_local       # with comm_merge the pick may be another rank's
    if positions is not None and local:         # index into the returned (accepted-only) arrays, as app.py:747
        d["index"] = int(positions[g - first_index])
    return d


def simulate_portfolios(mean_returns, cov_matrix, n_portfolios, *, risk_free=0.0, risk_target=0.30,
                        min_weights=None, max_weights=None, weights=None, seed=0, dtype="float32",
                        return_arrays=True, first_index=0, keep_last=False, max_tries=100,
                        device=None, out=None, n_bins=0, risk_range=None, devices=None, comm_merge=False,
                        philox_rounds=10, _group=None) -> PortfolioResult:
    """Random-weight portfolio sweep: the loop of app.py:699-722 in one call.

    weights=None      flat-Dirichlet weights generated in-kernel (Philox4x32-10, counter =
                      first_index + i), at most `max_tries` draws per portfolio against
                      [min_weights, max_weights] (app.py:700-707; skipped on exhaustion
                      unless keep_last, which is efficient_frontier's behaviour, app.py:277).
    weights=(P, N)    supplied-weights (parity) mode: evaluated as given, no RNG.  numpy
                      array (HOST space, copies inside the call) or CUDA torch tensor
                      (DEVICE space, no copies; arrays come back as torch tensors).  With
                      dtype='float32' and FP64 weights the FP32 sweep screens and the picks are
                      decided in FP64 among the near-ties (same index as the FP64 reference).
    risk_free         subtracted raw, as app.py:711 (pass 3.0 for the app's default widget value).
    return_arrays     False: selections only, zero HBM write-back (C3 sizes).  'device': arrays stay on
                      the GPU as torch tensors.  'device-metrics': only risks and returns stay on the GPU
                      (8 bytes per portfolio; skipped portfolios hold NaN) -- what the single-sweep
                      frontier envelope keeps in HBM.
    out               optional dict of preallocated arrays ('weights', 'returns', 'risks',
                      'sharpes', 'accepted') to reuse pinned buffers across calls.
    n_bins, risk_range  frontier envelope: per risk bin over [lo, hi] the maximum return and the
                      first global index attaining it -> result.extra['envelope'].
    devices           list of GPU indices: ONE process drives them all (one libmcp handle and host thread per GPU,
                      an NCCL communicator inside libmcp); the index range is sharded, the picks are merged in the
                      library, arrays (if requested) are concatenated in index order.  This is what the one-process
                      Streamlit app uses (Procfile:1).  See mcportfolio.multi.
    comm_merge        this call is one rank's shard of a job whose ranks share a libmcp communicator
                      (mcportfolio.dist.init_comm / mcp_comm_init): picks, counts, risk range and envelope bins are
                      merged across ranks inside the library.
    philox_rounds     10 (default; the generator oracle/philox_np.py restates) or 7 (Random123's Crush-resistant
                      minimum: a different, cheaper stream).
    Returns a PortfolioResult; arrays hold accepted portfolios only (P' <= P rows).
    """
    if devices is not None and len(devices) > 1:
        from . import multi
        return multi.simulate_portfolios(mean_returns, cov_matrix, n_portfolios, devices=list(devices), risk_free=risk_free,
                                         risk_target=risk_target, min_weights=min_weights, max_weights=max_weights,
                                         weights=weights, seed=seed, dtype=dtype, return_arrays=return_arrays,
                                         first_index=first_index, keep_last=keep_last, max_tries=max_tries, n_bins=n_bins,
                                         risk_range=risk_range, philox_rounds=philox_rounds, out=out)
    if devices is not None and len(devices) == 1:
        device = devices[0]
    mu, sigma, n = _mu_sigma(mean_returns, cov_matrix)
    code, npdt = _dtype(dtype)
    lo = _bounds(min_weights, n, "min_weights")
    hi = _bounds(max_weights, n, "max_weights")
    P = int(n_portfolios)
    if P < 0:
        raise ValueError("n_portfolios must be >= 0")

    if weights is not None and not _is_device_tensor(weights):
        wshape = np.shape(weights)
        if len(wshape) != 2 or wshape[1] != n:
            raise ValueError(f"weights must have shape (P, {n}), got {wshape}")
        if wshape[0] != P:
            raise ValueError(f"weights has {wshape[0]} rows but n_portfolios={P}")
    eng = get_engine(device) if _group is None else _group.engines[0]
    if _group is not None and (_is_device_tensor(weights) or return_arrays not in (True, False) or comm_merge):
        raise ValueError("devices=[...] takes host weights and returns host arrays or picks only (return_arrays True / False)")
    if weights is not None and not _is_device_tensor(weights) and return_arrays in ("device", "device-metrics"):
        # host weights with device-resident results: the call runs in DEVICE space, so the rows go up first
        import torch
        weights = torch.from_numpy(np.ascontiguousarray(np.asarray(weights))).to(torch.device("cuda", eng.device))
    device_mode = _is_device_tensor(weights)
    w_in = None
    recheck = None
    if weights is not None:
        if device_mode:
            import torch
            if weights.device.index != eng.device:
                raise ValueError(f"weights live on cuda:{weights.device.index} but the engine runs on cuda:{eng.device}")
            want = torch.float32 if code == MCP_F32 else torch.float64
            w_in = weights.to(want).contiguous()
            if code == MCP_F32 and weights.dtype == torch.float64:
                recheck = weights.contiguous()
            shape = tuple(w_in.shape)
        else:
            src = np.asarray(weights)
            w_in = np.ascontiguousarray(src, dtype=npdt)
            if code == MCP_F32 and src.dtype == np.float64:
                recheck = np.ascontiguousarray(src)
            shape = w_in.shape
        if len(shape) != 2 or shape[1] != n:
            raise ValueError(f"weights must have shape (P, {n}), got {shape}")
        if shape[0] != P:
            raise ValueError(f"weights has {shape[0]} rows but n_portfolios={P}")

    params = PortfolioParams()
    params.n_assets, params.dtype = n, code
    params.n_portfolios, params.first_index, params.seed = P, int(first_index), int(seed) & MCP_NO_INDEX
    params.risk_free, params.risk_target = float(risk_free), float(risk_target)
    params.min_weights, params.max_weights = _ptr(lo), _ptr(hi)
    params.max_tries, params.keep_last = int(max_tries), int(bool(keep_last))
    params.comm_merge, params.philox_rounds = int(bool(comm_merge)), int(philox_rounds)
    metrics_only = return_arrays == "device-metrics"
    params.space = MCP_DEVICE if device_mode or return_arrays == "device" or metrics_only else MCP_HOST
    params.weights_in = _ptr(w_in)
    params.weights_recheck = _ptr(recheck)
    n_bins = int(n_bins)
    if n_bins:
        if risk_range is None or not (np.isfinite(risk_range[0]) and np.isfinite(risk_range[1]) and risk_range[1] > risk_range[0]):
            raise ValueError("the envelope needs risk_range=(lo, hi) with finite lo < hi")
        params.n_bins, params.risk_lo, params.risk_hi = n_bins, float(risk_range[0]), float(risk_range[1])

    res = PortfolioOut()
    arrays = {}
    if return_arrays:
        names = (("weights", (P, n), npdt), ("returns", (P,), npdt), ("risks", (P,), npdt),
                 ("sharpes", (P,), npdt), ("accepted", (P,), np.uint8))
        if metrics_only:
            names = names[1:3]
        if params.space == MCP_DEVICE:
            import torch
            tdt = torch.float32 if code == MCP_F32 else torch.float64
            dev = torch.device("cuda", eng.device)
            for name, shape, dt in names:
                arrays[name] = torch.empty(shape, dtype=torch.uint8 if dt is np.uint8 else tdt, device=dev)
        else:
            for name, shape, dt in names:
                a = None if out is None else out.get(name)
                if a is not None and (a.shape != shape or a.dtype != dt or not a.flags.c_contiguous):
                    raise ValueError(f"out[{name!r}] must be C-contiguous {shape} {np.dtype(dt)}")
                arrays[name] = a if a is not None else _result_empty(shape, dt)
        for name in arrays:
            setattr(res, name, _ptr(arrays[name]))
    bin_ret = np.empty(n_bins)
    bin_idx = np.empty(n_bins, dtype=np.uint64)
    if n_bins:
        res.bin_best_return, res.bin_best_index = bin_ret.ctypes.data, bin_idx.ctypes.data
    ws = np.empty(n)
    wt = np.empty(n)
    res.max_sharpe.weights = ws.ctypes.data
    res.target_risk.weights = wt.ctypes.data

    if params.space == MCP_DEVICE:
        import torch
        eng.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
    per_device_ms = None
    if _group is not None:
        # ONE library call drives every GPU of the group: libmcp cuts the index range, runs a host thread per handle and merges
        # the results over its communicator (mcp_multi.cu); arrays are the whole job's, each device fills its slice
        for e in _group.engines:
            e.set_stream(None)
        per_device_ms = np.zeros(_group.world)
        check(eng.handle, lib().mcp_portfolios_multi(_group.handles, _group.world, C.byref(params), mu.ctypes.data, sigma.ctypes.data,
                                                     C.byref(res), per_device_ms.ctypes.data))
    else:
        check(eng.handle, lib().mcp_portfolios(eng.handle, C.byref(params), mu.ctypes.data, sigma.ctypes.data, C.byref(res)))
    if comm_merge:
        extra_merge = {"n_accepted_global": int(res.n_accepted_global)}
    else:
        extra_merge = {}
    if res.recheck_overflow:
        extra_merge["recheck_overflow"] = True

    n_acc = int(res.n_accepted)
    positions = None
    if metrics_only:
        positions = None                        # rows are not compacted: row i is portfolio first_index + i
    elif return_arrays and n_acc < P:
        # skipped portfolios (app.py:706-707): the reference's arrays simply do not contain them
        acc = arrays["accepted"]
        if params.space == MCP_DEVICE:
            mask = acc.bool()
            positions = (mask.cumsum(0) - 1).cpu().numpy()
        else:
            mask = acc.astype(bool)
            positions = np.cumsum(mask) - 1
        for name in ("weights", "returns", "risks", "sharpes"):
            arrays[name] = arrays[name][mask]
    elif return_arrays:
        positions = _Identity()
    extra = dict(extra_merge)
    if _group is not None:
        extra["devices"] = list(_group.devices)
        extra["kernel_ms_per_device"] = per_device_ms.tolist()
        extra["n_accepted_global"] = int(res.n_accepted_global)
    if n_bins:
        idx = bin_idx.astype(np.int64)
        idx[bin_idx == np.uint64(MCP_NO_INDEX)] = -1
        extra["envelope"] = {"edges": np.linspace(params.risk_lo, params.risk_hi, n_bins + 1),
                             "best_return": bin_ret, "best_index": idx}
    return PortfolioResult(
        risks=arrays.get("risks"), returns=arrays.get("returns"), weights=arrays.get("weights"),
        sharpes=arrays.get("sharpes"),
        max_sharpe=_selection_dict(res.max_sharpe, ws, int(first_index), positions, P if comm_merge else None),
        target_risk=_selection_dict(res.target_risk, wt, int(first_index), positions, P if comm_merge else None),
        n_requested=P, n_accepted=n_acc, risk_range=(res.risk_min, res.risk_max),
        kernel_ms=res.kernel_ms, accepted=arrays.get("accepted"), extra=extra)


class _Identity:
    def __getitem__(self, i):
        return i


def efficient_frontier(mean_returns, cov_matrix, points=200, min_weights=None, max_weights=None, *,
                       seed=0, dtype="float64", device=None):
    """Drop-in for the reference's ``efficient_frontier`` (app.py:265-284).

    Same signature, same return layout: ``results`` (3, points) = [std; return; return/std]
    (Sharpe WITHOUT risk-free, app.py:282) and ``weights`` (points, N), both FP64 ndarrays.
    Keeps the last draw when 100 tries are exhausted (app.py:277).  The reference draws
    from the global legacy MT19937 state; here the stream is Philox keyed by ``seed``.
    """
    r = simulate_portfolios(mean_returns, cov_matrix, points, risk_free=0.0, min_weights=min_weights,
                            max_weights=max_weights, seed=seed, dtype=dtype, keep_last=True,
                            return_arrays=True, device=device)
    results = np.zeros((3, int(points)))
    results[0] = r.risks
    results[1] = r.returns
    results[2] = r.sharpes
    return results, np.asarray(r.weights, dtype=np.float64)


def frontier_envelope(mean_returns, cov_matrix, n_portfolios, n_bins=512, *, risk_range=None, **kw):
    """Efficient-frontier envelope binned by risk (config C5), plus the two picks.

    Replaces the reference's scatter of every portfolio (app.py:726-736).  With risk_range=None the
    attained [risk_min, risk_max] defines the bins: when the per-portfolio (risk, return) pairs fit in
    HBM (8 bytes each) ONE sweep keeps them there and a bandwidth-bound pass bins them; otherwise a first
    sweep (no write-back) finds the range and the counter-based generator reproduces exactly the same
    portfolios for a second, binning sweep.  Both routes give identical bins.
    Returns the PortfolioResult of the binning sweep; result.extra['envelope'] holds
    'edges' (n_bins + 1), 'best_return' (-inf for empty bins) and 'best_index' (global, -1 if empty).
    """
    devices = kw.pop("devices", None)
    if devices is not None and len(devices) > 1:
        from . import multi
        return multi.frontier_envelope(mean_returns, cov_matrix, n_portfolios, n_bins, devices=list(devices), risk_range=risk_range, **kw)
    if devices is not None and len(devices) == 1:
        kw["device"] = devices[0]
    kw.setdefault("return_arrays", False)
    if risk_range is None and kw["return_arrays"] is False and kw.get("weights") is None and _metrics_fit(n_portfolios, kw):
        # ONE sweep: risks / returns stay in HBM (8 B per portfolio), the attained range comes out of the same
        # sweep, the binning pass then reads the arrays back at HBM speed (mcp_envelope_arrays)
        r = simulate_portfolios(mean_returns, cov_matrix, n_portfolios, **{**kw, "return_arrays": "device-metrics"})
        if r.n_accepted == 0:
            raise ValueError("no portfolio satisfied the bounds; the envelope is empty")
        r.extra["envelope"] = envelope_from_arrays(r.risks, r.returns, n_bins, _widen(r.risk_range),
                                                   first_index=kw.get("first_index", 0), device=kw.get("device"))
        r.risks = r.returns = None              # scratch of the envelope, not part of the result
        return r
    if risk_range is None:
        probe = simulate_portfolios(mean_returns, cov_matrix, n_portfolios, **{**kw, "return_arrays": False})
        if probe.n_accepted == 0:
            raise ValueError("no portfolio satisfied the bounds; the envelope is empty")
        risk_range = _widen(probe.risk_range)
    return simulate_portfolios(mean_returns, cov_matrix, n_portfolios, n_bins=n_bins, risk_range=risk_range, **kw)


def _widen(rng):
    lo, hi = rng
    if not hi > lo:
        hi = lo + max(abs(lo), 1.0) * 1e-6
    return (lo, hi)


def _metrics_fit(n_portfolios, kw) -> bool:
    """Do the risk / return arrays of the single-sweep envelope fit comfortably in free HBM?"""
    import torch
    eng = get_engine(kw.get("device"))
    es = 8 if _dtype(kw.get("dtype", "float32"))[0] == MCP_F64 else 4
    free, _ = torch.cuda.mem_get_info(eng.device)
    return 2 * es * int(n_portfolios) <= free // 2


def envelope_from_arrays(risks, returns, n_bins, risk_range, *, first_index=0, device=None):
    """Frontier envelope of (risk, return) arrays that are already on the GPU (CUDA torch tensors of one dtype):
    per risk bin over [lo, hi] the maximum return and the first global index (first_index + row) attaining it;
    NaN rows are ignored.  Same result as the n_bins / risk_range options of `simulate_portfolios`."""
    import torch
    if not (_is_device_tensor(risks) and _is_device_tensor(returns)):
        raise TypeError("envelope_from_arrays takes CUDA torch tensors")
    if risks.dtype != returns.dtype or risks.dtype not in (torch.float32, torch.float64) or risks.shape != returns.shape or risks.dim() != 1:
        raise ValueError("risks and returns must be 1-D tensors of the same length and dtype (float32 / float64)")
    n_bins = int(n_bins)
    lo, hi = float(risk_range[0]), float(risk_range[1])
    if not (np.isfinite(lo) and np.isfinite(hi) and hi > lo):
        raise ValueError("the envelope needs risk_range=(lo, hi) with finite lo < hi")
    risks, returns = risks.contiguous(), returns.contiguous()
    eng = get_engine(device if device is not None else risks.device.index)
    eng.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
    bin_ret = np.empty(n_bins)
    bin_idx = np.empty(n_bins, dtype=np.uint64)
    check(eng.handle, lib().mcp_envelope_arrays(eng.handle, MCP_F32 if risks.dtype == torch.float32 else MCP_F64,
                                                risks.data_ptr(), returns.data_ptr(), risks.numel(), int(first_index), lo, hi, n_bins,
                                                bin_ret.ctypes.data, bin_idx.ctypes.data))
    idx = bin_idx.astype(np.int64)
    idx[bin_idx == np.uint64(MCP_NO_INDEX)] = -1
    return {"edges": np.linspace(lo, hi, n_bins + 1), "best_return": bin_ret, "best_index": idx}


# ------------------------------------------------------------------------------------------
# correlated paths + VaR / CVaR
# ------------------------------------------------------------------------------------------

def simulate_paths(mean_returns, cov_matrix, weights, n_paths, n_steps=252, *, dt=1.0 / 252,
                   alphas=(0.95, 0.99), normals=None, seed=0, dtype="float32", first_index=0,
                   return_terminal=None, device=None, allreduce=None, n_total=None, devices=None, comm_merge=False,
                   philox_rounds=10, _group=None):
    """Correlated-return paths (Cholesky of Sigma, per-asset cumulative product) -> VaR / CVaR.

    Not in the reference: r = mu dt + sqrt(dt) L z, V *= 1 + r (the
    compounding of app.py:253), terminal = w . V_T - 1; VaR / CVaR follow app.py:258-263.
    normals=(M, S, N) switches to supplied-normals (parity) mode.
    One library call (mcp_paths_stats): the path kernel, the exact radix select and the tail sums run back to back on
    the GPU and the results come back with a single copy.
    devices=[...]   one process, several GPUs: paths sharded by index, histograms all-reduced inside libmcp.
    comm_merge      this call is one rank's shard (first_index, n_paths) of a job of `n_total` paths whose ranks share
                    a libmcp communicator: VaR / CVaR are the whole job's, exact, on every rank.
    allreduce       legacy hook: a Python callback that sums a device buffer across ranks (mcportfolio.dist.make_allreduce).
    Returns {'stats': {alpha: (var, cvar)}, 'terminal': ndarray | None, 'kernel_ms': float, ...}.
    """
    if devices is not None and len(devices) > 1:
        from . import multi
        return multi.simulate_paths(mean_returns, cov_matrix, weights, n_paths, n_steps, devices=list(devices), dt=dt, alphas=alphas,
                                    seed=seed, dtype=dtype, first_index=first_index, philox_rounds=philox_rounds)
    if devices is not None and len(devices) == 1:
        device = devices[0]
    mu, sigma, n = _mu_sigma(mean_returns, cov_matrix)
    code, npdt = _dtype(dtype)
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
    if w.shape != (n,):
        raise ValueError(f"weights must have shape ({n},), got {w.shape}")
    M, S = int(n_paths), int(n_steps)
    if M < (0 if comm_merge else 1) or S < 1:
        raise ValueError("n_paths and n_steps must be >= 1")
    alphas = tuple(float(a) for a in alphas)
    if not (1 <= len(alphas) <= _lib.MCP_MAX_ALPHAS):
        raise ValueError(f"between 1 and {_lib.MCP_MAX_ALPHAS} alphas are supported")
    eng = get_engine(device) if _group is None else _group.engines[0]
    if _group is not None:
        if normals is not None or allreduce is not None or comm_merge:
            raise ValueError("devices=[...] simulates Philox paths (no supplied normals, no caller-side merge)")
        return_terminal = False                  # terminal values stay on the devices
    if return_terminal is None:
        return_terminal = M <= (1 << 22)
    z_dev = terminal = None
    if normals is not None or return_terminal or allreduce is not None:
        import torch   # device memory for the normals / terminal values (plumbing only)
        dev = torch.device("cuda", eng.device)
        eng.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
        tdt = torch.float32 if code == MCP_F32 else torch.float64
        if normals is not None:
            if _is_device_tensor(normals):
                z_dev = normals.to(tdt).contiguous()
            else:
                z = np.ascontiguousarray(np.asarray(normals), dtype=npdt)
                if z.shape != (M, S, n):
                    raise ValueError(f"normals must have shape ({M}, {S}, {n}), got {z.shape}")
                z_dev = torch.from_numpy(z).to(dev)
            if tuple(z_dev.shape) != (M, S, n):
                raise ValueError(f"normals must have shape ({M}, {S}, {n}), got {tuple(z_dev.shape)}")
        if return_terminal or allreduce is not None:
            terminal = torch.empty(M, dtype=tdt, device=dev)
    else:
        eng.set_stream(None)                     # nothing of torch's is involved: the handle's own stream
    p = PathParams()
    p.n_assets, p.dtype, p.n_paths, p.first_index = n, code, M, int(first_index)
    p.seed, p.n_steps, p.space, p.dt = int(seed) & MCP_NO_INDEX, S, MCP_DEVICE, float(dt)
    p.normals_in = _ptr(z_dev)
    p.philox_rounds = int(philox_rounds)
    if allreduce is not None:
        # legacy route: the caller's callback sums the histograms (two library calls, a host round trip between them)
        ms = C.c_double()
        check(eng.handle, lib().mcp_paths(eng.handle, C.byref(p), mu.ctypes.data, sigma.ctypes.data, w.ctypes.data,
                                          terminal.data_ptr(), C.byref(ms)))
        stats = quantile_stats(terminal, alphas, device=eng.device, allreduce=allreduce, n_total=n_total)
        kernel_ms, quantile_ms = ms.value, eng.last_kernel_ms()
    else:
        st = _lib.PathStats()
        st.n_alphas, st.comm_merge = len(alphas), int(bool(comm_merge))
        st.n_total = int(n_total) if n_total is not None else M
        for i, a in enumerate(alphas):
            st.alphas[i] = a
        if _group is not None:
            for e in _group.engines:
                e.set_stream(None)
            check(eng.handle, lib().mcp_paths_stats_multi(_group.handles, _group.world, C.byref(p), mu.ctypes.data, sigma.ctypes.data,
                                                          w.ctypes.data, C.byref(st)))
        else:
            check(eng.handle, lib().mcp_paths_stats(eng.handle, C.byref(p), mu.ctypes.data, sigma.ctypes.data, w.ctypes.data,
                                                    terminal.data_ptr() if terminal is not None else None, C.byref(st)))
        stats = {a: (float(st.var[i]), float(st.cvar[i])) for i, a in enumerate(alphas)}
        kernel_ms, quantile_ms = st.kernel_ms, st.quantile_ms
    return {"stats": stats, "terminal": terminal.cpu().numpy() if (return_terminal and terminal is not None) else None,
            "terminal_device": terminal, "kernel_ms": kernel_ms, "quantile_ms": quantile_ms}


def quantile_stats(values, alphas=(0.95, 0.99), *, device=None, allreduce=None, n_total=None):
    """{alpha: (VaR, CVaR)} with app.py:258-263 conventions via the exact radix select.

    `values`: numpy array (HOST space) or CUDA torch tensor (DEVICE space), float32/float64.
    `allreduce(ptr, count, kind)`: in-place sum across ranks of a device buffer (see
    mcportfolio.dist.make_allreduce), or the string 'comm' = NCCL all-reduces issued by libmcp on the engine's own
    communicator (mcportfolio.dist.init_comm); `n_total` = global element count.
    """
    eng = get_engine(device)
    alphas = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
    if alphas.ndim != 1 or not (1 <= alphas.size <= _lib.MCP_MAX_ALPHAS):
        raise ValueError(f"between 1 and {_lib.MCP_MAX_ALPHAS} alphas are supported")
    if _is_device_tensor(values):
        import torch
        v = values.contiguous()
        code = {torch.float32: MCP_F32, torch.float64: MCP_F64}[v.dtype]
        space, n = MCP_DEVICE, v.numel()
        eng.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
    else:
        v = np.ascontiguousarray(values)
        if v.dtype not in (np.float32, np.float64):
            v = v.astype(np.float64)
        code = MCP_F32 if v.dtype == np.float32 else MCP_F64
        space, n = MCP_HOST, v.size
    var_out = np.empty(alphas.size)
    cvar_out = np.empty(alphas.size)
    cb = _lib.ALLREDUCE_FN(0)
    if isinstance(allreduce, str):
        if allreduce != "comm":
            raise ValueError("allreduce must be a callable or 'comm' (the engine's libmcp communicator)")
        cb = _lib.ALLREDUCE_COMM
    elif allreduce is not None:
        def _cb(ptr, count, kind, _user):
            try:
                allreduce(ptr, count, kind)
                return 0
            except Exception:      # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        cb = _lib.ALLREDUCE_FN(_cb)
    total = int(n_total) if n_total is not None else n
    check(eng.handle, lib().mcp_quantiles(eng.handle, _ptr(v), space, code, n, total, alphas.ctypes.data,
                                          alphas.size, var_out.ctypes.data, cvar_out.ctypes.data, cb, None))
    return {float(a): (float(var_out[i]), float(cvar_out[i])) for i, a in enumerate(alphas)}


# ------------------------------------------------------------------------------------------
# per-portfolio historical VaR / CVaR and the five "methods" of the app
# ------------------------------------------------------------------------------------------

def historical_var_cvar(returns_matrix, weights, alpha=0.95, *, dtype="float32", first_index=0,
                        return_arrays=True, device=None, negate=False, recheck=True, out=None):
    """VaR / CVaR of the historical series ``returns_matrix @ w`` for every portfolio.

    app.py:710-713 (definitions 258-263).  `weights`: (P, N) numpy array or CUDA torch tensor.
    Returns {'var': (P,), 'cvar': (P,), 'best_var': {...}, 'best_cvar': {...}, 'kernel_ms'}
    where best_var / best_cvar are the picks of the 'VaR' / 'CVaR' methods:
    ``np.argmin(-var)`` = first index of the largest VaR (app.py:673-674, 717, 747).
    negate=True: the arrays hold -var / -cvar, the app's metric (app.py:717), written by the kernel.
    recheck (float32 only): portfolios within FP32 rounding of the best are re-evaluated in FP64 on the device, so the
    pick is the index the reference's FP64 ``np.argmin`` returns.
    out: optional {'var': array, 'cvar': array} of preallocated host buffers.
    """
    R = np.ascontiguousarray(np.asarray(returns_matrix, dtype=np.float64))
    if R.ndim != 2 or R.shape[0] < 1:
        raise ValueError(f"returns_matrix must be (T, N) with T >= 1, got {R.shape}")
    T, n = R.shape
    code, npdt = _dtype(dtype)
    dev_mode = _is_device_tensor(weights)
    if dev_mode:
        import torch
        w = weights.to(torch.float32 if code == MCP_F32 else torch.float64).contiguous()
        shape = tuple(w.shape)
    else:
        w = np.ascontiguousarray(np.asarray(weights), dtype=npdt)
        shape = w.shape
    if len(shape) != 2 or shape[1] != n:
        raise ValueError(f"weights must have shape (P, {n}), got {shape}")
    P = shape[0]
    eng = get_engine(device)
    if dev_mode:
        if w.device.index != eng.device:
            raise ValueError(f"weights live on cuda:{w.device.index} but the engine runs on cuda:{eng.device}")
        eng.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
    p = _lib.HistParams()
    p.n_assets, p.n_periods, p.dtype = n, T, code
    p.space = MCP_DEVICE if dev_mode else MCP_HOST
    p.n_portfolios, p.first_index, p.alpha = P, int(first_index), float(alpha)
    p.weights_in = _ptr(w)
    p.negate, p.recheck = int(bool(negate)), int(bool(recheck) and code == MCP_F32)
    res = _lib.HistOut()
    var = cvar = None
    if return_arrays:
        if dev_mode:
            import torch
            var, cvar = torch.empty(P, dtype=w.dtype, device=w.device), torch.empty(P, dtype=w.dtype, device=w.device)
        else:
            var = out["var"] if out and out.get("var") is not None else _result_empty((P,), npdt)
            cvar = out["cvar"] if out and out.get("cvar") is not None else _result_empty((P,), npdt)
            for a in (var, cvar):
                if a.shape != (P,) or a.dtype != npdt or not a.flags.c_contiguous:
                    raise ValueError(f"out arrays must be C-contiguous ({P},) {np.dtype(npdt)}")
        res.var, res.cvar = _ptr(var), _ptr(cvar)
    check(eng.handle, lib().mcp_historical_var(eng.handle, C.byref(p), R.ctypes.data, C.byref(res)))
    pick = lambda i, v: None if i == MCP_NO_INDEX else {"index": int(i) - int(first_index), "global_index": int(i), "value": v}
    return {"var": var, "cvar": cvar, "best_var": pick(res.best_var_index, res.best_var),
            "best_cvar": pick(res.best_cvar_index, res.best_cvar), "kernel_ms": res.kernel_ms}


METHODS = ("Monte Carlo", "VaR", "CVaR", "MPT", "Equal Weight")


def estimate_moments(returns_matrix, annual_factor=12, *, device=None):
    """mu = mean(R) * A, Sigma = cov(R, ddof=1) * A on the device (app.py:679-680): FP64, one CTA per column pair.

    `returns_matrix`: (T, N) array or DataFrame (the app's `returns_df`, leading fillna(0) row included).
    Returns (mu (N,), sigma (N, N)) as FP64 numpy arrays."""
    R = np.ascontiguousarray(np.asarray(returns_matrix, dtype=np.float64))
    if R.ndim == 1:
        R = R[:, None]
    if R.ndim != 2 or R.shape[0] < 1 or R.shape[1] < 1:
        raise ValueError(f"returns_matrix must be (T, N) with T, N >= 1, got {R.shape}")
    T, n = R.shape
    mu, sigma = np.empty(n), np.empty((n, n))
    eng = get_engine(device)
    check(eng.handle, lib().mcp_moments(eng.handle, R.ctypes.data, T, n, float(annual_factor), mu.ctypes.data, sigma.ctypes.data))
    return mu, sigma


def simulate_method(returns_matrix, method="Monte Carlo", n_portfolios=2500, *, annual_factor=12,
                    risk_free=3.0, min_weights=None, max_weights=None, alpha=0.95, seed=0,
                    dtype="float32", device=None, weights=None, moments=None):
    """One pass of the app's per-method loop (app.py:682-722) and its pick (672-676, 747).

    Returns {'risks', 'returns', 'weights', 'metrics', 'opt_idx', 'opt_weights'} with the
    reference's array semantics: 'metrics' holds sharpe ('Monte Carlo', 'MPT', 'Equal Weight'),
    -var_95 ('VaR') or -cvar_95 ('CVaR') (app.py:717); opt_idx = argmax / argmin / 0.
    mu / Sigma estimation (app.py:679-680) runs on the device (`estimate_moments`; pass `moments=(mu, sigma)` to reuse
    them across the five methods).  For the 'VaR' / 'CVaR' methods the weights stay on the GPU between the sweep and the
    historical kernel, the kernel writes the NEGATED metric itself, and every array comes back with one copy into pooled
    page-locked memory.  `weights`: optional (P, N) rows evaluated instead of in-kernel draws (parity runs against the
    reference's legacy-seeded draws).
    """
    if method not in METHODS:
        raise KeyError(method)
    R = np.ascontiguousarray(np.asarray(returns_matrix, dtype=np.float64))
    T, n = R.shape
    mu, sigma = moments if moments is not None else estimate_moments(R, annual_factor, device=device)
    code, npdt = _dtype(dtype)
    if method == "Equal Weight":
        w = np.full((1, n), 1.0 / n)
        r = simulate_portfolios(mu, sigma, 1, weights=w, risk_free=risk_free, min_weights=min_weights,
                                max_weights=max_weights, dtype=dtype, device=device)
        if r.n_accepted == 0:
            raise IndexError("equal weights violate the bounds: the reference's arrays are empty (app.py:687, 747)")
        return {"risks": r.risks, "returns": r.returns, "weights": r.weights, "metrics": r.sharpes,
                "opt_idx": 0, "opt_weights": np.asarray(r.weights[0], dtype=np.float64)}
    P = int(n_portfolios) if weights is None else len(weights)
    if method not in ("VaR", "CVaR"):
        r = simulate_portfolios(mu, sigma, P, risk_free=risk_free, min_weights=min_weights, max_weights=max_weights,
                                seed=seed, dtype=dtype, device=device, weights=weights)
        if r.n_accepted == 0:
            raise ValueError("no portfolio satisfied the bounds (the reference raises at argmax of an empty array, app.py:747)")
        return {"risks": r.risks, "returns": r.returns, "weights": r.weights, "metrics": r.sharpes,
                "opt_idx": int(r.max_sharpe["index"]), "opt_weights": np.asarray(r.weights[r.max_sharpe["index"]], dtype=np.float64)}
    # ---- 'VaR' / 'CVaR': sweep with everything kept on the device, historical kernel on the same weights ----
    import torch
    eng = get_engine(device)
    r = simulate_portfolios(mu, sigma, P, risk_free=risk_free, min_weights=min_weights, max_weights=max_weights,
                            seed=seed, dtype=dtype, device=device, return_arrays="device", weights=weights)
    if r.n_accepted == 0:
        raise ValueError("no portfolio satisfied the bounds (the reference raises at argmax of an empty array, app.py:747)")
    Pa = r.n_accepted
    host = {"weights": _result_empty((Pa, n), npdt), "returns": _result_empty((Pa,), npdt), "risks": _result_empty((Pa,), npdt),
            "metrics": _result_empty((Pa,), npdt)}
    # The sweep's arrays are final: their copies back (the weights are 16 of the 19 values per portfolio) leave on a side stream
    # while the historical kernel runs on the same weights; only the metric follows the kernel.  One DMA each, into (pooled)
    # page-locked memory.
    main = torch.cuda.current_stream(eng.device)
    side = eng.copy_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        for name, src in (("weights", r.weights), ("returns", r.returns), ("risks", r.risks)):
            torch.from_numpy(host[name]).copy_(src, non_blocking=True)
    hv = historical_var_cvar(R, r.weights, alpha, dtype=dtype, device=device, negate=True)
    key, pick = ("var", "best_var") if method == "VaR" else ("cvar", "best_cvar")
    torch.from_numpy(host["metrics"]).copy_(hv[key], non_blocking=True)
    main.synchronize()
    side.synchronize()
    opt = int(hv[pick]["index"])
    return {"risks": host["risks"], "returns": host["returns"], "weights": host["weights"], "metrics": host["metrics"],
            "opt_idx": opt, "opt_weights": np.asarray(host["weights"][opt], dtype=np.float64)}


STATS_FIELDS = ("sharpe", "sortino", "volatility_ann", "total_return_ann", "mean_ann", "mean_month",
                "std_month", "min_month", "max_month", "max_drawdown", "var_95", "cvar_95")


def asset_stats(returns_matrix, *, annual_factor=12, risk_free=0.0, alpha=0.95, device=None):
    """Per-asset statistics of a (T, N) returns matrix: `calc_asset_stats` (app.py:286-335).

    Returns a list of N dicts with the reference's keys (plus min_ann / max_ann / std_ann /
    implied_vol, which app.py:301-310 derives from the same numbers)."""
    R = np.ascontiguousarray(np.asarray(returns_matrix, dtype=np.float64))
    if R.ndim == 1:
        R = R[:, None]
    if R.ndim != 2 or R.shape[0] < 1:
        raise ValueError(f"returns_matrix must be (T, N) with T >= 1, got {R.shape}")
    T, n = R.shape
    out = np.empty((n, len(STATS_FIELDS)))
    eng = get_engine(device)
    check(eng.handle, lib().mcp_asset_stats(eng.handle, R.ctypes.data, T, n, float(risk_free), float(annual_factor),
                                            float(alpha), out.ctypes.data))
    res = []
    for row in out:
        d = dict(zip(STATS_FIELDS, (float(v) for v in row)))
        d["implied_vol"] = d["std_ann"] = d["volatility_ann"]
        d["min_ann"], d["max_ann"] = d["min_month"] * annual_factor, d["max_month"] * annual_factor
        res.append(d)
    return res
