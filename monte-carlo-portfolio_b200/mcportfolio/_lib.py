"""ctypes binding of libmcp.so (the C ABI declared in include/mcp.h).

There is no CPU fallback: if the library is missing or no B200 is present the
first call raises.  ``build()`` compiles it in-tree (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libmcp.so")

MCP_OK, MCP_ERR_INVALID, MCP_ERR_CUDA, MCP_ERR_NUMERIC, MCP_ERR_NOMEM, MCP_ERR_COMM = 0, -1, -2, -3, -4, -5
MCP_ABI_VERSION = 2
MCP_F32, MCP_F64 = 0, 1
MCP_HOST, MCP_DEVICE = 0, 1
MCP_NO_INDEX = 0xFFFFFFFFFFFFFFFF
MCP_MAX_ALPHAS = 8
MCP_MAX_TARGETS = 16
MCP_COMM_ID_BYTES = 128
MCP_REDUCE_U64_SUM, MCP_REDUCE_F64_SUM, MCP_REDUCE_F64_MIN, MCP_REDUCE_F64_MAX, MCP_REDUCE_U64_MAX = 0, 1, 2, 3, 4


class McpError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libmcp error {code}: {message}")
        self.code = code


class DeviceInfo(C.Structure):
    _fields_ = [("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("max_smem_per_block", C.c_int32), ("total_mem", C.c_uint64), ("name", C.c_char * 64)]


class PortfolioParams(C.Structure):
    _fields_ = [("n_assets", C.c_int32), ("dtype", C.c_int32),
                ("n_portfolios", C.c_uint64), ("first_index", C.c_uint64), ("seed", C.c_uint64),
                ("risk_free", C.c_double), ("risk_target", C.c_double),
                ("min_weights", C.c_void_p), ("max_weights", C.c_void_p),
                ("max_tries", C.c_int32), ("keep_last", C.c_int32),
                ("space", C.c_int32), ("comm_merge", C.c_int32),
                ("weights_in", C.c_void_p), ("weights_recheck", C.c_void_p),
                ("n_bins", C.c_int32), ("philox_rounds", C.c_int32),
                ("risk_lo", C.c_double), ("risk_hi", C.c_double)]


class Selection(C.Structure):
    _fields_ = [("index", C.c_uint64), ("key", C.c_double), ("ret", C.c_double),
                ("risk", C.c_double), ("sharpe", C.c_double), ("weights", C.c_void_p)]


class PortfolioOut(C.Structure):
    _fields_ = [("weights", C.c_void_p), ("returns", C.c_void_p), ("risks", C.c_void_p),
                ("sharpes", C.c_void_p), ("accepted", C.c_void_p),
                ("bin_best_return", C.c_void_p), ("bin_best_index", C.c_void_p),
                ("n_accepted", C.c_uint64), ("risk_min", C.c_double), ("risk_max", C.c_double),
                ("max_sharpe", Selection), ("target_risk", Selection), ("kernel_ms", C.c_double),
                ("n_accepted_global", C.c_uint64), ("recheck_overflow", C.c_int32), ("reserved", C.c_int32)]


class PathParams(C.Structure):
    _fields_ = [("n_assets", C.c_int32), ("dtype", C.c_int32),
                ("n_paths", C.c_uint64), ("first_index", C.c_uint64), ("seed", C.c_uint64),
                ("n_steps", C.c_int32), ("space", C.c_int32), ("dt", C.c_double),
                ("normals_in", C.c_void_p), ("philox_rounds", C.c_int32), ("reserved", C.c_int32)]


class PathStats(C.Structure):
    _fields_ = [("n_alphas", C.c_int32), ("comm_merge", C.c_int32), ("n_total", C.c_uint64),
                ("alphas", C.c_double * MCP_MAX_ALPHAS), ("var", C.c_double * MCP_MAX_ALPHAS),
                ("cvar", C.c_double * MCP_MAX_ALPHAS), ("kernel_ms", C.c_double), ("quantile_ms", C.c_double)]


class SelectState(C.Structure):
    _fields_ = [("key_bits", C.c_int32), ("bits_done", C.c_int32), ("n_targets", C.c_int32),
                ("n_slots", C.c_int32), ("rank", C.c_uint64 * MCP_MAX_TARGETS),
                ("prefix", C.c_uint64 * MCP_MAX_TARGETS), ("slot_of", C.c_int32 * MCP_MAX_TARGETS),
                ("slot_prefix", C.c_uint64 * MCP_MAX_TARGETS)]


class HistParams(C.Structure):
    _fields_ = [("n_assets", C.c_int32), ("n_periods", C.c_int32), ("dtype", C.c_int32),
                ("space", C.c_int32), ("n_portfolios", C.c_uint64), ("first_index", C.c_uint64),
                ("alpha", C.c_double), ("weights_in", C.c_void_p), ("negate", C.c_int32), ("recheck", C.c_int32)]


class HistOut(C.Structure):
    _fields_ = [("var", C.c_void_p), ("cvar", C.c_void_p),
                ("best_var_index", C.c_uint64), ("best_cvar_index", C.c_uint64),
                ("best_var", C.c_double), ("best_cvar", C.c_double), ("kernel_ms", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)
ALLREDUCE_COMM = C.cast(C.c_void_p(1), ALLREDUCE_FN)       # MCP_ALLREDUCE_COMM: sum over the handle's own NCCL communicator

# every symbol include/mcp.h declares: (restype, argtypes)
SYMBOLS = {
    "mcp_abi_version": (C.c_int, []),
    "mcp_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mcp_destroy": (C.c_int, [C.c_void_p]),
    "mcp_last_error": (C.c_char_p, [C.c_void_p]),
    "mcp_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mcp_synchronize": (C.c_int, [C.c_void_p]),
    "mcp_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "mcp_host_free": (C.c_int, [C.c_void_p]),
    "mcp_device_info": (C.c_int, [C.c_void_p, C.POINTER(DeviceInfo)]),
    "mcp_launch_count": (C.c_uint64, [C.c_void_p]),
    "mcp_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "mcp_portfolios": (C.c_int, [C.c_void_p, C.POINTER(PortfolioParams), C.c_void_p, C.c_void_p,
                                 C.POINTER(PortfolioOut)]),
    "mcp_paths": (C.c_int, [C.c_void_p, C.POINTER(PathParams), C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.POINTER(C.c_double)]),
    "mcp_quantiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, ALLREDUCE_FN, C.c_void_p]),
    "mcp_select_init": (C.c_int, [C.POINTER(SelectState), C.c_int, C.c_void_p, C.c_int]),
    "mcp_select_pass_bits": (C.c_int, [C.POINTER(SelectState)]),
    "mcp_select_advance": (C.c_int, [C.POINTER(SelectState), C.c_void_p]),
    "mcp_select_hist": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.POINTER(SelectState),
                                  C.c_void_p]),
    "mcp_key_to_value": (C.c_double, [C.c_uint64, C.c_int]),
    "mcp_historical_var": (C.c_int, [C.c_void_p, C.POINTER(HistParams), C.c_void_p, C.POINTER(HistOut)]),
    "mcp_asset_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "mcp_measure_fma_peak": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "mcp_set_allreduce_stream_ordered": (C.c_int, [C.c_void_p, C.c_int]),
    "mcp_envelope_arrays": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.c_double,
                                      C.c_int, C.c_void_p, C.c_void_p]),
    "mcp_paths_stats": (C.c_int, [C.c_void_p, C.POINTER(PathParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(PathStats)]),
    "mcp_moments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "mcp_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mcp_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mcp_comm_init_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "mcp_comm_destroy": (C.c_int, [C.c_void_p]),
    "mcp_comm_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mcp_comm_nccl_version": (C.c_int, [C.POINTER(C.c_int)]),
    "mcp_comm_allgather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mcp_comm_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "mcp_portfolios_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(PortfolioParams), C.c_void_p, C.c_void_p,
                                       C.POINTER(PortfolioOut), C.c_void_p]),
    "mcp_paths_stats_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(PathParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(PathStats)]),
}

_lib = None


def build(force: bool = False) -> str:
    """Compile libmcp.so in-tree (no GPU needed)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mcp_build", os.path.join(PKG_ROOT, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force)


def lib():
    """The loaded library with prototypes set.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise McpError(MCP_ERR_INVALID,
                           f"{LIB_PATH} not found: run `python monte-carlo-portfolio_b200/build.py` "
                           "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError here = ABI drift between header and binary
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(handle, rc):
    if rc != MCP_OK:
        msg = lib().mcp_last_error(handle)
        raise McpError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")
