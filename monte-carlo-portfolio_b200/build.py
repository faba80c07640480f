"""Build libmcp.so (sm_100a only) in-tree with explicit nvcc commands.

    python monte-carlo-portfolio_b200/build.py [--force] [--verbose]

Objects go to ``monte-carlo-portfolio_b200/build/``, the library to
``monte-carlo-portfolio_b200/lib/libmcp.so`` (git-ignored, but it travels to the GPU box
with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmcp.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE,
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]

SMALL_NP = (4, 8, 16, 24, 32)


def translation_units():
    """(object name, source, extra defines)"""
    tus = [("mcp_context", "mcp_context.cu", []),
           ("mcp_comm", "mcp_comm.cu", []),
           ("mcp_multi", "mcp_multi.cu", []),
           ("mcp_portfolio", "mcp_portfolio.cu", []),
           ("mcp_portfolio_large", "mcp_portfolio_large.cu", []),
           ("mcp_portfolio_large_tc", "mcp_portfolio_large_tc.cu", []),
           ("mcp_envelope", "mcp_envelope.cu", []),
           ("mcp_recheck", "mcp_recheck.cu", []),
           ("mcp_paths", "mcp_paths.cu", []),
           ("mcp_paths_tc", "mcp_paths_tc.cu", []),
           ("mcp_paths_tc16", "mcp_paths_tc16.cu", []),
           ("mcp_quantile", "mcp_quantile.cu", []),
           ("mcp_historical", "mcp_historical.cu", []),
           ("mcp_stats", "mcp_stats.cu", [])]
    for t, tag in (("float", "f32"), ("double", "f64")):
        for np_ in SMALL_NP:
            tus.append((f"mcp_small_{tag}_{np_}", "mcp_portfolio_small_inst.cu",
                        [f"-DMCP_INST_T={t}", f"-DMCP_INST_NP={np_}"]))
    return tus


def _deps_digest():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(a for a in ARCH + CFLAGS if a != INCLUDE).encode())     # path-independent: the tree moves
    return h.hexdigest()


def _compile(tu, verbose):
    name, src, defs = tu
    obj = os.path.join(BUILD, name + ".o")
    cmd = [NVCC, *ARCH, *CFLAGS, *defs, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(BUILD, name + ".log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    import fcntl
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(BUILD, ".lock"), "w") as lock:        # one builder at a time (torchrun ranks)
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build_locked(force, verbose)


def _build_locked(force: bool, verbose: bool) -> str:
    stamp = os.path.join(BUILD, "stamp")
    digest = _deps_digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read() == digest:
        return LIB
    tus = translation_units()
    with cf.ThreadPoolExecutor(max_workers=min(len(tus), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda t: _compile(t, verbose), tus))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl", "-lpthread", "-Xlinker", "--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib, os.path.getsize(lib), "bytes")
