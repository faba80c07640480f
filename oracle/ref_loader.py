"""Run the reference's OWN lines for the path, in this container only.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``import app`` is impossible
(streamlit / plotly / yfinance are absent and the module body is the UI), but
the path's functions are pure numpy: the slice ``app.py:231-284``
(``sharpe_ratio`` ... ``efficient_frontier``, incl. ``var`` and ``cvar``) is
``exec``-ed into a namespace holding only ``np`` and ``pd``.  Nothing is copied
into the repo; the outputs are recorded by ``oracle/make_golden.py``.

``/root/reference`` does not exist on the GPU box: only ``make_golden.py`` and
the ``needs_reference`` CPU tests call this.
"""
from __future__ import annotations

import os

REFERENCE_APP = os.environ.get("MCP_REFERENCE_APP", "/root/reference/app.py")
REFERENCE_DATA = os.path.join(os.path.dirname(REFERENCE_APP), "data")
_SLICE = (231, 284)            # 1-based inclusive line range of the stats + frontier functions


def available() -> bool:
    return os.path.isfile(REFERENCE_APP)


def load_reference_functions():
    """Namespace with the reference's var / cvar / efficient_frontier / max_drawdown ..."""
    import numpy as np
    import pandas as pd
    with open(REFERENCE_APP, encoding="utf-8") as fh:
        lines = fh.readlines()
    src = "".join(lines[_SLICE[0] - 1:_SLICE[1]])
    ns = {"np": np, "pd": pd}
    exec(compile(src, REFERENCE_APP, "exec"), ns)
    for name in ("var", "cvar", "efficient_frontier", "max_drawdown", "annual_return"):
        assert callable(ns.get(name)), f"reference slice no longer defines {name}"
    return ns
