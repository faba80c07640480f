// mcp_historical_var: per-portfolio historical VaR / CVaR (app.py:710-713) and the 'VaR' /
// 'CVaR' method selections (app.py:673-674).  SURVEY.md 8(f) row f1 -- scheduled after the
// core path; fails loudly until the kernel lands (no CPU fallback).
#include "mcp_context.h"

extern "C" int mcp_historical_var(mcp_handle h, const mcp_hist_params* params, const double* returns_matrix_host,
                                  mcp_hist_out* out) {
    if (!h) return MCP_ERR_INVALID;
    (void)params; (void)returns_matrix_host; (void)out;
    return mcp_fail(h, MCP_ERR_INVALID, "mcp_historical_var: kernel not built yet");
}
