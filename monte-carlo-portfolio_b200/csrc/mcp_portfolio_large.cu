// Block-cooperative sweep for N > 32 assets (C5: N = 256).  Placeholder until the tiled
// kernel lands: fails loudly instead of falling back to anything slower.
#include "mcp_portfolio.h"

namespace mcp {

int pf_large_launch(mcp_context* h, PfJob& job) {
    return mcp_fail(h, MCP_ERR_INVALID, "n_assets=%d: the large-N sweep kernel is not built yet (N <= %d supported)",
                    job.n, PF_SMALL_MAX_N);
}

int pf_large_replay(mcp_context* h, const PfJob& job, const PfReplay&) {
    return mcp_fail(h, MCP_ERR_INVALID, "n_assets=%d: the large-N sweep kernel is not built yet", job.n);
}

}  // namespace mcp
