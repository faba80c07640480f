// One (arithmetic type, padded N) instantiation of the register-resident sweep per
// translation unit, so the variants compile in parallel:
//   nvcc ... -DMCP_INST_T=float -DMCP_INST_NP=16 -c mcp_portfolio_small_inst.cu
#include "mcp_portfolio_small.cuh"

#if !defined(MCP_INST_T) || !defined(MCP_INST_NP)
#error "define MCP_INST_T (float|double) and MCP_INST_NP (4|8|16|24|32)"
#endif

namespace mcp {
template int pf_small_launch_t<MCP_INST_T, MCP_INST_NP>(mcp_context*, PfJob&);
template int pf_small_replay_t<MCP_INST_T, MCP_INST_NP>(mcp_context*, const PfJob&, const PfReplay&);
}  // namespace mcp
