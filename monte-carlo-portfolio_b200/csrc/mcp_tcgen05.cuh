// tcgen05 / TMEM / mbarrier PTX wrappers shared by the tensor-core kernels (mcp_portfolio_large_tc.cu: the N <= 256 portfolio
// sweep; mcp_paths_tc.cu: the correlated-path simulator).  sm_100a only.  Conventions pinned on hardware by tools/tc_probe:
// A operand in tensor memory (written with tcgen05.st, 32x32b shape: thread t of warp w owns lane 32 (w % 4) + t), B operand in
// shared memory in the canonical K-major no-swizzle core-matrix layout, FP32 accumulators in tensor memory, cta_group::1.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcp {

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t addr = smem_u32(b);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// warp-convergent issue: every lane executes the block, one elected lane issues (no divergent-branch lane loop around UTCHMMA)
__device__ __forceinline__ void mma_tf32_ts_elect(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16_ts_elect(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// One chunk of the FP16 split under a single elect: h1 S1 + h2 S1 + h1 S2 for both K = 16 halves, then the commit.  dlo* are the low
// words of the four shared-memory descriptors (start address and LBO), dhi their common high word (SBO, version).
__device__ __forceinline__ void mma6_f16_commit_elect(uint32_t d, uint32_t a_h1, uint32_t a_h2, uint32_t dlo10, uint32_t dlo11, uint32_t dlo20,
                                                      uint32_t dlo21, uint32_t dhi, uint32_t idesc, uint32_t acc, uint32_t bar) {
    asm volatile("{\n\t.reg .pred p, q, t;\n\t.reg .b64 e10, e11, e20, e21;\n\t.reg .b32 b1, b2;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "setp.ne.b32 p, %9, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                 "mov.b64 e10, {%3, %7};\n\tmov.b64 e11, {%4, %7};\n\tmov.b64 e20, {%5, %7};\n\tmov.b64 e21, {%6, %7};\n\t"
                 "add.u32 b1, %1, 8;\n\tadd.u32 b2, %2, 8;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], e10, %8, p;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], e10, %8, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], e20, %8, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [b1], e11, %8, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [b2], e11, %8, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [b1], e21, %8, t;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t}"
                 ::"r"(d), "r"(a_h1), "r"(a_h2), "r"(dlo10), "r"(dlo11), "r"(dlo20), "r"(dlo21), "r"(dhi), "r"(idesc), "r"(acc), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* b) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(b)) : "memory");
}
#define TC_R32(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),   \
                  "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),     \
                  "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
                 "%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr), TC_R32(v) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
                 "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
                 "%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
// instruction descriptor (cute::UMMA::InstrDescriptor, cute/arch/mma_sm100_desc.hpp): c_format [4,6) = 1 (F32); a_format [7,10) and
// b_format [10,13): 1 = BF16, 2 = TF32; a/b major bits 15/16 = 0 (K-major); n_dim [17,23) = N >> 3; m_dim [24,29) = M >> 4
__device__ __forceinline__ uint32_t tc_idesc(uint32_t fmt, uint32_t N) { return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24); }
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, no swizzle: start >> 4 [0,14); LBO >> 4 [16,30) = byte distance
// between the two 8x16-byte core matrices along K; SBO >> 4 [32,46) = distance between 8-row groups along N; version 1 at [46,48)
__device__ __forceinline__ uint64_t tc_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

}  // namespace mcp
