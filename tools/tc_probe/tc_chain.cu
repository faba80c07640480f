// tcgen05 issue / latency probe for NARROW MMAs (development tool, not part of the library):
// how long does a chain of dependent 128 x N x 8 TF32 MMAs (same accumulator) take per MMA, and how much of that is
// latency that independent accumulators (other row tiles) can hide when their MMAs are interleaved in issue order?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_chain tc_chain.cu && ./tc_chain
// Output: cycles per MMA (issue -> completion of the whole batch) for N in {16, 32}, kind in {tf32 K=8, f16 K=16},
// NT interleaved accumulators in {1, 2, 4, 7}.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mma_tf32(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ inline uint32_t make_idesc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ inline uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}

// `chain` dependent MMAs per accumulator, NT accumulators interleaved in issue order (mma c of every accumulator, then mma c + 1
// of every accumulator ...), each round of NT MMAs ONE asm block (no per-MMA issue code).  A = whatever is in TMEM, B = zeros.
#define MMA_LINE(D) "tcgen05.mma.cta_group::1.kind::tf32 [" D "], [%7], %8, %9, p;\n\t"
template <int NT>
__device__ __forceinline__ void round_block(const uint32_t* d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if constexpr (NT == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %10, 0;\n\t" MMA_LINE("%0") "}" ::"r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]),
                     "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else if constexpr (NT == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %10, 0;\n\t" MMA_LINE("%0") MMA_LINE("%1") "}" ::"r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]),
                     "r"(d[6]), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else if constexpr (NT == 4)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %10, 0;\n\t" MMA_LINE("%0") MMA_LINE("%1") MMA_LINE("%2") MMA_LINE("%3") "}" ::"r"(d[0]), "r"(d[1]), "r"(d[2]),
                     "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %10, 0;\n\t" MMA_LINE("%0") MMA_LINE("%1") MMA_LINE("%2") MMA_LINE("%3") MMA_LINE("%4") MMA_LINE("%5") MMA_LINE("%6") "}"
                     ::"r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(160, 1) chain_probe(int N, int chain, int reps, long long* timing) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 32 * N; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    if (warp == 4 && lane == 0) {
        const uint32_t lbo = (uint32_t)(N / 8) * 128, sbo = 128;
        const uint64_t bd = make_sdesc(smem_u32(smem), lbo, sbo);
        const uint32_t idesc = make_idesc(2, 128, N);
        uint32_t d[7];
        for (int t = 0; t < 7; ++t) d[t] = tb + 64u * t;
        const uint32_t a = tb + 448u;
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            for (int c = 0; c < chain; ++c) round_block<NT>(d, a, bd, idesc, c > 0);
            const long long t1 = clock64();
            tc_commit(&bar_done);
            mbar_wait(&bar_done, phase);
            phase ^= 1u;
            const long long t2 = clock64();
            if (r == reps - 1) { timing[0] = t1 - t0; timing[1] = t2 - t0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int NT>
static void run(int N, int chain, long long* dT) {
    CK(cudaFuncSetAttribute(chain_probe<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 256 * 4));
    chain_probe<NT><<<1, 160, 32 * N * 4>>>(N, chain, 8, dT);
    CK(cudaDeviceSynchronize());
    long long T[2]; CK(cudaMemcpy(T, dT, 16, cudaMemcpyDeviceToHost));
    printf("tf32 N=%3d chain=%2d nt=%d: issue %6lld cyc, to completion %6lld cyc = %6.1f cyc per MMA (%7.1f per tile-chain)\n", N, chain, NT, T[0], T[1],
           (double)T[1] / (chain * NT), (double)T[1] / NT);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sm_%d%d\n", prop.name, prop.major, prop.minor);
    long long* dT; CK(cudaMalloc(&dT, 16));
    for (int N : {16, 64, 256})
        for (int chain : {1, 7, 28, 112}) {
            run<1>(N, chain, dT);
            run<2>(N, chain, dT);
            run<4>(N, chain, dT);
            run<7>(N, chain, dT);
        }
    return 0;
}
