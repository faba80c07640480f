// tcgen05 convention probe (development tool, not part of the library):
//   D[128 x N] (TMEM, fp32) = A[128 x 32] (TMEM, written with tcgen05.st) x B[32 x N] (smem, K-major,
//   no-swizzle canonical layout), as tf32 hi/lo + bf16 correction -- the arithmetic the N<=256 sweep uses.
// Prints max errors per mode so a wrong descriptor / layout convention is visible at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu && ./tc_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }

__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]),
                 "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6) 1 = F32; a_format [7,10), b_format [10,13): 0 F16, 1 BF16, 2 TF32;
// a_major bit 15, b_major bit 16 (0 = K-major); n_dim [17,23) = N >> 3; m_dim [24,29) = M >> 4
__host__ __device__ inline uint32_t make_idesc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// smem matrix descriptor (cute::UMMA::SmemDescriptor), no swizzle: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48)
__device__ inline uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}

constexpr int KC = 32;   // K per chunk

// mode bit 0: tf32 hi x Bhi; bit 1: tf32 lo x Bhi; bit 2: bf16(a) x Blo
// out[row][col] row-major fp32; timing[0] = cycles for `reps` repetitions of the chunk's MMAs (commit to completion)
__global__ void __launch_bounds__(160, 1) probe(const float* A, const float* Bhi, const __nv_bfloat16* Blo, float* out, int N, int mode, int reps,
                                                long long* timing) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sBhi = reinterpret_cast<float*>(smem);                                   // [8 kcore][N/8][8][4] floats
    __nv_bfloat16* sBlo = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)KC * N * 4);  // [4 kcore][N/8][8][8] bf16
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // B[k][n] (global, row-major K x N) -> canonical K-major no-swizzle: core matrix = 8 n-rows x 16 bytes of k
    for (int i = tid; i < KC * N; i += blockDim.x) {
        const int k = i / N, n = i % N;
        sBhi[((size_t)(k / 4) * (N / 8) + n / 8) * 32 + (n % 8) * 4 + (k % 4)] = Bhi[i];
        sBlo[((size_t)(k / 8) * (N / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8)] = Blo[i];
    }
    if (tid == 0) { mbar_init(&bar_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t colD = 0, colAhi = 256, colAlo = 288, colAbf = 320;

    if (warp < 4) {
        // thread = row; write A hi / lo (tf32 words) and packed bf16 pairs into TMEM
        const int row = tid;
        const uint32_t lane_addr = tbase + ((uint32_t)(32 * warp) << 16);
        for (int k0 = 0; k0 < KC; k0 += 8) {
            uint32_t hi[8], lo[8];
            for (int j = 0; j < 8; ++j) {
                const float a = A[row * KC + k0 + j];
                const uint32_t h = __float_as_uint(a) & 0xffffe000u;
                hi[j] = h;
                lo[j] = __float_as_uint(a - __uint_as_float(h));
            }
            tmem_st8(lane_addr + colAhi + k0, hi);
            tmem_st8(lane_addr + colAlo + k0, lo);
        }
        for (int k0 = 0; k0 < KC; k0 += 16) {
            uint32_t bf[8];
            for (int j = 0; j < 8; ++j) {
                const __nv_bfloat162 p = __floats2bfloat162_rn(A[row * KC + k0 + 2 * j], A[row * KC + k0 + 2 * j + 1]);   // .x = low half = even k
                bf[j] = *reinterpret_cast<const uint32_t*>(&p);
            }
            tmem_st8(lane_addr + colAbf + k0 / 2, bf);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 4) {
        const uint32_t lbo_hi = (uint32_t)(N / 8) * 128, lbo_lo = (uint32_t)(N / 8) * 128, sbo = 128;
        const uint32_t id32 = make_idesc(2, 128, N), id16 = make_idesc(1, 128, N);
        long long t0 = 0, t1 = 0;
        if (lane == 0) {
            // descriptors precomputed: the timed loop is only the ten MMAs (issue-rate / dependent-accumulate floor)
            uint64_t bdh[4], bdl[2];
            for (int ks = 0; ks < 4; ++ks) bdh[ks] = make_sdesc(smem_u32(sBhi) + ks * 2 * lbo_hi, lbo_hi, sbo);
            for (int ks = 0; ks < 2; ++ks) bdl[ks] = make_sdesc(smem_u32(sBlo) + ks * 2 * lbo_lo, lbo_lo, sbo);
            const uint32_t m1 = mode & 1, m2 = mode & 2, m4 = mode & 4;
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                uint32_t acc = 0;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (m1) { mma_tf32_ts(tbase + colD, tbase + colAhi + 8 * ks, bdh[ks], id32, acc); acc = 1; }
                    if (m2) { mma_tf32_ts(tbase + colD, tbase + colAlo + 8 * ks, bdh[ks], id32, acc); acc = 1; }
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    if (m4) { mma_f16_ts(tbase + colD, tbase + colAbf + 8 * ks, bdl[ks], id16, acc); acc = 1; }
                }
            }
            const long long t_issue = clock64();
            timing[2] = t_issue - t0;
            tc_commit(&bar_done);
        }
        __syncwarp();
        mbar_wait(&bar_done, 0);
        if (lane == 0) { t1 = clock64(); timing[0] = t1 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 4) {
        const int row = tid;
        const uint32_t lane_addr = tbase + ((uint32_t)(32 * warp) << 16);
        long long t0 = clock64();
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(lane_addr + colD + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) out[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
        }
        long long t1 = clock64();
        if (tid == 0) timing[1] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000fffu + ((u >> 13) & 1u); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }

int main() {
    int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s sm_%d%d\n", prop.name, prop.major, prop.minor);
    for (int N : {32, 64, 160, 256}) {
        std::vector<float> A(128 * KC), B(KC * N), Bhi(KC * N);
        std::vector<__nv_bfloat16> Blo(KC * N);
        std::vector<float> Blo_f(KC * N);
        srand(N);
        for (auto& a : A) a = -23.f * (float)rand() / RAND_MAX;
        for (int i = 0; i < KC * N; ++i) {
            B[i] = ((float)rand() / RAND_MAX - 0.3f) * 0.4f;
            Bhi[i] = tf32_rn(B[i]);
            Blo[i] = __float2bfloat16(B[i] - Bhi[i]);
            Blo_f[i] = __bfloat162float(Blo[i]);
        }
        float *dA, *dBhi, *dOut; __nv_bfloat16* dBlo; long long* dT;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBhi, B.size() * 4)); CK(cudaMalloc(&dBlo, B.size() * 2));
        CK(cudaMalloc(&dOut, 128 * N * 4)); CK(cudaMalloc(&dT, 32));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBhi, Bhi.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBlo, Blo.data(), B.size() * 2, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)KC * N * 6;
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int mode : {1, 2, 4, 7}) {
            for (int reps : {1, 64}) {
                CK(cudaMemset(dOut, 0xff, 128 * N * 4));
                probe<<<1, 160, smem>>>(dA, dBhi, dBlo, dOut, N, mode, reps, dT);
                CK(cudaDeviceSynchronize());
                std::vector<float> out(128 * N); long long T[3];
                CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(T, dT, 24, cudaMemcpyDeviceToHost));
                // references: what this mode should compute (exact, fp64) and the full product
                double max_err_mode = 0, max_err_full = 0, max_ref = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < N; ++n) {
                        double want = 0, full = 0;
                        for (int k = 0; k < KC; ++k) {
                            const float a = A[m * KC + k];
                            uint32_t u; memcpy(&u, &a, 4); u &= 0xffffe000u; float ah; memcpy(&ah, &u, 4);
                            const float al = a - ah;
                            uint32_t ul; memcpy(&ul, &al, 4); ul &= 0xffffe000u; float alt; memcpy(&alt, &ul, 4);   // lo as the tensor core sees it (truncated)
                            const float abf = __bfloat162float(__float2bfloat16(a));
                            if (mode & 1) want += (double)ah * Bhi[k * N + n];
                            if (mode & 2) want += (double)alt * Bhi[k * N + n];
                            if (mode & 4) want += (double)abf * Blo_f[k * N + n];
                            full += (double)a * B[k * N + n];
                        }
                        const double got = out[m * N + n];
                        const double exp_val = want;   // every repetition restarts with accumulate = 0
                        max_err_mode = fmax(max_err_mode, fabs(got - exp_val));
                        max_err_full = fmax(max_err_full, fabs(got - full));
                        max_ref = fmax(max_ref, fabs(full));
                    }
                printf("N=%3d mode=%d reps=%2d  max|D-mode_ref|=%.3e  max|D-full|=%.3e  mma_cycles(to completion)=%lld  issue_cycles=%lld\n", N, mode, reps,
                       max_err_mode, max_err_full, T[0], T[2]);
            }
        }
        cudaFree(dA); cudaFree(dBhi); cudaFree(dBlo); cudaFree(dOut); cudaFree(dT);
    }
    return 0;
}
