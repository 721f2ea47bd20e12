// Stand-alone bring-up test of the tcgen05 building blocks used by the tensor-core GRU kernel:
// D[128 x N] (fp32, TMEM) = A[128 x K] (fp16) * B[N x K]^T (fp16), K-major operands without swizzle.
//   mode 0: A from shared memory (SS)      mode 1: A from tensor memory (TS, written with tcgen05.st)
// Prints the max abs error against a CPU double-precision product.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no-swizzle canonical layout: 8-row x 16-byte core matrices; LBO = stride between the core matrices
// adjacent in K, SBO = stride between 8-row groups (bytes)
__host__ __device__ inline uint32_t kmajor_off(int r, int k, int K) {
    return (uint32_t)((r >> 3) * ((K >> 3) * 128) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N, int K>
__global__ void __launch_bounds__(128, 1) gemm_test(const __half *A, const __half *B, float *D, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar;
    unsigned char *sA = smem;                       // 128 x K fp16
    unsigned char *sB = smem + 128 * K * 2;         // N x K fp16
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half *>(sA + kmajor_off(r, k, K)) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half *>(sB + kmajor_off(r, k, K)) = B[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t d_col = 0, a_col = 256;
    if (mode == 1) {
        // A into tensor memory: lane = row, 32-bit column c holds elements (2c, 2c+1)
        const int r = tid;
#pragma unroll 1
        for (int c = 0; c < K / 2; c += 8) {
            uint32_t v[8];
            for (int q = 0; q < 8; ++q) {
                const __half2 h2 = __halves2half2(A[r * K + 2 * (c + q)], A[r * K + 2 * (c + q) + 1]);
                v[q] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(tmem + lane_base + a_col + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                         "r"(v[6]), "r"(v[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint32_t sbo = (K / 8) * 128, lbo = 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t bdesc = make_desc(smem_u32(sB) + ks * 2 * lbo, lbo, sbo);
            const uint32_t acc = ks > 0;
            if (mode == 0) {
                const uint64_t adesc = make_desc(smem_u32(sA) + ks * 2 * lbo, lbo, sbo);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + d_col), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem + d_col), "r"(tmem + a_col + ks * 8), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(&s_bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + lane_base + d_col + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 8; ++q) D[tid * N + c + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int N, int K>
static double run(int mode) {
    std::vector<__half> hA(128 * K), hB(N * K);
    std::vector<float> fA(128 * K), fB(N * K);
    srand(1 + N + K);
    for (int i = 0; i < 128 * K; ++i) { fA[i] = (rand() % 2001 - 1000) / 1000.0f; hA[i] = __float2half(fA[i]); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { fB[i] = (rand() % 2001 - 1000) / 1000.0f; hB[i] = __float2half(fB[i]); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float *dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * N * 4);
    const size_t smem = (128 + N) * K * 2;
    cudaFuncSetAttribute(gemm_test<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gemm_test<N, K><<<1, 128, smem>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
    std::vector<float> hD(128 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)fA[m * K + k] * fB[n * K + k];
            const double err = fabs(acc - hD[m * N + n]);
            if (err > maxerr) maxerr = err;
        }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr;
}

int main() {
    printf("SS N=64  K=64 : max err %.3e\n", run<64, 64>(0));
    printf("SS N=192 K=16 : max err %.3e\n", run<192, 16>(0));
    printf("SS N=128 K=64 : max err %.3e\n", run<128, 64>(0));
    printf("SS N=16  K=64 : max err %.3e\n", run<16, 64>(0));
    printf("TS N=64  K=64 : max err %.3e\n", run<64, 64>(1));
    printf("TS N=192 K=16 : max err %.3e\n", run<192, 16>(1));
    return 0;
}
