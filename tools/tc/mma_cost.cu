// Micro-benchmark: what does one tcgen05.mma (M = 128, K = 16, kind::f16) cost as a function of N, of where the A operand
// lives (tensor memory / shared memory), and of whether consecutive MMAs accumulate into the SAME tensor-memory columns
// (a dependent chain) or rotate over independent accumulators?  One CTA per SM, one issuing thread; cycles per MMA from
// clock64 around {issue L MMAs; commit; wait}.  Garbage operands (timing only).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t idesc(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

template <int N, int SS, int NACC>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long *out) {
    const int L = 12;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0), bar = smem_u32(&s_bar), sm0 = smem_u32(smem);
    if (warp == 0) {
        uint32_t phase = 0;
        long long best = 1LL << 60, best_issue = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const uint32_t d = tmem + (uint32_t)((i & (NACC - 1)) * N);
                const uint64_t b = make_desc(sm0 + 16384 + 256 * (i & 3), 1024);
                const uint32_t acc = i >= NACC;
                if (SS) {
                    const uint64_t a = make_desc(sm0 + 256 * (i & 3), 1024);
                    asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(a), "l"(b), "r"(idesc(N)), "r"(acc) : "memory");
                } else {
                    asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(d), "r"(tmem + 448 + 8 * (i & 3)), "l"(b), "r"(idesc(N)), "r"(acc) : "memory");
                }
            }
            const long long t1 = clock64();
            asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
            phase ^= 1;
            const long long t2 = clock64();
            if (t2 - t0 < best) { best = t2 - t0; best_issue = t1 - t0; }
        }
        if (blockIdx.x == 0 && tid == 0) { out[0] = best; out[1] = best_issue; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int N, int SS, int NACC>
void run(long long *d_out) {
    long long h[2];
    cudaFuncSetAttribute(bench<N, SS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    bench<N, SS, NACC><<<148, 128, 98304>>>(20, d_out);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    printf("   N=%3d %5.0f/%-5.0f", N, (double)h[0] / 12, (double)h[1] / 12);
}
template <int SS, int NACC>
void row(long long *d_out) {
    printf("%s  accumulators=%d  L=12 unrolled:", SS ? "A smem" : "A tmem", NACC);
    run<16, SS, NACC>(d_out); run<32, SS, NACC>(d_out); run<64, SS, NACC>(d_out); run<96, SS, NACC>(d_out);
    run<128, SS, NACC>(d_out); run<192, SS, NACC>(d_out);
    if (NACC == 1) run<256, SS, NACC>(d_out);
    printf("\n");
}
int main() {
    long long *d_out;
    cudaMalloc(&d_out, 16);
    printf("cycles per MMA (M=128, K=16): total/issue-only\n");
    row<0, 1>(d_out); row<0, 2>(d_out); row<1, 1>(d_out); row<1, 2>(d_out);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
