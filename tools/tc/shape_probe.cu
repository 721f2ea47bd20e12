// Probe: which (lane, column) does each thread of a warp get from the 16-lane tensor-memory load / store shapes
// (tcgen05.ld/st .16x256b, .16x128b, .16x64b), and may the lane field of the address be quarter base + 16?
// Pattern written with .32x32b stores: value(lane, column) = lane * 1000 + column.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(int *out) {
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tl = s_tmem + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < 64; ++c) {
        const uint32_t v = (uint32_t)((32 * warp + lane) * 1000 + c);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tl + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (warp == 1) {   // quarter 1: lanes 32..63
        uint32_t r[4];
        for (int half = 0; half < 2; ++half) {
            const uint32_t a = tl + ((uint32_t)(16 * half) << 16) + 8;   // column 8
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 4; ++i) out[(0 * 2 + half) * 128 + lane * 4 + i] = (int)r[i];
            asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 2; ++i) out[(1 * 2 + half) * 128 + lane * 4 + i] = (int)r[i];
            asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            out[(2 * 2 + half) * 128 + lane * 4] = (int)r[0];
        }
        // stores: 16x128b.x1 at column 40 (half 0 and 1), 16x64b.x1 at column 48; read back with 32x32b
        for (int half = 0; half < 2; ++half) {
            const uint32_t a = tl + ((uint32_t)(16 * half) << 16);
            uint32_t v0 = 100000u + (uint32_t)(lane * 10), v1 = v0 + 1;
            asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(a + 40), "r"(v0), "r"(v1) : "memory");
            uint32_t w0 = 200000u + (uint32_t)(lane * 10);
            asm volatile("tcgen05.st.sync.aligned.16x64b.x1.b32 [%0], {%1};" ::"r"(a + 48), "r"(w0) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        for (int c = 0; c < 16; ++c) {
            uint32_t v;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tl + 40 + c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            out[6 * 128 + lane * 16 + c] = (int)v;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(64));
}

int main() {
    int *d, h[6 * 128 + 512];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0xff, sizeof(h));
    probe<<<1, 128>>>(d);
    cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *nm[3] = {"ld 16x256b.x1", "ld 16x128b.x1", "ld 16x64b.x1"};
    const int nr[3] = {4, 2, 1};
    for (int s = 0; s < 3; ++s)
        for (int half = 0; half < 2; ++half) {
            printf("%s, address lane = quarter + %d, column 8: thread -> (lane.column) per register\n", nm[s], 16 * half);
            for (int t = 0; t < 32; ++t) {
                printf("  t%02d:", t);
                for (int i = 0; i < nr[s]; ++i) { int v = h[(s * 2 + half) * 128 + t * 4 + i]; printf(" %d.%d", v / 1000, v % 1000); }
                if (t % 4 == 3) printf("\n");
            }
        }
    printf("after st 16x128b.x1 {100000 + 10 t, +1} at column 40 and st 16x64b.x1 {200000 + 10 t} at column 48 (both halves): lanes 32..63 x columns 40..55\n");
    for (int l = 0; l < 32; ++l) {
        printf("  lane %d:", 32 + l);
        for (int c = 0; c < 16; ++c) printf(" %d", h[6 * 128 + l * 16 + c]);
        printf("\n");
    }
    return 0;
}
