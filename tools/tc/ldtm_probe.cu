// Micro-benchmark: what bounds tcgen05.ld (LDTM) -- bytes or instructions, per SM or per scheduler / lane quarter?
// W warps (warp w reads lane quarter w % 4) each issue REPS x UNROLL loads of one shape and wait; cycles per load and bytes
// per cycle and SM are printed for W = 1, 4, 8, 16.  Optionally a background warp keeps the tensor pipe busy with N = 96 MMAs
// (A operand in tensor memory) to see whether MMA traffic slows the loads down.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc/ldtm_probe tools/tc/ldtm_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__device__ __forceinline__ uint32_t do_load(uint32_t addr) {
    uint32_t v[16];
    if (SHAPE == 0) {   // 16x256b.x1: 16 lanes x 8 columns = 512 B
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
        return v[0] ^ v[1] ^ v[2] ^ v[3];
    } else if (SHAPE == 1) {   // 16x256b.x2: 16 lanes x 16 columns = 1024 B
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
        return v[0] ^ v[1] ^ v[2] ^ v[3] ^ v[4] ^ v[5] ^ v[6] ^ v[7];
    } else if (SHAPE == 2) {   // 32x32b.x8: 32 lanes x 8 columns = 1024 B
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
        return v[0] ^ v[1] ^ v[2] ^ v[3] ^ v[4] ^ v[5] ^ v[6] ^ v[7];
    } else if (SHAPE == 3) {   // 16x128b.x1: 16 lanes x 4 columns = 256 B
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(addr));
        return v[0] ^ v[1];
    } else {   // 32x32b.x16: 32 lanes x 16 columns = 2048 B
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr));
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) x ^= v[i];
        return x;
    }
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

template <int SHAPE>
__global__ void __launch_bounds__(544, 1) probe(int warps_active, int reps, int mma_bg, long long *out) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    __shared__ long long s_cyc[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) s_stop = 0;
    for (int i = tid; i < 8192; i += 544) reinterpret_cast<uint32_t *>(dsm)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = s_tmem;
    if (warp == 16) {
        if (mma_bg) {
            const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0), sm0 = smem_u32(dsm);
            const uint32_t idesc = (1u << 4) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            while (!s_stop) {
#pragma unroll
                for (int i = 0; i < 12; ++i) {
                    const uint32_t d = tmu + 256 + 128 * (i & 1), a = tmu + 448 + 8 * (i & 3);
                    const uint64_t b = make_desc(sm0 + 256 * (i & 3), 1024);
                    asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
                                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
                }
            }
        }
    } else if (warp < warps_active) {
        const uint32_t tl = tm + ((uint32_t)(32 * (warp & 3)) << 16);
        uint32_t acc = 0;
        __syncwarp();
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc ^= do_load<SHAPE>(tl + 16 * i);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        const long long t1 = clock64();
        if (lane == 0) s_cyc[warp] = t1 - t0;
        if (acc == 0x12345u) out[15] = 1;
        asm volatile("bar.sync 1, %0;" ::"r"(32 * warps_active));   // the timed warps only
        if (tid == 0) {
            long long m = 0;
            for (int w = 0; w < warps_active; ++w) m = s_cyc[w] > m ? s_cyc[w] : m;
            out[0] = m;
            s_stop = 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
}

template <int SHAPE>
void run(const char *name, int bytes) {
    long long *out;
    cudaMallocManaged(&out, 16 * sizeof(long long));
    cudaFuncSetAttribute(probe<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int reps = 2000;
    for (int bg = 0; bg < 2; ++bg)
        for (int w : {1, 4, 8, 16}) {
            out[0] = 0;
            probe<SHAPE><<<1, 544, 65536>>>(w, reps, bg, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
            const double per = (double)out[0] / (reps * 8.0);
            printf("%-12s warps %2d mma_bg %d: %6.1f cycles per load and warp, %7.1f B/cycle/SM\n", name, w, bg, per, bytes * (double)w / per);
        }
}

int main() {
    run<0>("16x256b.x1", 512);
    run<1>("16x256b.x2", 1024);
    run<3>("16x128b.x1", 256);
    run<2>("32x32b.x8", 1024);
    run<4>("32x32b.x16", 2048);
    return 0;
}
