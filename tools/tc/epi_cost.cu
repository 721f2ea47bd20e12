// Micro-benchmark of the tensor-core GRU kernel's gate epilogue (gru_epilogue) in isolation: cycles per call for one
// warp per scheduler, and how it splits into tensor-memory loads, arithmetic and stores.
#include "../../cartpolesimulation_b200/csrc/cps_net_tc.cu"
thread_local std::string g_create_err;   // cps_lib.cu's (this TU links alone)

// mma_bg != 0: warp 16 keeps the tensor pipe busy with N = 96 MMAs (A in tensor memory, accumulators in columns 128..) while
// the epilogue warps are timed -- does tensor-pipe traffic slow the epilogue's tensor-memory loads down?
__global__ void __launch_bounds__(544, 1) epi_bench(int warps_active, int reps, int mode, long long *out, int mma_bg) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ volatile int s_stop;
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_cst[512];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 512; i += 512) s_cst[i] = 0.01f * (i % 7);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { s_stop = 0; bar_init(smem_u32(&s_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 16384; i += 544) reinterpret_cast<uint32_t *>(dsm)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 16) {
        if (mma_bg) {
            const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0), sm0 = smem_u32(dsm);
            long long n = 0;
            while (!s_stop) {
#pragma unroll
                for (int i = 0; i < 12; ++i)
                    tc_mma(tm + 128 + 128 * (i & 1), tm + C_AH2_HI + 8 * (i & 3), make_desc(sm0 + 256 * (i & 3), 1024), idesc_f16(96), 1);
                n += 12;
                if (mma_bg == 2) __nanosleep(2000);   // bursts
            }
            tc_commit(smem_u32(&s_bar));
            bar_wait(smem_u32(&s_bar), 0);
            if (tid == 512) out[1] = n;
        }
    }
    const uint32_t tl = s_tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    if (warp < warps_active) {
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 512; c += 8) st8(tl + c, z);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        const int sub = (warp >> 2) & 3;
        long long best = 1LL << 60, tot = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            const int lane = tid & 31, hs = (warp >> 2) & 1;
            if (mode == 0) {          // 32 live rollouts per CTA: stacked hi / lo rows, one job of a warp
                gru_epilogue<2, true, true>(tl, 0, (uint32_t)(16 * hs), s_cst, 1e-3f, -1.4e-3f, C_AH1_HI, C_AH1_LO, 16 * hs, lane);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else if (mode == 3) {   // 64 live rollouts per CTA
                gru_epilogue<2, true>(tl, 0, (uint32_t)(16 * hs), s_cst, 1e-3f, -1.4e-3f, C_AH1_HI, C_AH1_LO, 16 * hs, lane);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else if (mode == 4) {   // 128 live rollouts per CTA: two calls per job
                for (int e = 0; e < 2; ++e)
                    gru_epilogue<2, false>(tl, 0, (uint32_t)(16 * hs + 8 * e), s_cst, 1e-3f, -1.4e-3f, C_AH1_HI, C_AH1_LO, 16 * hs + 8 * e, lane);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else if (mode == 1) {   // loads only
                uint32_t R[8], Z[8], NI[8], NH[8], PH[4], PL[4];
                ld8(tl + C_R + 8 * sub, R); ld8(tl + C_Z + 8 * sub, Z); ld8(tl + C_NI + 8 * sub, NI); ld8(tl + C_NH + 8 * sub, NH);
                ld4(tl + C_AH1_HI + 4 * sub, PH); ld4(tl + C_AH1_LO + 4 * sub, PL);
                ld_wait();
                if ((R[0] ^ Z[1] ^ NI[2] ^ NH[3] ^ PH[0] ^ PL[1]) == 0x12345u) out[3] = 1;
            } else {                  // stores only
                uint32_t P[4] = {1, 2, 3, 4};
                st4(tl + C_AH1_HI + 4 * sub, P); st4(tl + C_AH1_LO + 4 * sub, P);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
            tot += t1 - t0;
        }
        if (tid == 0) { out[0] = best; out[2] = tot / reps; }
    }
    __syncwarp();
    if (tid == 0) s_stop = 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512));
}

int main() {
    long long *d, h[4];
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute((const void *)epi_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int bg = 0; bg < 3; ++bg)
    for (int mode = 0; mode < 5; ++mode)
        for (int w : {1, 4, 8, 16}) {
            epi_bench<<<1, 544, 65536>>>(w, 50, mode, d, bg);
            cudaDeviceSynchronize();
            cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
            printf("bg MMAs %d: %s, %2d warps: best %lld mean %lld cycles per call (50 calls)\n", bg, mode == 0 ? "job, 32 rollouts (stacked)" : mode == 3 ? "job, 64 rollouts" : mode == 4 ? "job, 128 rollouts" : (mode == 1 ? "6 loads + wait " : "2 stores + wait"), w, h[0], h[2]);
        }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
