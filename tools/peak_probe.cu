// peak_probe.cu -- FP32 pipe throughput on this device for the instruction forms the rollout kernels can use:
//   ffma_const  d = fma(d, c[0][a], c[0][b])      (one register operand; what cps_measure_peaks uses)
//   ffma_reg    d = fma(d, r1, r2)                (three register operands)
//   ffma2_reg   fma.rn.f32x2 on 64-bit register pairs (two IEEE FMAs per issue slot)
//   mix         FFMA + MUFU.RCP + FMNMX mix shaped like one rotation substep
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peak_probe tools/peak_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) {
    u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d;
}
__device__ __forceinline__ float lo(u64 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x + y; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}

template <int CH>
__global__ void __launch_bounds__(256) k_const(float *out, int iters, float a, float b) {
    float x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x + i;
#pragma unroll 4
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fmaf(x[i], a, b);
    float r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += x[i];
    if (r == 123.456f) out[0] = r;
}

template <int CH>
__global__ void __launch_bounds__(256) k_reg(float *out, int iters, const float *p) {
    float x[CH];
    const float a = p[threadIdx.x & 1], b = p[2 + (threadIdx.x & 1)];  // loaded -> live registers
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x + i;
#pragma unroll 4
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fmaf(x[i], a, b);
    float r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += x[i];
    if (r == 123.456f) out[0] = r;
}

// three DISTINCT register operands per FFMA, rotating (defeats the operand reuse cache)
template <int CH>
__global__ void __launch_bounds__(256) k_reg3(float *out, int iters, const float *p) {
    float x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = p[(threadIdx.x + i) & 3];
#pragma unroll 2
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fmaf(x[(i + 1) % CH], x[(i + 2) % CH], x[i]);
    float r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += x[i];
    if (r == 123.456f) out[0] = r;
}

template <int CH>
__global__ void __launch_bounds__(256) k_ffma2(float *out, int iters, const float *p) {
    u64 x[CH];
    const u64 a = pk(p[threadIdx.x & 1], p[1]), b = pk(p[2 + (threadIdx.x & 1)], p[3]);
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = pk(threadIdx.x + i, i);
#pragma unroll 4
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = ffma2(x[i], a, b);
    float r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += lo(x[i]);
    if (r == 123.456f) out[0] = r;
}

template <int CH>
__global__ void __launch_bounds__(256) k_ffma2_3(float *out, int iters, const float *p) {
    u64 x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = pk(p[(threadIdx.x + i) & 3], p[i & 3]);
#pragma unroll 2
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = ffma2(x[(i + 1) % CH], x[(i + 2) % CH], x[i]);
    float r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += lo(x[i]);
    if (r == 123.456f) out[0] = r;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 8, block = 256, iters = 1 << 13;
    float *d; cudaMalloc(&d, 64);
    float hp[4] = {1.000001f, 1.000002f, 1e-7f, 2e-7f};
    cudaMemcpy(d, hp, sizeof(hp), cudaMemcpyHostToDevice);
    const double thr = (double)grid * block * iters;
    double ms;
    ms = time_ms([&] { k_const<8><<<grid, block>>>(d + 8, iters, 1.000001f, 1e-7f); });
    printf("ffma const-operand  : %7.2f TFLOP/s\n", 2 * 8 * thr / ms / 1e9);
    ms = time_ms([&] { k_reg<8><<<grid, block>>>(d + 8, iters, d); });
    printf("ffma reg (2 shared) : %7.2f TFLOP/s\n", 2 * 8 * thr / ms / 1e9);
    ms = time_ms([&] { k_reg3<8><<<grid, block>>>(d + 8, iters, d); });
    printf("ffma 3 distinct regs: %7.2f TFLOP/s\n", 2 * 8 * thr / ms / 1e9);
    ms = time_ms([&] { k_ffma2<8><<<grid, block>>>(d + 8, iters, d); });
    printf("ffma2 reg (2 shared): %7.2f TFLOP/s\n", 4 * 8 * thr / ms / 1e9);
    ms = time_ms([&] { k_ffma2_3<8><<<grid, block>>>(d + 8, iters, d); });
    printf("ffma2 3 distinct    : %7.2f TFLOP/s\n", 4 * 8 * thr / ms / 1e9);
    printf("SMs %d, clock %.0f MHz (max)\n", prop.multiProcessorCount, prop.clockRate / 1e3);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
