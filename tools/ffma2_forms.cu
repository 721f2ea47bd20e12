// ffma2_forms.cu -- what a packed FP32 instruction costs on this device, by operand form.  The rollout kernels are bound by
// the FMA pipe ("math pipe throttle" is their top stall), yet the pipe reports 70-75 % active: this probe measures the
// issue-to-issue cost per warp instruction and scheduler (SM sub-partition) of each form the substep uses, with 6 independent chains per thread,
// from CUDA-event times of long launches (the per-warp clock64() stamps are kept for inspection only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_forms tools/ffma2_forms.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <algorithm>
#include <vector>

typedef unsigned long long u64;
#define DI __device__ __forceinline__
DI u64 pk(float a, float b) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }
DI float sum2(u64 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x + y; }
DI u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
DI u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DI u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int CH = 6;
enum Form { F_FMA_1U2S, F_FMA_2U1S, F_FMA_3U, F_FMA_UR, F_FMA_IMM, F_MUL_2U, F_MUL_UR, F_ADD_2U, F_FMA_RING, F_S_FMA_3U, F_S_FMA_UR,
            F_MIX_ALU, F_FMA_SQ, N_FORMS };
const char *NAMES[N_FORMS] = {
    "FFMA2 x = x*a + b        (a, b shared by the chains)", "FFMA2 x = x*y_i + b      (two per-chain pairs, b shared)",
    "FFMA2 x = x*y_i + z_i    (three distinct pairs)", "FFMA2 x = x*U + z_i      (uniform scalar, two pairs)",
    "FFMA2 x = x*y_i + 1.0    (immediate, two pairs)", "FMUL2 x = x*y_i", "FMUL2 x = x*U", "FADD2 x = x + y_i",
    "FFMA2 a = b*c + d ring   (three distinct pairs, fourth written)", "FFMA  x = x*y_i + z_i    (scalar, three registers)",
    "FFMA  x = x*U + z_i      (scalar, uniform operand)", "FFMA2 x*U + z_i  with 1.5 ALU instructions per FFMA2 beside it",
    "FFMA2 x = y_i*y_i + x    (one pair read twice)"};

template <int F>
__global__ void __launch_bounds__(256) probe(float *out, long long *stamps, int iters, const float *p, float U) {
    u64 x[CH], y[CH], z[CH], w[CH];
    float sx[CH], sy[CH], sz[CH], m = 0.0f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {   // every operand from its own memory location: nothing for the compiler to merge
        const float *q = p + 8 + 16 * i + (threadIdx.x & 1);
        x[i] = pk(q[0], q[2]); y[i] = pk(q[4], q[6]); z[i] = pk(q[8], q[10]); w[i] = pk(q[12], q[14]);
        sx[i] = q[1]; sy[i] = q[5]; sz[i] = q[9];
    }
    const u64 a = pk(p[4], p[5]), b = pk(p[6], p[7]), one = pk(1.0f, 1.0f);
    const u64 Uu = pk(U, U);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < iters; ++it) {   // 16 x CH instructions per trip: the moves ptxas puts at the loop end are < 10 %
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (F == F_FMA_1U2S) x[i] = fma2(x[i], a, b);
            if (F == F_FMA_2U1S) x[i] = fma2(x[i], y[i], b);
            if (F == F_FMA_3U) x[i] = fma2(x[i], y[i], z[i]);
            if (F == F_FMA_UR) x[i] = fma2(x[i], Uu, z[i]);
            if (F == F_FMA_IMM) x[i] = fma2(x[i], y[i], one);
            if (F == F_MUL_2U) x[i] = mul2(x[i], y[i]);
            if (F == F_MUL_UR) x[i] = mul2(x[i], Uu);
            if (F == F_ADD_2U) x[i] = add2(x[i], y[i]);
            if (F == F_FMA_RING) {   // four instructions per chain and iteration, no register moves
                x[i] = fma2(y[i], z[i], w[i]); y[i] = fma2(z[i], w[i], x[i]); z[i] = fma2(w[i], x[i], y[i]); w[i] = fma2(x[i], y[i], z[i]);
            }
            if (F == F_S_FMA_3U) sx[i] = fmaf(sx[i], sy[i], sz[i]);
            if (F == F_S_FMA_UR) sx[i] = fmaf(sx[i], U, sz[i]);
            if (F == F_MIX_ALU) { x[i] = fma2(x[i], Uu, z[i]); m = fmaxf(m, fabsf(sx[i])); sx[i] = __int_as_float(__float_as_int(sx[i]) + 12345); }
            if (F == F_FMA_SQ) x[i] = fma2(y[i], y[i], x[i]);
        }
    }
    const long long t1 = clock64();
    float r = m;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += sum2(x[i]) + sum2(y[i]) + sum2(z[i]) + sum2(w[i]) + sx[i];
    if (r == 123.456f) out[0] = r;
    if ((threadIdx.x & 31) == 0) {
        const int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        stamps[2 * wi] = t0; stamps[2 * wi + 1] = t1;
    }
}

typedef void (*kern)(float *, long long *, int, const float *, float);
template <int F> struct Tab { static void fill(kern *t) { t[F] = probe<F>; Tab<F + 1>::fill(t); } };
template <> struct Tab<N_FORMS> { static void fill(kern *) {} };

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, bps = 4, block = 256, iters = 4096;
    const int grid = sms * bps, warps = grid * block / 32;
    float hp[8 + 16 * CH + 2] = {1.0f, 0.5f, 0.25f, 0.125f, 0.9999999f, 0.9999998f, 1e-8f, 2e-8f};
    for (int i = 8; i < 8 + 16 * CH + 2; ++i) hp[i] = ((i - 8) % 16 < 4) ? 0.5f + 1e-3f * i : ((i - 8) % 16 < 8) ? 1.0f - 1e-7f * i : 1e-9f * i;
    float *dp, *dout;
    long long *dst;
    cudaMalloc(&dp, sizeof(hp)); cudaMemcpy(dp, hp, sizeof(hp), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 4); cudaMalloc(&dst, sizeof(long long) * 2 * warps);
    kern tab[N_FORMS];
    Tab<0>::fill(tab);
    // Event-timed (robust against how many blocks are co-resident): cycles = elapsed x SM clock x schedulers / warp
    // instructions, with the SM clock read through the attribute (boost clock; the rollout bench holds it under load).
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%d SMs at %d MHz, %d chains per thread; cycles per warp instruction and scheduler (1.0 = one issue slot per cycle)\n",
           sms, khz / 1000, CH);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int long_iters = 1 << 16;
    for (int f = 0; f < N_FORMS; ++f) {
        double best = 1e30;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            tab[f]<<<grid, block>>>(dout, dst, long_iters, dp, 0.99999f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double instr = (double)long_iters * CH * warps * (f == F_FMA_RING ? 4.0 : 1.0);
            const double per = ms * 1e-3 * khz * 1e3 * (sms * 4.0) / instr;
            if (rep > 0) best = std::min(best, per);
        }
        printf("  %-68s %6.3f\n", NAMES[f], best);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
