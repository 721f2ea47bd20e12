// cps_net_tc.cu -- the GRU predictor on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// net_tc_kernel<MPPI>: one CTA advances 128 rollouts (one per TMEM lane) of a 2 x 64 GRU through the horizon.
//   * the gate GEMMs [128 rollouts] x [192 gates] x [K = 16 | 64] run as tcgen05.mma.kind::f16 with fp32 accumulators
//     in tensor memory; weights (B operand) are resident in shared memory for the whole launch in the UMMA
//     K-major core-matrix layout; activations (A operand) live in TENSOR MEMORY (".ts" form), written by the thread
//     that owns the rollout with tcgen05.st -- no shared-memory round trip, no bank conflicts;
//   * fp32 accuracy from fp16 tensor cores: every operand is split x = hi + lo (two fp16 values after a power-of-two
//     pre-scale that keeps lo out of the subnormal range: activations x 2^7, weight rows scaled to [64, 128)), and
//     each product runs as three MMAs  hi*hi + lo*hi + hi*lo  into the same accumulator (the dropped lo*lo term is
//     2^-22 relative).  The scales are undone, and the biases added, by one FMA in the epilogue;
//   * 8 warps; thread (rollout, half) reads its 32 hidden units' accumulators with tcgen05.ld, applies the GRU
//     non-linearities, and writes h(t) back as the next A operand; the "lead" half of the threads also carries the
//     rollout's bookkeeping (feedback, de-normalisation, trajectory, cost, MPPI control) overlapped with the next
//     step's first-layer MMAs;
//   * MPPI = true: block partials, last-block merge and the stored-hidden-state update as in net_kernel.
// Reference: see cps_net.cu.  The tcgen05 building blocks were brought up with tools/tc/tc_test.cu.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "cps_net.cuh"

namespace {

constexpr int TC_H = 64, TC_ROWS = 128, TC_NT = 256;
// tensor-memory columns
constexpr uint32_t C_ACC = 0, C_OUT = 256, C_AX_HI = 288, C_AX_LO = 296, C_AH1_HI = 304, C_AH1_LO = 336, C_AH2_HI = 368,
                   C_AH2_LO = 400;
// image offsets (bytes)
constexpr uint32_t O_WIH1_HI = 0, O_WIH1_LO = 6144, O_WHH1_HI = 12288, O_WHH1_LO = 36864, O_WIH2_HI = 61440,
                   O_WIH2_LO = 86016, O_WHH2_HI = 110592, O_WHH2_LO = 135168, O_WOUT_HI = 159744, O_WOUT_LO = 161792,
                   O_CST1 = 163840, O_CST2 = 165888, O_CSTO = 167936, TC_IMAGE_BYTES = 168064;
constexpr float A_SCALE = 128.0f, A_INV = 1.0f / 128.0f;

__host__ __device__ inline uint32_t kmajor_off(int r, int k, int K) {  // UMMA K-major, no swizzle: 8 x 16 B core matrices
    return (uint32_t)((r >> 3) * ((K >> 3) * 128) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;   // LBO: K-adjacent core matrices are contiguous
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell)
    return d;
}
__device__ __forceinline__ constexpr uint32_t idesc_f16(int N) {  // D fp32, A/B fp16 K-major, M = 128
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// all threads: order this thread's tensor-memory accesses before the barrier and the issuer's MMAs after it
__device__ __forceinline__ void tc_sync() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}
__device__ __forceinline__ void ld16(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr));
}
__device__ __forceinline__ void ld8(uint32_t addr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
}
__device__ __forceinline__ void st8(uint32_t addr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// e^{-x} for x clamped to [-20, 20]: one FMUL-free MUFU.EX2 (ex2.approx.ftz of x * -log2 e)
__device__ __forceinline__ float ex2_neg(float x) {
    const float t = fminf(fmaxf(x, -20.0f), 20.0f) * -1.4426950408889634f;
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}
__device__ __forceinline__ float rcp_f(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// v (already scaled) -> fp16 hi + fp16 lo
__device__ __forceinline__ void split_h(float v, __half &hi, __half &lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    const __half2 h = __halves2half2(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&v));
}

// 16 consecutive values of one rollout -> A operand (8 hi + 8 lo columns at column offset c0 of the two regions)
__device__ __forceinline__ void write_operand16(uint32_t tl, uint32_t c_hi, uint32_t c_lo, const float (&v)[16]) {
    uint32_t ph[8], pl[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        __half h0, l0, h1, l1;
        split_h(v[2 * q] * A_SCALE, h0, l0);
        split_h(v[2 * q + 1] * A_SCALE, h1, l1);
        ph[q] = pack_h2(h0, h1);
        pl[q] = pack_h2(l0, l1);
    }
    st8(tl + c_hi, ph);
    st8(tl + c_lo, pl);
}
// ... and back: (hi + lo) / scale
__device__ __forceinline__ void read_operand16(uint32_t tl, uint32_t c_hi, uint32_t c_lo, float (&v)[16]) {
    uint32_t ph[8], pl[8];
    ld8(tl + c_hi, ph);
    ld8(tl + c_lo, pl);
    ld_wait();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float2 h = unpack_h2(ph[q]), l = unpack_h2(pl[q]);
        v[2 * q] = (h.x + l.x) * A_INV;
        v[2 * q + 1] = (h.y + l.y) * A_INV;
    }
}

// One GRU layer's MMAs: x-part (A = ax, K = Kx, W_ih rows 0..191 -> ACC[0,192)) then h-part (A = ah, K = 64, W_hh rows
// 0..127 -> ACC[0,128) accumulating, rows 128..191 -> ACC[192,256)); three split passes each.  One thread.
__device__ __forceinline__ void issue_layer(uint32_t tmem, uint32_t ax_hi, uint32_t ax_lo, int Kx, uint32_t bx_hi,
                                            uint32_t bx_lo, uint32_t ah_hi, uint32_t ah_lo, uint32_t bh_hi, uint32_t bh_lo,
                                            uint32_t bar) {
    uint32_t acc = 0;
    const uint32_t sbo_x = (uint32_t)(Kx >> 3) * 128u;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = (pass == 1) ? ax_lo : ax_hi, b = (pass == 2) ? bx_lo : bx_hi;
        for (int ks = 0; ks < (Kx >> 4); ++ks) {
            tc_mma(tmem + C_ACC, tmem + a + 8 * ks, make_desc(b + 256 * ks, sbo_x), idesc_f16(192), acc);
            acc = 1;
        }
    }
    uint32_t accn = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = (pass == 1) ? ah_lo : ah_hi, b = (pass == 2) ? bh_lo : bh_hi;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            tc_mma(tmem + C_ACC, tmem + a + 8 * ks, make_desc(b + 256 * ks, 1024), idesc_f16(128), 1);
            tc_mma(tmem + C_ACC + 192, tmem + a + 8 * ks, make_desc(b + 16 * 1024 + 256 * ks, 1024), idesc_f16(64), accn);
            accn = 1;
        }
    }
    tc_commit(bar);
}
__device__ __forceinline__ void issue_out(uint32_t tmem, uint32_t b_hi, uint32_t b_lo, uint32_t bar) {
    uint32_t acc = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = (pass == 1) ? C_AH2_LO : C_AH2_HI, b = (pass == 2) ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            tc_mma(tmem + C_OUT, tmem + a + 8 * ks, make_desc(b + 256 * ks, 1024), idesc_f16(16), acc);
            acc = 1;
        }
    }
    tc_commit(bar);
}

// GRU non-linearities for this thread's 32 hidden units (torch GRUCell, gate order r, z, n):
// reads ACC = {R, Z, NI, NH} and h(t-1) (the A operand itself), writes h(t) as the new A operand.
__device__ __forceinline__ void gru_epilogue(uint32_t tl, const float *cst, uint32_t c_hi, uint32_t c_lo, int half) {
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
        const int u0 = 32 * half + 16 * ch;
        uint32_t R[16], Z[16], NI[16], NH[16];
        ld16(tl + C_ACC + u0, R);
        ld16(tl + C_ACC + 64 + u0, Z);
        ld16(tl + C_ACC + 128 + u0, NI);
        ld16(tl + C_ACC + 192 + u0, NH);
        float h[16];
        read_operand16(tl, c_hi + (u0 >> 1), c_lo + (u0 >> 1), h);   // includes the tcgen05.wait::ld
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            // two units at a time: 6 exponentials + 2 reciprocals (instead of 6 + 6): 1/a = (b c d) / (a b c d).
            // Pre-activations are clamped to +-20 (sigmoid(-20) = 2e-9, tanh(10) = 1 - 4e-9), which bounds the products.
            const float4 c0 = *reinterpret_cast<const float4 *>(cst + (u0 + i) * 8);
            const float4 c1 = *reinterpret_cast<const float4 *>(cst + (u0 + i) * 8 + 4);
            const float4 d0 = *reinterpret_cast<const float4 *>(cst + (u0 + i + 1) * 8);
            const float4 d1 = *reinterpret_cast<const float4 *>(cst + (u0 + i + 1) * 8 + 4);
            const float ar = 1.0f + ex2_neg(fmaf(__uint_as_float(R[i]), c0.x, c0.y));
            const float az = 1.0f + ex2_neg(fmaf(__uint_as_float(Z[i]), c0.z, c0.w));
            const float br = 1.0f + ex2_neg(fmaf(__uint_as_float(R[i + 1]), d0.x, d0.y));
            const float bz = 1.0f + ex2_neg(fmaf(__uint_as_float(Z[i + 1]), d0.z, d0.w));
            const float pa = ar * az, pb = br * bz;
            const float inv = rcp_f(pa * pb);
            const float ia = pb * inv, ib = pa * inv;           // 1/(ar az), 1/(br bz)
            const float r0 = az * ia, z0 = ar * ia, r1 = bz * ib, z1 = br * ib;
            const float n0p = fmaf(r0, fmaf(__uint_as_float(NH[i]), c1.z, c1.w), fmaf(__uint_as_float(NI[i]), c1.x, c1.y));
            const float n1p = fmaf(r1, fmaf(__uint_as_float(NH[i + 1]), d1.z, d1.w), fmaf(__uint_as_float(NI[i + 1]), d1.x, d1.y));
            const float e0 = 1.0f + ex2_neg(-2.0f * n0p), e1 = 1.0f + ex2_neg(-2.0f * n1p);   // 1 + e^{2 n}
            const float inv2 = rcp_f(e0 * e1);
            const float n0 = fmaf(-2.0f * e1, inv2, 1.0f), n1 = fmaf(-2.0f * e0, inv2, 1.0f);  // tanh = 1 - 2 / (1 + e^{2n})
            h[i] = fmaf(h[i] - n0, z0, n0);
            h[i + 1] = fmaf(h[i + 1] - n1, z1, n1);
        }
        write_operand16(tl, c_hi + (u0 >> 1), c_lo + (u0 >> 1), h);
    }
}

}  // namespace

template <bool MPPI>
__global__ void __launch_bounds__(TC_NT, 1) net_tc_kernel(const __grid_constant__ NetArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_wbar, s_mbar;
    __shared__ unsigned s_ticket;
    __shared__ float s_bmin[4];
    const NetDev &N = a.net;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = warp >> 2, row = 32 * (warp & 3) + lane;
    const bool lead = half == 0;
    const int T = a.T;
    const int row0 = blockIdx.x * TC_ROWS;
    const int k = row0 + row;
    const bool active = lead && k < a.B;
    const int kc = min(k, a.B - 1);
    float *s_f = reinterpret_cast<float *>(smem + TC_IMAGE_BYTES);
    float *s_unom = s_f;                                   // MPPI: [T] shifted nominal inputs, [p] w0, [p] w1, scratch
    float *s_w0 = s_unom + (MPPI ? a.mp.T : 0);
    float *s_w1 = s_w0 + (MPPI ? a.mp.p : 0);
    float *s_red = s_w1 + (MPPI ? a.mp.p : 0);             // [4][n_red + 2] then [n_red + 2]
    const float *cst1 = reinterpret_cast<const float *>(smem + O_CST1);
    const float *cst2 = reinterpret_cast<const float *>(smem + O_CST2);
    const float *csto = reinterpret_cast<const float *>(smem + O_CSTO);
    const uint32_t sm0 = smem_u32(smem), wbar = smem_u32(&s_wbar), mbar = smem_u32(&s_mbar);

    // ---- weights: bulk asynchronous copy; tensor memory; MPPI tables ------------------------------------------------
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wbar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar), "r"(TC_IMAGE_BYTES) : "memory");
        for (uint32_t done = 0; done < TC_IMAGE_BYTES; done += 32768u) {
            const uint32_t chunk = min(TC_IMAGE_BYTES - done, 32768u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sm0 + done), "l"(a.tc + done), "r"(chunk), "r"(wbar) : "memory");
        }
    }
    __syncwarp();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (MPPI) {
        for (int t = tid; t < T; t += TC_NT) s_unom[t] = a.u_nom[min(t + 1, T - 1)];   // warm-start shift (:183)
        for (int j = tid; j < a.mp.p; j += TC_NT) {
            s_w0[j] = (float)(a.mp.p - j) / (float)a.mp.p;
            s_w1[j] = (float)j / (float)a.mp.p;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);   // this warp's lane quarter

    // ---- row bookkeeping (lead threads) ---------------------------------------------------------------------------------
    float Jacc = 0.0f, corr = 0.0f, up = a.u_prev, u_cur = 0.0f, du_cur = 0.0f, u_nxt = 0.0f, du_nxt = 0.0f;
    int seg = 0, jj = 0;
    float na = 0.0f, nb = 0.0f;
    const float *nz = nullptr, *qrow = nullptr;
    float *traj = (a.traj_out && active) ? a.traj_out + (long long)k * a.ts_k : nullptr;
    if (MPPI) {
        nz = a.noise + (long long)kc * a.ns_k;
        na = nz[0] * a.mp.sigma;
        nb = (a.mp.n_ind > 1) ? nz[a.ns_i] * a.mp.sigma : 0.0f;
    } else {
        qrow = a.Q + (long long)kc * a.qs_b;
    }
    auto next_control = [&](int t) {
        if (MPPI) {
            const MppiParams &mp = a.mp;
            du_nxt = (seg == mp.n_ind - 1) ? na * mp.inv_p : fmaf(na, s_w0[jj], nb * s_w1[jj]);
            if (++jj == mp.p) {
                jj = 0; ++seg; na = nb;
                nb = (seg + 1 < mp.n_ind) ? nz[(long long)(seg + 1) * a.ns_i] * mp.sigma : 0.0f;
            }
            u_nxt = clampf(s_unom[t] + du_nxt, mp.lo, mp.hi);
        } else {
            u_nxt = qrow[(long long)t * a.qs_t];
        }
    };
    // network input x = [control, state features] -> A operand (K padded to 16)
    auto write_x = [&](float ctrl, const float (&feat)[6]) {
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = 0.0f;
        x[0] = fmaf(N.norm_a[0], ctrl, N.norm_b[0]);
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i < N.n_state_in) x[1 + i] = feat[i];
        write_operand16(tl, C_AX_HI, C_AX_LO, x);
    };

    // ---- initial operands -------------------------------------------------------------------------------------------------
    {
        float v[16];
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                const int u0 = 32 * half + 16 * ch;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = a.h0[(long long)kc * a.hs_b + l * TC_H + u0 + i];
                write_operand16(tl, (l ? C_AH2_HI : C_AH1_HI) + (u0 >> 1), (l ? C_AH2_LO : C_AH1_LO) + (u0 >> 1), v);
            }
        }
    }
    float st[6], y[6], feat[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) st[c] = a.s0[(long long)kc * a.ss_b + c];
    if (lead) {
        next_control(0);
#pragma unroll
        for (int i = 0; i < 6; ++i)
            feat[i] = (i < N.n_state_in) ? fmaf(N.norm_a[1 + i], a.s0[(long long)kc * a.ss_b + N.in_idx[i]], N.norm_b[1 + i]) : 0.0f;
        write_x(u_nxt, feat);
    }
    bar_wait(wbar, 0);   // weights have landed
    tc_sync();

    // ---- the horizon ------------------------------------------------------------------------------------------------------
    uint32_t phase = 0;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        if (tid == 0)
            issue_layer(tmem, C_AX_HI, C_AX_LO, 16, sm0 + O_WIH1_HI, sm0 + O_WIH1_LO, C_AH1_HI, C_AH1_LO, sm0 + O_WHH1_HI,
                        sm0 + O_WHH1_LO, mbar);
        if (lead) {  // under the first-layer MMAs: state s_t -> trajectory row, stage cost; next control
            if (t > 0) compose_state(N, y, st);
            if (traj) {
#pragma unroll
                for (int c = 0; c < 6; ++c) traj[(long long)t * a.ts_t + c * a.ts_c] = st[c];
            }
            u_cur = u_nxt; du_cur = du_nxt;
            if (MPPI) {
                Jacc += stage_cost_rt(a.cost_id, a.cost, cosf(st[IDX_ANGLE]), st[IDX_ANGLED], st[IDX_POS], u_cur, up);
                corr = fmaf(a.mp.cc_half_nu * du_cur, du_cur,
                            fmaf(a.mp.cc_R * u_cur, du_cur, fmaf(a.mp.cc_half_R * u_cur, u_cur, corr)));
                if (a.u_run_out && active) a.u_run_out[(long long)k * T + t] = u_cur;
                up = u_cur;
            }
            if (t + 1 < T) next_control(t + 1);
        }
        bar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        gru_epilogue(tl, cst1, C_AH1_HI, C_AH1_LO, half);
        tc_sync();
        if (tid == 0)
            issue_layer(tmem, C_AH1_HI, C_AH1_LO, 64, sm0 + O_WIH2_HI, sm0 + O_WIH2_LO, C_AH2_HI, C_AH2_LO, sm0 + O_WHH2_HI,
                        sm0 + O_WHH2_LO, mbar);
        bar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        gru_epilogue(tl, cst2, C_AH2_HI, C_AH2_LO, half);
        tc_sync();
        if (tid == 0) issue_out(tmem, sm0 + O_WOUT_HI, sm0 + O_WOUT_LO, mbar);
        bar_wait(mbar, phase); phase ^= 1;
        tc_fence_after();
        if (lead) {  // linear output layer -> feedback as the next input (autoregression.py:94-98)
            uint32_t o[8];
            ld8(tl + C_OUT, o);
            ld_wait();
#pragma unroll
            for (int i = 0; i < 6; ++i) y[i] = (i < N.n_out) ? fmaf(__uint_as_float(o[i]), csto[2 * i], csto[2 * i + 1]) : 0.0f;
            if (t + 1 < T) write_x(u_nxt, y);
        }
        tc_sync();
    }

    // ---- last state, costs, hidden state out ---------------------------------------------------------------------------
    float J = 0.0f;
    if (lead) {
        compose_state(N, y, st);
        if (traj) {
#pragma unroll
            for (int c = 0; c < 6; ++c) traj[(long long)T * a.ts_t + c * a.ts_c] = st[c];
        }
        if (MPPI) {
            if (a.cost_id == CPS_COST_DEFAULT || a.cost_id == CPS_COST_QUADRATIC_BOUNDARY)
                Jacc += terminal_cost<COST_DEFAULT>(a.cost, st[IDX_ANGLE], st[IDX_POS]);
            J = __fdiv_rn(Jacc, a.mp.T1) + corr;   // mean over T+1 entries (true division, as torch.mean) + correction
            if (active) {
                if (a.J_out) a.J_out[k] = J;
                if (!isfinite(J)) atomicAdd(a.nonfinite, 1);
            }
        }
    }
    auto store_hidden = [&](float *dst) {  // this thread's 2 x 32 units of the current hidden state
        float v[16];
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                const int u0 = 32 * half + 16 * ch;
                read_operand16(tl, (l ? C_AH2_HI : C_AH1_HI) + (u0 >> 1), (l ? C_AH2_LO : C_AH1_LO) + (u0 >> 1), v);
                if (dst) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) dst[l * TC_H + u0 + i] = v[i];
                }
            }
        }
    };
    if (a.h_final) store_hidden((k < a.B) ? a.h_final + (long long)k * (2 * TC_H) : nullptr);

    bool last = false;
    if (MPPI) {
        // ---- block partial {min J, sum w, sum w*eps[.]} over the 128 rollouts (lead warps 0..3) ----------------------
        const MppiParams &mp = a.mp;
        const int rec = 2 + mp.n_red;
        if (lead) {
            const float m = warp_min(active ? J : INFINITY);
            if (lane == 0) s_bmin[warp] = m;
        }
        __syncthreads();
        const float m = fminf(fminf(s_bmin[0], s_bmin[1]), fminf(s_bmin[2], s_bmin[3]));
        if (lead) {
            const float wgt = active ? expf(-(J - m) * mp.inv_lambda) : 0.0f;
            const float S = warp_sum(wgt);
            if (lane == 0) s_red[warp * rec + 1] = S;
            for (int i = 0; i < mp.n_red; ++i) {
                const float e = active ? nz[(long long)i * a.ns_i] : 0.0f;
                const float v = warp_sum(wgt * e);
                if (lane == 0) s_red[warp * rec + 2 + i] = v;
            }
        }
        __syncthreads();
        float *part = a.partials + (size_t)blockIdx.x * rec;
        for (int c = tid; c < mp.n_red + 1; c += TC_NT)
            part[1 + c] = (s_red[1 + c] + s_red[rec + 1 + c]) + (s_red[2 * rec + 1 + c] + s_red[3 * rec + 1 + c]);
        if (tid == 0) part[0] = m;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u);
        __syncthreads();
        last = s_ticket == gridDim.x - 1;
        if (last) {
            __threadfence();
            merge_and_finish(mp, a.partials, gridDim.x, s_red, s_unom, a.u_nom, a.u_out, a.shard_out, false);
            if (tid == 0) *a.ticket = 0u;
        }
        if (last && !a.shard_out && a.h_ref) {
            // ---- advance the stored hidden state by one step on (u, s) (optimizer_mppi.py:191,194-196) ----------------
            __syncthreads();
            const float u_sel = __ldcg(a.u_out);
            float v[16];
#pragma unroll 1
            for (int l = 0; l < 2; ++l) {
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    const int u0 = 32 * half + 16 * ch;
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = a.h_ref[l * TC_H + u0 + i];
                    write_operand16(tl, (l ? C_AH2_HI : C_AH1_HI) + (u0 >> 1), (l ? C_AH2_LO : C_AH1_LO) + (u0 >> 1), v);
                }
            }
            if (lead) {
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    feat[i] = (i < N.n_state_in) ? fmaf(N.norm_a[1 + i], a.s0[N.in_idx[i]], N.norm_b[1 + i]) : 0.0f;
                write_x(u_sel, feat);
            }
            tc_sync();
            if (tid == 0)
                issue_layer(tmem, C_AX_HI, C_AX_LO, 16, sm0 + O_WIH1_HI, sm0 + O_WIH1_LO, C_AH1_HI, C_AH1_LO, sm0 + O_WHH1_HI,
                            sm0 + O_WHH1_LO, mbar);
            bar_wait(mbar, phase); phase ^= 1;
            tc_fence_after();
            gru_epilogue(tl, cst1, C_AH1_HI, C_AH1_LO, half);
            tc_sync();
            if (tid == 0)
                issue_layer(tmem, C_AH1_HI, C_AH1_LO, 64, sm0 + O_WIH2_HI, sm0 + O_WIH2_LO, C_AH2_HI, C_AH2_LO, sm0 + O_WHH2_HI,
                            sm0 + O_WHH2_LO, mbar);
            bar_wait(mbar, phase); phase ^= 1;
            tc_fence_after();
            gru_epilogue(tl, cst2, C_AH2_HI, C_AH2_LO, half);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            store_hidden(row == 0 ? a.h_ref : nullptr);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// =====================================================================================================
// host side
// =====================================================================================================
bool cps_net_tc_eligible(const NetDev &N) {
    return N.type == CPS_NET_GRU && N.n_layers == 2 && N.hsz[0] == TC_H && N.hsz[1] == TC_H && N.n_in <= 7 && N.n_out <= 6
           && !N.differential;
}

size_t cps_net_tc_smem(const MppiParams *mp, bool mppi) {
    size_t f = 0;
    if (mppi) f = (size_t)mp->T + 2 * (size_t)mp->p + 5 * ((size_t)mp->n_red + 2) + 8;
    return TC_IMAGE_BYTES + f * sizeof(float);
}

// Builds the kernel image from the torch-order weights: fp16 hi/lo parts of the row-scaled matrices in UMMA core-matrix
// order + the per-column epilogue constants {c = 1 / (128 s_n), b = bias}.
void cps_net_tc_build_image(const NetDev &N, const float *w, std::vector<unsigned char> &img) {
    img.assign(TC_IMAGE_BYTES, 0);
    const int H = TC_H, n_in = N.n_in;
    const float *w_ih1 = w, *w_hh1 = w_ih1 + 3 * H * n_in, *b_ih1 = w_hh1 + 3 * H * H, *b_hh1 = b_ih1 + 3 * H;
    const float *w_ih2 = b_hh1 + 3 * H, *w_hh2 = w_ih2 + 3 * H * H, *b_ih2 = w_hh2 + 3 * H * H, *b_hh2 = b_ih2 + 3 * H;
    const float *w_out = b_hh2 + 3 * H, *b_out = w_out + N.n_out * H;
    auto row_scale = [](const float *r1, int n1, const float *r2, int n2) {
        float m = 0.0f;
        for (int i = 0; i < n1; ++i) m = fmaxf(m, fabsf(r1[i]));
        for (int i = 0; i < n2; ++i) m = fmaxf(m, fabsf(r2[i]));
        if (!(m > 0.0f)) return 1.0f;
        int e;
        frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)
        return ldexpf(1.0f, 7 - e);  // s * m in [64, 128)
    };
    auto put = [&](uint32_t o_hi, uint32_t o_lo, int r, int kk, int K, float v) {
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        memcpy(&img[o_hi + kmajor_off(r, kk, K)], &hi, 2);
        memcpy(&img[o_lo + kmajor_off(r, kk, K)], &lo, 2);
    };
    auto layer = [&](const float *wih, int kin, int Kpad, const float *whh, const float *bih, const float *bhh, uint32_t oih_hi,
                     uint32_t oih_lo, uint32_t ohh_hi, uint32_t ohh_lo, uint32_t ocst) {
        float *cst = reinterpret_cast<float *>(&img[ocst]);
        for (int n = 0; n < 3 * H; ++n) {
            const float s = row_scale(wih + (size_t)n * kin, kin, whh + (size_t)n * H, H);
            for (int kk = 0; kk < kin; ++kk) put(oih_hi, oih_lo, n, kk, Kpad, wih[(size_t)n * kin + kk] * s);
            for (int kk = 0; kk < H; ++kk) put(ohh_hi, ohh_lo, n, kk, H, whh[(size_t)n * H + kk] * s);
            const float c = 1.0f / (A_SCALE * s);
            const int u = n % H, g = n / H;   // gate 0 = r, 1 = z, 2 = n
            if (g == 0) { cst[u * 8 + 0] = c; cst[u * 8 + 1] = bih[n] + bhh[n]; }
            else if (g == 1) { cst[u * 8 + 2] = c; cst[u * 8 + 3] = bih[n] + bhh[n]; }
            else { cst[u * 8 + 4] = c; cst[u * 8 + 5] = bih[n]; cst[u * 8 + 6] = c; cst[u * 8 + 7] = bhh[n]; }
        }
    };
    layer(w_ih1, n_in, 16, w_hh1, b_ih1, b_hh1, O_WIH1_HI, O_WIH1_LO, O_WHH1_HI, O_WHH1_LO, O_CST1);
    layer(w_ih2, H, H, w_hh2, b_ih2, b_hh2, O_WIH2_HI, O_WIH2_LO, O_WHH2_HI, O_WHH2_LO, O_CST2);
    float *co = reinterpret_cast<float *>(&img[O_CSTO]);
    for (int o = 0; o < 16; ++o) { co[2 * o] = 0.0f; co[2 * o + 1] = 0.0f; }
    for (int o = 0; o < N.n_out; ++o) {
        const float s = row_scale(w_out + (size_t)o * H, H, nullptr, 0);
        for (int kk = 0; kk < H; ++kk) put(O_WOUT_HI, O_WOUT_LO, o, kk, H, w_out[(size_t)o * H + kk] * s);
        co[2 * o] = 1.0f / (A_SCALE * s);
        co[2 * o + 1] = b_out[o];
    }
}

int cps_net_tc_image_bytes() { return (int)TC_IMAGE_BYTES; }

int cps_net_tc_launch(cps_handle *h, NetArgs &a, bool mppi, int n_rows) {
    const size_t smem = cps_net_tc_smem(&h->mp, mppi);
    void (*fn)(const NetArgs) = mppi ? net_tc_kernel<true> : net_tc_kernel<false>;
    CUDA_TRY(h, cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (n_rows + TC_ROWS - 1) / TC_ROWS;
    fn<<<grid, TC_NT, smem, h->stream>>>(a);
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return CPS_OK;
}
