// cps_net_tc_common.cuh -- building blocks shared by the tensor-core network kernels (cps_net_tc.cu, cps_net_tc2.cu): image
// layout of the fp16 hi/lo weights, UMMA descriptors, tcgen05 issue / commit / tensor-memory access wrappers, mbarrier
// helpers and the GRU gate epilogue.  Include inside an anonymous namespace.

constexpr int TC_H = 64, TC_ROWS = 128;
constexpr int TC_EPI = 512, TC_ROWT = 128, TC_NT = TC_EPI + TC_ROWT + 32;   // epilogue warps 0-15, row warps 16-19, issuer warp 20
constexpr int TC_EW = TC_EPI / 32, TC_EU = 32 / (TC_EW / 4);            // epilogue warps; units per epilogue thread and job (8)
// tensor-memory columns: accumulator region r at 128 r = {NH [0,32), R [32,64), Z [64,96), NI [96,128)}: the recurrent
// product of a job is ONE N = 96 MMA per k-step into {NH, R, Z}, the input product one N = 96 MMA into {R, Z, NI}
constexpr uint32_t C_NH = 0, C_R = 32, C_Z = 64, C_NI = 96, C_AH1_HI = 384, C_AH1_LO = 416, C_AH2_HI = 448, C_AH2_LO = 480;
// image offsets (bytes).  Weight rows are ordered by half-layer job jh (units 32 jh .. 32 jh + 31): 96 rows per job, gate
// blocks [r | z | n] in the input matrices W_ih and [n | r | z] in the recurrent matrices W_hh (matching the columns above).
constexpr uint32_t O_WIH1_HI = 0, O_WIH1_LO = 6144, O_WHH1_HI = 12288, O_WHH1_LO = 36864, O_WIH2_HI = 61440,
                   O_WIH2_LO = 86016, O_WHH2_HI = 110592, O_WHH2_LO = 135168, O_WOUT_HI = 159744, O_WOUT_LO = 161792,
                   O_CST1 = 163840, O_CST2 = 165888, O_CSTO = 167936, TC_IMAGE_BYTES = 168192;
// after the image: the first layer's A operand x = [control, state features] (K = 16, hi and lo tiles), then float scratch
constexpr uint32_t O_X_HI = TC_IMAGE_BYTES, O_X_LO = O_X_HI + 4096, O_FLOATS = O_X_LO + 4096;
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
// The MMA-issuing code is executed by ALL lanes of the issuer warp, converged, on warp-uniform operands; one elected lane
// executes the instruction.  (Issued from a divergent `if (lane == 0)` branch the compiler wraps every UTCHMMA into a
// lane-uniformisation loop and the descriptor arithmetic leaves the uniform datapath: measured ~200 cycles per MMA
// instead of N / 2 + 10, tools/tc/mma_cost.cu.)
// A operand in tensor memory
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// A operand in shared memory
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {   // whole warp, one elected lane arrives
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Waiting warps share their scheduler with working ones: the suspend-time hint parks the thread in hardware until the phase
// completes (or the hint expires) instead of polling, which would take issue slots from the epilogue warps.
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
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
// one warp -> the issuer: this warp's tensor-memory stores / loads are done (count one arrival per warp)
__device__ __forceinline__ void warp_signal(uint32_t bar, int lane) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) bar_arrive(bar);
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
__device__ __forceinline__ void ld4(uint32_t addr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void st4(uint32_t addr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16-lane shapes (layouts verified with tools/tc/shape_probe.cu): the address names lane L = quarter base (+ 16) and column c.
//   .16x256b.x1: thread t gets {(L + t/4, c + 2 (t%4)), (L + t/4, c + 2 (t%4) + 1), (L + 8 + t/4, same two columns)}
//   .16x128b.x1: thread t gets {(L + t/4, c + t%4), (L + 8 + t/4, c + t%4)}
// i.e. four threads share a rollout: the gate epilogue of 16 rollouts x 8 units is spread over the whole warp.
__device__ __forceinline__ void ld16x256(uint32_t addr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void ld16x128(uint32_t addr, uint32_t (&v)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(addr));
}
__device__ __forceinline__ void st16x128(uint32_t addr, const uint32_t (&v)[2]) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v[0]), "r"(v[1]) : "memory");
}

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

// 8 consecutive values of one rollout -> A operand (4 hi + 4 lo columns at column offset c0 of the two regions)
__device__ __forceinline__ void write_operand8(uint32_t tl, uint32_t c_hi, uint32_t c_lo, const float (&v)[8]) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __half h0, l0, h1, l1;
        split_h(v[2 * q] * A_SCALE, h0, l0);
        split_h(v[2 * q + 1] * A_SCALE, h1, l1);
        ph[q] = pack_h2(h0, h1);
        pl[q] = pack_h2(l0, l1);
    }
    st4(tl + c_hi, ph);
    st4(tl + c_lo, pl);
}
// ... and back: (hi + lo) / scale
__device__ __forceinline__ void read_operand8(uint32_t tl, uint32_t c_hi, uint32_t c_lo, float (&v)[8]) {
    uint32_t ph[4], pl[4];
    ld4(tl + c_hi, ph);
    ld4(tl + c_lo, pl);
    ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 h = unpack_h2(ph[q]), l = unpack_h2(pl[q]);
        v[2 * q] = (h.x + l.x) * A_INV;
        v[2 * q + 1] = (h.y + l.y) * A_INV;
    }
}

// Stacked layout (see gru_epilogue): row r of a 16-row group holds the hi parts, row r + 8 the lo parts, both in the HI region.
// v: the values of rollout (lane & 7) of the quarter; lanes with bit 3 set write the lo parts.
__device__ __forceinline__ void write_operand8_stacked(uint32_t tl, uint32_t c_hi, const float (&v)[8], int lane) {
    uint32_t p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __half h0, l0, h1, l1;
        split_h(v[2 * q] * A_SCALE, h0, l0);
        split_h(v[2 * q + 1] * A_SCALE, h1, l1);
        p[q] = (lane & 8) ? pack_h2(l0, l1) : pack_h2(h0, h1);
    }
    st4(tl + c_hi, p);
}
// ... and back: valid in the lanes r < 8 of every 16-row group (whole warp calls)
__device__ __forceinline__ void read_operand8_stacked(uint32_t tl, uint32_t c_hi, float (&v)[8]) {
    uint32_t p[4];
    ld4(tl + c_hi, p);
    ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 h = unpack_h2(p[q]);
        v[2 * q] = (h.x + __shfl_down_sync(0xffffffffu, h.x, 8)) * A_INV;
        v[2 * q + 1] = (h.y + __shfl_down_sync(0xffffffffu, h.y, 8)) * A_INV;
    }
}

// ---- MMA issue (issuer warp, converged).  Half-layer job jh of a layer: weight rows [96 jh, 96 jh + 96). ----------------
// Recurrent part W_hh h: one N = 96 MMA per pass and k-step -> {NH, R, Z} of the region, overwriting it.
// stk (warp-uniform): the A operand carries the hi parts in rows r and the lo parts in rows r + 8 of every 16-row group
// ("stacked", see gru_epilogue): the a_lo pass is skipped, rows r + 8 accumulate it beside rows r in the SAME two MMAs.
// NKS (compile time): k-steps of 16 units that carry live units of the A operand's layer -- 4 for 64 units, 2 when the layer has
// at most 32 (the k-steps of a narrower layer's zero padding would add exact zeros).  A run-time bound here costs the
// full-width kernel a third of its speed: the issue sequence has to stay straight-line.
template <int NKS = 4>
__device__ __forceinline__ void issue_H(uint32_t region, uint32_t ah_hi, uint32_t ah_lo, uint32_t bh_hi, uint32_t bh_lo, int jh, bool stk) {
    const uint32_t row = (uint32_t)jh * 12288u;   // 96 rows x (64 / 8) x 128 B / 8
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        if (stk && pass == 1) continue;
        const uint32_t a = (pass == 1) ? ah_lo : ah_hi, b = ((pass == 2) ? bh_lo : bh_hi) + row;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks)
            tc_mma(region + C_NH, a + 8 * ks, make_desc(b + 256 * ks, 1024), idesc_f16(96), (pass | ks) != 0);
    }
}
// Input part of the second layer, W_ih2 h1 (K = 64): {R, Z} accumulate on top of the recurrent part, NI is written fresh
// by the first MMA (split in two for that) and accumulated by the rest.
template <int NKS = 4>
__device__ __forceinline__ void issue_X2(uint32_t region, uint32_t ah_hi, uint32_t ah_lo, uint32_t bx_hi, uint32_t bx_lo, int jh, bool stk) {
    const uint32_t row = (uint32_t)jh * 12288u;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        if (stk && pass == 1) continue;
        const uint32_t a = (pass == 1) ? ah_lo : ah_hi, b = ((pass == 2) ? bx_lo : bx_hi) + row;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            if ((pass | ks) == 0) {
                tc_mma(region + C_R, a, make_desc(b, 1024), idesc_f16(64), 1);
                tc_mma(region + C_NI, a, make_desc(b + 8192, 1024), idesc_f16(32), 0);
            } else {
                tc_mma(region + C_R, a + 8 * ks, make_desc(b + 256 * ks, 1024), idesc_f16(96), 1);
            }
        }
    }
}
// Input part of the first layer, W_ih1 x (K = 16, A operand x in shared memory).
__device__ __forceinline__ void issue_X1(uint32_t region, uint32_t ax_hi, uint32_t ax_lo, uint32_t bx_hi, uint32_t bx_lo, int jh, bool stk) {
    const uint32_t row = (uint32_t)jh * 3072u;    // 96 rows x (16 / 8) x 128 B / 8
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        if (stk && pass == 1) continue;
        const uint64_t a = make_desc((pass == 1) ? ax_lo : ax_hi, 256);
        const uint32_t b = ((pass == 2) ? bx_lo : bx_hi) + row;
        if (pass == 0) {
            tc_mma_ss(region + C_R, a, make_desc(b, 256), idesc_f16(64), 1);
            tc_mma_ss(region + C_NI, a, make_desc(b + 2048, 256), idesc_f16(32), 0);
        } else {
            tc_mma_ss(region + C_R, a, make_desc(b, 256), idesc_f16(96), 1);
        }
    }
}
// Linear output layer W_out h2 -> 16 columns at `dst`.
template <int NKS = 4>
__device__ __forceinline__ void issue_OUT(uint32_t dst, uint32_t ah_hi, uint32_t ah_lo, uint32_t b_hi, uint32_t b_lo, bool stk) {
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        if (stk && pass == 1) continue;
        const uint32_t a = (pass == 1) ? ah_lo : ah_hi, b = (pass == 2) ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks)
            tc_mma(dst, a + 8 * ks, make_desc(b + 256 * ks, 1024), idesc_f16(16), (pass | ks) != 0);
    }
}

__device__ __forceinline__ F2 f2bits(uint32_t a, uint32_t b) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ float ex2_f(float t) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo_v, float hi_v) {   // two floats -> fp16x2, round to nearest
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_v), "f"(lo_v));
    return d;
}

// GRU non-linearities (torch GRUCell, gate order r, z, n) of 8 hidden units x 16 HALVES rollouts, by one warp: thread t
// takes the pair of units 2 (t % 4), + 1 of the rollouts (tensor-memory lanes) t / 4 and 8 + t / 4 of each 16-lane half
// (the .16x256b / .16x128b access shapes above).  Reads the job's region {NH, R, Z, NI} at column offset `cu` (= first
// unit inside the half layer) and h(t-1) (the A operand itself), writes h(t) as the new A operand.  u0 = index of the
// first of the 8 units inside the layer (constants, operand columns); tq = tensor-memory address of the warp's lane quarter.
// Two units per step in packed FP32 (FFMA2 / FMUL2 / FADD2: the FMA-pipe work of the epilogue halves, which leaves the 4
// MUFU operations per unit -- 3 exponentials and, shared between two units, 2 reciprocals: 1/a = (b c d) / (a b c d) --
// as its bound: 16 MUFU lanes per clock and SM).  c = 1 / (128 S) undoes the operand scales (uniform per layer),
// cn = -c log2 e; per pair of units the constants are {-log2e (b_ir + b_hr), -log2e (b_iz + b_hz), b_in, b_hn} x 2.
// Exponents are capped at 2^30, which bounds the shared-reciprocal products; everything stays in the operand scale (h is
// kept as 128 h).
// N chunks of 16 rollouts x 8 units per call, interleaved for instruction-level parallelism (2 N independent dependence
// chains per thread): UNITS -> the chunks are consecutive 8-unit groups of the same 16 rollouts (64 live rollouts per CTA),
// else -> the same 8 units of the two 16-lane halves of the quarter (128 live rollouts).
// ldbar != 0: mbarrier that counts this warp once its tensor-memory loads have completed, i.e. before the arithmetic.
// STACK: 32 live rollouts per CTA, 8 per lane quarter (rows t / 4), and the dead rows 8 + t / 4 put to work: the A operand
// holds the fp16 hi part of a rollout's hidden state in row r and the lo part in row r + 8, so ONE MMA per weight part
// computes hi * W in row r and lo * W in row r + 8 -- two passes (W_hi, W_lo) instead of three, and the dropped lo * lo term
// comes for free.  The 16-lane access shapes hand a thread exactly that pair of rows: the two accumulator rows are added
// here, and one 16x128b access moves the {hi, lo} operand pair (no separate lo region).
template <int N, bool UNITS, bool STACK = false>
__device__ __forceinline__ void gru_epilogue(uint32_t tq, uint32_t region, uint32_t cu, const float *cst, float c, float cn,
                                              uint32_t c_hi, uint32_t c_lo, int u0, int lane, uint32_t ldbar = 0) {
    uint32_t R[N][4], Z[N][4], NI[N][4], NH[N][4], PH[N][2], PL[N][2];
    float4 k0[N], k1[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const uint32_t th = tq + ((uint32_t)(UNITS ? 0 : 16 * j) << 16), cj = cu + (UNITS ? 8 * j : 0);
        const int uj = u0 + (UNITS ? 8 * j : 0);
        ld16x256(th + region + C_R + cj, R[j]);
        ld16x256(th + region + C_Z + cj, Z[j]);
        ld16x256(th + region + C_NI + cj, NI[j]);
        ld16x256(th + region + C_NH + cj, NH[j]);
        ld16x128(th + c_hi + (uj >> 1), PH[j]);
        if (!STACK) ld16x128(th + c_lo + (uj >> 1), PL[j]);
        const float *kp = cst + ((uj >> 1) + (lane & 3)) * 8;
        k0[j] = *reinterpret_cast<const float4 *>(kp);       // brn0, brn1, bzn0, bzn1
        k1[j] = *reinterpret_cast<const float4 *>(kp + 4);   // bni0, bni1, bnh0, bnh1
    }
    const F2 C2 = f2(c), CN2 = f2(cn), ONE = f2(1.0f);
    ld_wait();
    if (ldbar) {   // the region is in registers: the issuer may overwrite it (one arrival per warp)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(ldbar);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int q = 0; q < (STACK ? 1 : 2); ++q) {   // the two rollouts of this thread in the chunk (STACK: one, in two rows)
            const F2 aR = STACK ? add2(f2bits(R[j][0], R[j][1]), f2bits(R[j][2], R[j][3])) : f2bits(R[j][2 * q], R[j][2 * q + 1]);
            const F2 aZ = STACK ? add2(f2bits(Z[j][0], Z[j][1]), f2bits(Z[j][2], Z[j][3])) : f2bits(Z[j][2 * q], Z[j][2 * q + 1]);
            const F2 aNH = STACK ? add2(f2bits(NH[j][0], NH[j][1]), f2bits(NH[j][2], NH[j][3])) : f2bits(NH[j][2 * q], NH[j][2 * q + 1]);
            const F2 aNI = STACK ? add2(f2bits(NI[j][0], NI[j][1]), f2bits(NI[j][2], NI[j][3])) : f2bits(NI[j][2 * q], NI[j][2 * q + 1]);
            const F2 tr = fma2(aR, CN2, f2(k0[j].x, k0[j].y));    // -log2e * pre-activation
            const F2 tz = fma2(aZ, CN2, f2(k0[j].z, k0[j].w));
            const F2 AR = add2(f2(ex2_f(fminf(lo(tr), 30.0f)), ex2_f(fminf(hi(tr), 30.0f))), ONE);   // 1 + e^{-r}
            const F2 AZ = add2(f2(ex2_f(fminf(lo(tz), 30.0f)), ex2_f(fminf(hi(tz), 30.0f))), ONE);
            const F2 PA = mul2(AR, AZ);
            const float inv = rcp_f(lo(PA) * hi(PA));
            const F2 IAB = mul2(f2(hi(PA), lo(PA)), f2(inv));          // 1 / (ar az) of each unit
            const F2 R2 = mul2(AZ, IAB), Z2 = mul2(AR, IAB);           // sigmoids
            const F2 tnh = fma2(aNH, C2, f2(k1[j].z, k1[j].w));
            const F2 tni = fma2(aNI, C2, f2(k1[j].x, k1[j].y));
            const F2 ta = mul2(fma2(R2, tnh, tni), f2(2.885390081777927f));               // 2 log2e * n pre-activation
            const F2 E = add2(f2(ex2_f(fminf(lo(ta), 30.0f)), ex2_f(fminf(hi(ta), 30.0f))), ONE);   // 1 + e^{2n}
            const float m2 = -256.0f * rcp_f(lo(E) * hi(E));
            const F2 N128 = fma2(f2(hi(E), lo(E)), f2(m2), f2(128.0f));   // 128 tanh = 128 - 256 / (1 + e^{2n})
            const float2 hh_ = unpack_h2(PH[j][STACK ? 0 : q]), hl = unpack_h2(STACK ? PH[j][1] : PL[j][q]);
            const F2 HS = add2(f2(hh_.x, hh_.y), f2(hl.x, hl.y));        // 128 h(t-1)
            const F2 HN = fma2(HS, Z2, fma2(neg2(N128), Z2, N128));      // 128 h(t) = 128 ((h - n) z + n)
            const uint32_t p_hi = pack_f16x2(lo(HN), hi(HN));
            const float2 hf = unpack_h2(p_hi);
            const F2 L = add2(HN, f2(-hf.x, -hf.y));
            const uint32_t p_lo = pack_f16x2(lo(L), hi(L));
            if (STACK) { PH[j][0] = p_hi; PH[j][1] = p_lo; }
            else { PH[j][q] = p_hi; PL[j][q] = p_lo; }
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const uint32_t th = tq + ((uint32_t)(UNITS ? 0 : 16 * j) << 16);
        const int uj = u0 + (UNITS ? 8 * j : 0);
        st16x128(th + c_hi + (uj >> 1), PH[j]);
        if (!STACK) st16x128(th + c_lo + (uj >> 1), PL[j]);
    }
}

// Optional pipeline trace (-DCPS_TC_TRACE): cycle stamps of one step of block 0, printed by the issuer / one epilogue
// thread / one row thread.  Used to find what a step waits for.
#ifdef CPS_TC_TRACE
#define TC_TRACE_STEP 10
#define TC_TR(i) do { if (blockIdx.x == 0 && lane == 0 && t == TC_TRACE_STEP) s_tr[i] = clock64(); } while (0)
#else
#define TC_TR(i)
#endif

