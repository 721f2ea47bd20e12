// cps_mppi_inst.cuh -- the MPPI solve kernels (K1 + K2 of the design: rollouts, cost, block partials, last-block merge) and
// their dispatch for ONE cost plugin.  Included by cps_mppi_<plugin>.cu with CPS_MPPI_COST / CPS_MPPI_NAME defined: one
// translation unit per plugin keeps the build parallel (40 + 2 kernels each).
#include <cuda_runtime.h>

#include "cps_internal.cuh"

using namespace cps;

// NSUB = 10: the substeps of the predictors' operating point unrolled (control_step); 0: a.ode.n substeps in a loop.
template <int INTEG, int COST, int SC, int NOISE, bool FAST_DIV, bool EXACT_ATAN2, int NSUB = 0>
__global__ void __launch_bounds__(256, 4) mppi_kernel(const __grid_constant__ MppiArgs a) {
    extern __shared__ float smem[];
    SolveIO io = a.io;
    if (a.use_inline) io.s = a.s_inline;   // constant-bank reads instead of a global load of the state
    mppi_solve_block<INTEG, COST, SC, NOISE, FAST_DIV, EXACT_ATAN2, NSUB>(a.ode, a.cost, a.mp, io, smem, blockIdx.x, gridDim.x);
}

// The throughput form: two rollouts per thread in packed FP32 (mppi_solve_block2); chosen by cps_mppi_step for large K.
template <int INTEG, int COST, int NSUB = 0>
__global__ void __launch_bounds__(128, 4) mppi_pair_kernel(const __grid_constant__ MppiArgs a) {
    extern __shared__ float smem[];
    SolveIO io = a.io;
    if (a.use_inline) io.s = a.s_inline;
    mppi_solve_block2<INTEG, COST, NSUB>(a.ode, a.cost, a.mp, io, smem, blockIdx.x, gridDim.x);
}

template <int INTEG, int COST, int SC, int NOISE>
static mppi_fn pick_mppi3(unsigned flags) {
    const bool fd = flags & CPS_FLAG_FAST_DIV, ea = flags & CPS_FLAG_EXACT_ATAN2;
    if (fd) return ea ? mppi_kernel<INTEG, COST, SC, NOISE, true, true> : mppi_kernel<INTEG, COST, SC, NOISE, true, false>;
    return ea ? mppi_kernel<INTEG, COST, SC, NOISE, false, true> : mppi_kernel<INTEG, COST, SC, NOISE, false, false>;
}
template <int INTEG, int NOISE>
static mppi_fn pick_mppi2b(unsigned flags, bool n10) {
    constexpr int COST = CPS_MPPI_COST;
    switch (sc_mode(flags)) {
    case SC_ACCURATE: return pick_mppi3<INTEG, COST, SC_ACCURATE, NOISE>(flags);
    case SC_MUFU: return pick_mppi3<INTEG, COST, SC_MUFU, NOISE>(flags);
    default:
        if (flags & CPS_FLAG_FAST_DIV) return mppi_kernel<INTEG, COST, SC_ROTATE, NOISE, true, false>;
        return n10 ? mppi_kernel<INTEG, COST, SC_ROTATE, NOISE, false, false, 10> : mppi_kernel<INTEG, COST, SC_ROTATE, NOISE, false, false>;
    }
}
template <int INTEG>
static mppi_fn pick_mppi2(int noise, unsigned flags, bool n10) {
    if (noise == CPS_NOISE_INDUCING) return pick_mppi2b<INTEG, CPS_NOISE_INDUCING>(flags, n10);
    return pick_mppi2b<INTEG, CPS_NOISE_DIRECT>(flags, n10);
}
#define CPS_CAT2(a, b) a##b
#define CPS_CAT(a, b) CPS_CAT2(a, b)
mppi_fn CPS_CAT(cps_pick_mppi_, CPS_MPPI_NAME)(int integ, int noise, unsigned flags, int n_sub) {
    return integ == CPS_EULER_V0 ? pick_mppi2<0>(noise, flags, n_sub == 10) : pick_mppi2<1>(noise, flags, n_sub == 10);
}
mppi_fn CPS_CAT(cps_pick_mppi_pair_, CPS_MPPI_NAME)(int integ, int n_sub) {
    if (n_sub == 10) return integ == CPS_EULER_V0 ? mppi_pair_kernel<0, CPS_MPPI_COST, 10> : mppi_pair_kernel<1, CPS_MPPI_COST, 10>;
    return integ == CPS_EULER_V0 ? mppi_pair_kernel<0, CPS_MPPI_COST> : mppi_pair_kernel<1, CPS_MPPI_COST>;
}
