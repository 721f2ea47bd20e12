// cps_internal.cuh -- declarations shared by the translation units of libcps_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/cps.h"
#include "cps_device.cuh"

using namespace cps;

// =====================================================================================================
// kernel argument blocks (passed by value: they live in the constant bank)
// =====================================================================================================
struct MppiArgs {
    OdeParams ode;
    CostParams cost;
    MppiParams mp;
    SolveIO io;
    float s_inline[6];   // cps_mppi_step_host: the state travels in the parameter block (no host-to-device copy)
    int use_inline;
};

typedef void (*mppi_fn)(const MppiArgs);
static inline int sc_mode(unsigned flags) {
    if (flags & CPS_FLAG_SUBSTEP_SINCOS) return (flags & CPS_FLAG_FAST_SINCOS) ? SC_MUFU : SC_ACCURATE;
    return SC_ROTATE;
}

struct RolloutArgs {
    OdeParams ode;
    const float *s0;
    long long ss_b;          // 6 if batched, 0 if one shared state
    const float *Q;
    long long qs_b, qs_t;
    int B, T;
    float *traj_out;
    long long ts_k, ts_t, ts_c;
    float *final_out;        // [B][6] or null
};

struct CostArgs {
    CostParams cost;
    const float *traj;       // [K][rows][6], rows = T+1 (or T for get_stage_cost on states[:, :-1])
    const float *Q;          // [K][T]
    float u_prev;
    int K, T, rows;
    float inv_T1;
    float *J;                // [K] or null
    float *stage;            // [K][T] or null
    int unshifted;
};

struct FinalizeArgs {
    MppiParams mp;
    const float *partials;   // [n_parts][2 + n_red]
    int n_parts;
    float *u_nom;
    float *u_out;
    float *shard_out;
};


struct NetState;    // cps_net.cu
struct FleetState;  // cps_fleet.cu
struct PlanState;   // cps_plan.cu
struct GmmState;    // cps_gmm.cu
struct GradState;   // cps_grad.cu

struct cps_handle {
    cps_config cfg;
    int n_ind, n_red;
    float phys[CPS_PH_COUNT];
    float cost_in[24];
    int cost_in_n;
    float mppi_in[7];  // cc_weight, R, LBD, NU, sigma, lo, hi
    float target_position, target_equilibrium, L_var, m_pole_var;
    OdeParams ode;
    CostParams cost;
    MppiParams mp;
    cudaStream_t stream;
    // scratch owned by the handle
    float *d_partials;
    unsigned *d_ticket;
    int *d_nonfinite;
    float *d_s, *d_unom, *d_u;
    float *d_uprev;   // legacy front-end: previous nominal sequence [T]
    float *d_ldu;     // legacy front-end: staging of delta_u for cps_legacy_step_host [K][T]
    float *d_lknots;  // legacy front-end: knot draws of the interpolated sampler [K][n_knots]
    size_t n_lknots;  // bytes allocated at d_lknots
    float *h_pin;  // pinned: [0..6) s, [8] u
    float *h_pin_dev;        // device view of h_pin (mapped): the solve writes u straight into host memory
    const float *inline_s;   // set by cps_mppi_step_host around its cps_mppi_step call
    int grid, block;
    size_t smem;
    int shard;
    float *shard_out;
    // growable buffers of cps_rollout_host
    float *d_rs0, *d_rQ, *d_rtraj, *d_rfinal;
    size_t cap_rs0, cap_rQ, cap_rtraj, cap_rfinal;
    // cps_rollout_host pipeline: copy-in / copy-out streams and per-chunk events (created on first use)
    cudaStream_t st_in, st_out;
    cudaEvent_t ev_start, ev_in[16], ev_k[16];
    int pipe_ready;
    long long launches;
    int net_last_kernel;       // 0 none, 1 net_kernel (FP32), 2 net_tc_kernel
    int rollout_last_kernel;   // 0 none, 1 rollout_kernel, 2 rollout_pair_kernel
    PeerExchange px;           // cps_mppi_set_peers (world <= 1: off); px.epoch = solves exchanged so far
    int *d_px_timeouts;
    std::string err;
    NetState *net;      // neural predictor (cps_net_load), owned
    FleetState *fleet;  // closed-loop experiments (cps_fleet_create), owned
    PlanState *plan;    // forward-only planners (cps_plan_*, cps_cem_*), owned
    GmmState *gmm;      // CEM with a Gaussian-mixture sampling distribution (cps_cem_gmm_*), owned
    GradState *grad;    // adjoint workspace and Adam moments (cps_plan_cost_grad, cps_rpgd_*), owned
};

extern thread_local std::string g_create_err;

// Dynamic shared memory (in floats) of mppi_solve_block / mppi_solve_block2 for a block of `block` threads carrying
// `per_thread` rollouts each: [T] shifted nominal inputs, 2 [p] tent weights, reduction scratch, `extra` floats the
// caller appends (fleet: Philox draws) and, for the MAX_COST plugins, the row-sum slots (cps_device.cuh RowSumPlan).
// Sets mp.rs_off.
static inline size_t mppi_smem_floats(MppiParams &mp, int cost_id, int block, int per_thread, size_t extra = 0) {
    size_t fl = (size_t)mp.T + 2 * (size_t)mp.p + (size_t)(block / 32) * (mp.n_red + 2) + (size_t)mp.n_red + 4;
    fl = ((fl + 1) & ~(size_t)1) + ((extra + 1) & ~(size_t)1);
    mp.rs_off = (int)fl;
    if (cost_id == CPS_COST_DEFAULT || cost_id == CPS_COST_QUADRATIC_BOUNDARY)
        fl += (size_t)row_sum_slots(mp.T + 1) * block * per_thread;
    return fl;
}

// cps_fleet.cu
void cps_fleet_free(cps_handle *h);
// cps_plan.cu
void cps_plan_free(cps_handle *h);
// cps_gmm.cu
void cps_gmm_free(cps_handle *h);
// cps_grad.cu
void cps_grad_free(cps_handle *h);
// cps_lib.cu: cost parameters folded for a given target equilibrium
int cps_fold_cost_for(cps_handle *h, float target_equilibrium, CostParams *out);
// cps_net.cu
void cps_net_free(cps_handle *h);
int cps_net_mppi_step(cps_handle *h, const float *s_dev, const float *noise_dev, int noise_layout, float u_prev,
                      float *u_nom_dev, float *u_out_dev, float *J_out_dev, float *traj_out_dev, int traj_layout,
                      float *u_run_out_dev);

static inline int fail(cps_handle *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_err = buf;
    return code;
}

#define CUDA_TRY(h, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) return fail(h, CPS_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

