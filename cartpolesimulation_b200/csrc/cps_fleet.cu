// cps_fleet.cu -- E independent closed-loop experiments advanced together (BASELINE configs[4]: data_generator with
// 8192 MPPI-controlled cartpoles).
//
// One launch = one controller period of EVERY experiment:
//   grid (blocks_per_experiment, E).  All blocks of experiment e run its MPPI solve (mppi_solve_block, the same code
//   as mppi_kernel) on the experiment's own state, nominal inputs, last control and targets; the block that finishes
//   last merges the partials, obtains Q = u_nom[0], writes the experiment's record row and then integrates the PLANT
//   (CartPole.update_state, CartPole/__init__.py:283-324: Euler-Cromer at dt_simulation, edge bounce, cos/sin, fmod
//   wrap; second derivatives refreshed every tick) for one controller period in the reference's mixed
//   float32/float64 arithmetic, and stores the new state.  Nothing returns to the host between periods.
//
// Perturbations are either supplied ([E][n_ind][K] standard-normal draws per period -- the "identical injected noise"
// hook of the parity tests) or generated in the kernel: Philox4x32-10 keyed by the seed, counter = (rollout, draw group,
// period, GLOBAL experiment index) + Box-Muller, staged in shared memory so the integration loop keeps its registers.
// The streams depend only on the global experiment index, so results do not change with the sharding over GPUs.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>

#include "cps_internal.cuh"

struct PlantParams {
    float k, m_cart, g, J_fric, M_fric, u_max, thl;
    double L, m_pole, dt;
    int n_sim;
};

// Plant-side models between the plant and the controller (all OFF in the shipped configuration,
// cartpole_physical_parameters.yml:13-24): control disturbance (CartPole/noise_control_signal.py:5-26), measurement
// noise (CartPole/noise_adder.py:69-84), latency (CartPole/latency_adder.py:10-74).  Per plant tick the reference pushes
// the true state into a ring buffer, reads it back `latency` seconds late (linear interpolation between the two
// neighbouring ticks, all six components), adds the measurement noise and recomputes cos / sin; the controller is handed
// that observation, the plant keeps integrating the true state.
struct PlantModels {
    int on;                    // any of the three enabled: the solve reads `obs`, the plant stage maintains ring / obs
    int ctrl_mode;             // 0 OFF, 1 additive, 2 truncnorm
    float ctrl_mult, ctrl_add;
    int meas_on;
    float sig_a, sig_p, sig_aD, sig_pD;
    int lat_int;               // int(latency / dt_simulation)
    double lat_frac;           // latency / dt_simulation - lat_int
    int ring_len;              // lat_int + 2 slots of 6 floats per experiment
    float *obs;                // [E][8] the controller's view of the state for the next solve
    float *ring;               // [E][ring_len][6]
    long long pushed;          // states pushed per experiment before this launch
    const float *ctrl_draws;   // [E] standard normal (additive) / uniform (truncnorm) draw of this period, or null: Philox
    const float *meas_draws;   // [n_sim][E][4] standard normals of this period (angle, position, angleD, positionD), or null
};

struct FleetArgs {
    PlantModels pm;
    OdeParams ode;              // the controller's model (L, m_pole "for controller")
    CostParams cost_up, cost_dn;  // folded for target_equilibrium = +1 / -1 (quadratic_boundary_grad selects its set)
    MppiParams mp;
    PlantParams plant;
    int bpe;                    // blocks per experiment
    float *s;                   // [E][8] current state (6 used)
    float *u_nom;               // [E][T]
    float *u_prev;              // [E] last returned control
    const float *tp, *te;       // [E] targets of this period (null: 0 / +1)
    const float *noise;         // supplied draws [E][n_ind][K] of this period, or null (Philox)
    unsigned long long seed;
    unsigned period;            // controller period index (Philox counter)
    unsigned e_offset;          // global index of local experiment 0
    float *partials;            // [E][bpe][2 + n_red]
    unsigned *tickets;          // [E]
    int *nonfinite;
    float *record;              // [E][CPS_FLEET_RECORD] of this period, or null
    float *J_out;               // [E][K] or null
    double time;                // period * dt_control
    // relabel mode (cps_fleet_relabel): states come from a recording, the plant is not integrated
    const float *replay_s;      // [E][6] recorded states of this row, or null (closed loop)
    const float *L_row, *mp_row;  // [E] controller-model pole length / mass of this row, or null (handle's values)
    float *Q_out;               // [E] controls computed for this row
    const int *active;          // [E] or null: files with 0 sit this row out (controller state untouched, no output)
    float k, m_cart, g, J_fric, M_fric, u_max;  // physical constants for the device-side fold of (L, m_pole)
    float m_pole_fixed;         // ODE_v0 takes only L from the variable parameters
    float L_default, mp_default;  // the handle's L / m_pole "for controller" when only one of the arrays is given
    double h_step;                // substep dt / n in double, for the device-side fold
};

// fold_ode (cps_lib.cu) on the device, same double-precision expressions: the controller's model constants for a
// per-experiment pole length / mass (add_control_along_trajectories feeds L per recorded row).
template <int INTEG>
__device__ __forceinline__ void fold_ode_device(const FleetArgs &a, double L, double mp_in, OdeParams &o) {
    const double k = a.k, mc = a.m_cart, g = a.g, J = a.J_fric, M = a.M_fric;
    const double mp = (INTEG == 0) ? (double)a.m_pole_fixed : mp_in;
    const double kp1 = k + 1.0, Lh = L / 2.0;
    o.KM = (float)(kp1 * (mc + mp));
    o.m_p = (float)mp;
    o.c1 = (float)(mp * g);
    o.c2 = (float)(kp1 * mp * Lh);
    o.c3 = (float)(J / Lh);
    o.c5 = (float)(kp1 * M);
    o.d1 = (float)(g / (kp1 * Lh));
    o.d2 = (float)(1.0 / (kp1 * Lh));
    o.d3 = (float)(J / (mp * Lh * kp1 * Lh));
    o.bounce = (float)(2.0 / (0.5 * L));
    const double hh = a.h_step;
    o.hd1 = (float)(hh * (g / (kp1 * Lh)));
    o.hd2 = (float)(hh * (1.0 / (kp1 * Lh)));
    o.hd3 = (float)(hh * (J / (mp * Lh * kp1 * Lh)));
}

// ---- Philox4x32-10 (Salmon et al., SC'11), the counter-based generator torch / TF / cuRAND also use ---------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// Two standard normals from two 32-bit words (Box-Muller on 24-bit uniforms in (0, 1)).
__device__ __forceinline__ void box_muller(unsigned a, unsigned b, float &n0, float &n1) {
    const float u1 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-8f;  // 2^-24
    const float u2 = ((float)(b >> 8) + 0.5f) * 5.9604644775390625e-8f;
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    n0 = r * cs;
    n1 = r * sn;
}

// Draws i = 0..n_ind-1 of rollout k, experiment e_global, controller period `period`; dst[i * stride].
__device__ __forceinline__ void fleet_draws(unsigned long long seed, unsigned period, unsigned e_global, unsigned k,
                                            int n_ind, float *dst, long long stride) {
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    for (int g = 0; 4 * g < n_ind; ++g) {
        const uint4 r = philox4x32_10(make_uint4(k, (unsigned)g, period, e_global), key);
        float n[4];
        box_muller(r.x, r.y, n[0], n[1]);
        box_muller(r.z, r.w, n[2], n[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (4 * g + q < n_ind) dst[(long long)(4 * g + q) * stride] = n[q];
    }
}

// cps_fleet_noise: materialise the draws the kernel generates, [E][n_ind][K]
__global__ void __launch_bounds__(256) fleet_noise_kernel(unsigned long long seed, unsigned period, unsigned e_offset, int E,
                                                          int K, int n_ind, float *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, e = blockIdx.y;
    if (k >= K || e >= E) return;
    fleet_draws(seed, period, e_offset + (unsigned)e, (unsigned)k, n_ind, out + ((size_t)e * n_ind) * K + k, K);
}

// ---- the plant (device restatement of oracle/cps_oracle.c: ode_plant, plant_integrate; no FMA contraction) ------------
__device__ __forceinline__ void plant_ode(const PlantParams &P, float ca, float sa, float angleD, float positionD, float u,
                                          double &angleDD, double &positionDD) {
    const double kp1 = (double)__fadd_rn(P.k, 1.0f);
    const float ca2 = __fmul_rn(ca, ca), aD2 = __fmul_rn(angleD, angleD);
    const double A = __dsub_rn(__dmul_rn(kp1, __dadd_rn((double)P.m_cart, P.m_pole)), __dmul_rn(P.m_pole, (double)ca2));
    const float F_fric = __fmul_rn(-P.M_fric, positionD);
    const float T_fric = __fmul_rn(-P.J_fric, angleD);
    const double L_half = P.L / 2.0;
    const double grav = __dmul_rn(__dmul_rn(__dmul_rn(P.m_pole, (double)P.g), (double)sa), (double)ca);
    const double fric = (double)__fmul_rn(T_fric, ca) / L_half;
    const double centr = -__dmul_rn(__dmul_rn(__dmul_rn(P.m_pole, L_half), (double)aD2), (double)sa);
    const double inner = __dadd_rn(__dadd_rn(centr, (double)F_fric), (double)u);
    const double pDD = __dadd_rn(__dadd_rn(grav, fric), __dmul_rn(kp1, inner)) / A;
    positionDD = pDD;
    const double num = __dadd_rn(__dadd_rn((double)__fmul_rn(P.g, sa), __dmul_rn(pDD, (double)ca)),
                                 (double)T_fric / __dmul_rn(P.m_pole, L_half));
    angleDD = num / __dmul_rn(kp1, L_half);
}

__device__ __forceinline__ double plant_wrap(double angle) {
    const double two_pi = 6.283185307179586, pi = 3.141592653589793;
    const double m = fmod(angle, two_pi);
    if (m < -pi) return m + two_pi;
    if (m > pi) return m - two_pi;
    return m;
}

__device__ __forceinline__ void plant_tick(const PlantParams &P, float *s, double aDD, double pDD) {
    const float angle = s[IDX_ANGLE], angleD = s[IDX_ANGLED], position = s[IDX_POS], positionD = s[IDX_POSD];
    const double angleD_next = __dadd_rn((double)angleD, __dmul_rn(aDD, P.dt));
    const double positionD_next = __dadd_rn((double)positionD, __dmul_rn(pDD, P.dt));
    const double angle_next = __dadd_rn((double)angle, __dmul_rn(angleD_next, P.dt));
    const double position_next = __dadd_rn((double)position, __dmul_rn(positionD_next, P.dt));
    s[IDX_ANGLE] = (float)angle_next; s[IDX_ANGLED] = (float)angleD_next;
    s[IDX_POS] = (float)position_next; s[IDX_POSD] = (float)positionD_next;
    if (s[IDX_POS] >= P.thl || -s[IDX_POS] >= P.thl) {  // edge_bounce (cartpole_equations.py:341-347)
        const float a = s[IDX_ANGLE], aD = s[IDX_ANGLED], p = s[IDX_POS], pD = s[IDX_POSD];
        const float ca = cosf(a);
        const double aD2 = __dsub_rn((double)aD, __dmul_rn(2.0, (double)__fmul_rn(pD, ca)) / __dmul_rn(0.5, P.L));
        const double a2 = __dadd_rn((double)a, __dmul_rn(aD2, P.dt));
        const float pD2 = -pD;
        const double p2 = __dadd_rn((double)p, __dmul_rn((double)pD2, P.dt));
        s[IDX_ANGLE] = (float)a2; s[IDX_ANGLED] = (float)aD2; s[IDX_POS] = (float)p2; s[IDX_POSD] = pD2;
    }
    s[IDX_COS] = cosf(s[IDX_ANGLE]);   // of the UNWRAPPED angle (CartPole/__init__.py:329-334)
    s[IDX_SIN] = sinf(s[IDX_ANGLE]);
    s[IDX_ANGLE] = (float)plant_wrap((double)s[IDX_ANGLE]);
}

// The controller's view after `pushed` ring pushes: delayed + interpolated state (latency_adder.py:67-71), measurement
// noise (noise_adder.py:69-84, draws n[0..3] = angle, position, angleD, positionD; null: none) and the cos / sin refresh of
// update_vertical_angle_offset with a zero offset (CartPole/__init__.py:348-358).  float64 like the reference's buffer.
__device__ __forceinline__ void plant_observe(const PlantModels &M, const float *ring, long long pushed, const float *n,
                                              float *obs) {
    const int RL = M.ring_len;
    const int i1 = (int)(((pushed - 1 - M.lat_int) % RL + RL) % RL), i2 = (int)(((pushed - 2 - M.lat_int) % RL + RL) % RL);
    double v[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const double s1 = (double)ring[i1 * 6 + c], s2 = (double)ring[i2 * 6 + c];
        v[c] = __dadd_rn(s1, __dmul_rn(M.lat_frac, __dsub_rn(s2, s1)));
    }
    if (M.meas_on && n) {
        v[IDX_ANGLE] = plant_wrap(__dadd_rn(v[IDX_ANGLE], (double)__fmul_rn(M.sig_a, n[0])));
        v[IDX_POS] = __dadd_rn(v[IDX_POS], (double)__fmul_rn(M.sig_p, n[1]));
        v[IDX_ANGLED] = __dadd_rn(v[IDX_ANGLED], (double)__fmul_rn(M.sig_aD, n[2]));
        v[IDX_POSD] = __dadd_rn(v[IDX_POSD], (double)__fmul_rn(M.sig_pD, n[3]));
    }
    v[IDX_ANGLE] = plant_wrap(v[IDX_ANGLE]);
    v[IDX_COS] = cos(v[IDX_ANGLE]);
    v[IDX_SIN] = sin(v[IDX_ANGLE]);
#pragma unroll
    for (int c = 0; c < 6; ++c) obs[c] = (float)v[c];
}

// Philox draws of the plant-side models: counter (0xFFFFFF00 + j, 0xFFFFFFFF, period, experiment) -- disjoint from the
// rollout draws, whose first word is a rollout index < K.  j = tick inside the period (measurement) or 0xFF (control).
__device__ __forceinline__ uint4 plant_philox(unsigned long long seed, unsigned period, unsigned e_global, unsigned j) {
    return philox4x32_10(make_uint4(0xFFFFFF00u + j, 0xFFFFFFFFu, period, e_global), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
}

// add_control_noise (CartPole/noise_control_signal.py:5-26): d = a standard normal (additive) or a uniform (truncnorm)
__device__ __forceinline__ float plant_control_noise(const PlantModels &M, float Q, float d) {
    if (M.ctrl_mode == 1) return __fadd_rn(__fadd_rn(Q, __fmul_rn(M.ctrl_mult, d)), M.ctrl_add);
    if (M.ctrl_mode == 2) {   // truncnorm.rvs((-1 - loc) / scale, (1 - loc) / scale, loc, scale): inverse-CDF of the draw
        const double scale = (double)M.ctrl_mult, loc = (double)Q + (double)M.ctrl_add;
        const double ca = normcdf((-1.0 - loc) / scale), cb = normcdf((1.0 - loc) / scale);
        const double z = normcdfinv(ca + (double)d * (cb - ca));
        return (float)fmin(fmax(loc + scale * z, -1.0), 1.0);
    }
    return Q;
}

// NSUB = 10 (packed solve only): the substeps of the predictors' operating point unrolled; 0: ode.n substeps in a loop.
template <int INTEG, int COST, bool PHILOX, bool PAIR, int NSUB = 0>
__global__ void __launch_bounds__(256, 4) fleet_kernel(const __grid_constant__ FleetArgs a) {
    extern __shared__ float smem[];
    __shared__ CostParams s_cost;
    const MppiParams &mp = a.mp;
    const int e = blockIdx.y, tid = threadIdx.x;
    if (a.active && a.active[e] == 0) return;   // whole experiment (all its blocks): warm start, last control and ticket stay
    const float tp = a.tp ? a.tp[e] : 0.0f, te = a.te ? a.te[e] : 1.0f;
    if (tid == 0) {
        s_cost = (te == 1.0f) ? a.cost_up : a.cost_dn;
        s_cost.target_position = tp;
        s_cost.target_equilibrium = te;
    }
    const int nwarps = blockDim.x >> 5;
    // [n_ind][rollouts of this block]; even offset: a pair's draws are read as one 8-byte word (layout: mppi_smem_floats)
    float *s_eps = smem + ((mp.T + 2 * mp.p + nwarps * (mp.n_red + 2) + mp.n_red + 4 + 1) & ~1);
    const int per_block = PAIR ? 2 * (int)blockDim.x : (int)blockDim.x;   // rollouts per block
    if (PHILOX) {
        if (PAIR) {
            const int k = 2 * (blockIdx.x * blockDim.x + tid);
            fleet_draws(a.seed, a.period, a.e_offset + (unsigned)e, (unsigned)min(k, mp.K - 2), mp.n_ind, s_eps + 2 * tid, per_block);
            fleet_draws(a.seed, a.period, a.e_offset + (unsigned)e, (unsigned)min(k + 1, mp.K - 1), mp.n_ind, s_eps + 2 * tid + 1, per_block);
        } else {
            const int k = blockIdx.x * blockDim.x + tid;
            fleet_draws(a.seed, a.period, a.e_offset + (unsigned)e, (unsigned)min(k, mp.K - 1), mp.n_ind, s_eps + tid, per_block);
        }
    }
    __syncthreads();

    SolveIO io;
    io.s = a.replay_s ? a.replay_s + (size_t)e * 6 : ((a.pm.on ? a.pm.obs : a.s) + (size_t)e * 8);
    if (PHILOX) {  // rollout k of the experiment sits at s_eps[i * per_block + (k - first rollout of this block)]
        io.noise = s_eps - (long long)blockIdx.x * per_block;
        io.ns_i = per_block; io.ns_k = 1;
    } else {
        io.noise = a.noise + (size_t)e * mp.n_ind * mp.K;
        io.ns_i = mp.K; io.ns_k = 1;
    }
    io.u_prev = a.u_prev[e];
    io.u_nom = a.u_nom + (size_t)e * mp.T;
    io.u_out = a.u_prev + e;   // the returned control becomes the next period's u_prev (optimizer_mppi.py:210)
    io.J_out = a.J_out ? a.J_out + (size_t)e * mp.K : nullptr;
    io.traj_out = nullptr; io.ts_k = io.ts_t = io.ts_c = 0; io.u_run_out = nullptr;
    io.partials = a.partials + (size_t)e * a.bpe * (mp.n_red + 2);
    io.ticket = a.tickets + e;
    io.nonfinite = a.nonfinite;
    io.shard_out = nullptr;
    io.px.world = 0;
    OdeParams ode = a.ode;
    if (a.L_row || a.mp_row)
        fold_ode_device<INTEG>(a, a.L_row ? (double)a.L_row[e] : (double)a.L_default,
                               a.mp_row ? (double)a.mp_row[e] : (double)a.mp_default, ode);
    const bool last = PAIR ? mppi_solve_block2<INTEG, COST, NSUB, false>(ode, s_cost, mp, io, smem, blockIdx.x, a.bpe)
                           : mppi_solve_block<INTEG, COST, SC_ROTATE, CPS_NOISE_INDUCING, false, false, 0, false>(ode, s_cost, mp, io, smem,
                                                                                                       blockIdx.x, a.bpe);
    if (!last) return;
    __syncthreads();
    if (tid != 0) return;
    if (a.replay_s) {   // relabel: the control is the result; the next row brings its own state
        a.Q_out[e] = *io.u_out;
        return;
    }

    // ---- the plant: one controller period with Q held (CartPole/__init__.py:283-324, 475-527) -------------------------
    const PlantParams &P = a.plant;
    const PlantModels &M = a.pm;
    float s[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = a.s[(size_t)e * 8 + c];   // the TRUE state (io.s is the controller's view when M.on)
    const float Qc = *io.u_out;   // written by this thread in merge_and_finish
    float Q = Qc;
    if (M.on && M.ctrl_mode) {
        float d;
        if (M.ctrl_draws) d = M.ctrl_draws[e];
        else {
            const uint4 r = plant_philox(a.seed, a.period, a.e_offset + (unsigned)e, 0xFFu);
            float n0, n1;
            box_muller(r.x, r.y, n0, n1);
            d = (M.ctrl_mode == 1) ? n0 : ((float)(r.z >> 8) + 0.5f) * 5.9604644775390625e-8f;
        }
        Q = plant_control_noise(M, Qc, d);
    }
    const float u = __fmul_rn(P.u_max, Q);
    double aDD, pDD;
    plant_ode(P, s[IDX_COS], s[IDX_SIN], s[IDX_ANGLED], s[IDX_POSD], u, aDD, pDD);
    if (a.record) {
        float *r = a.record + (size_t)e * CPS_FLEET_RECORD;
        r[0] = (float)a.time;
        r[1] = s[IDX_ANGLE]; r[2] = s[IDX_ANGLED]; r[3] = (float)aDD; r[4] = s[IDX_COS]; r[5] = s[IDX_SIN];
        r[6] = s[IDX_POS]; r[7] = s[IDX_POSD]; r[8] = (float)pDD;
        r[9] = Qc; r[10] = Q; r[11] = u; r[12] = tp; r[13] = te; r[14] = 0.0f; r[15] = 0.0f;
    }
    float *ring = M.on ? M.ring + (size_t)e * M.ring_len * 6 : nullptr;
    for (int i = 0; i < P.n_sim; ++i) {
        plant_tick(P, s, aDD, pDD);
        plant_ode(P, s[IDX_COS], s[IDX_SIN], s[IDX_ANGLED], s[IDX_POSD], u, aDD, pDD);
        if (M.on) {   // add_current_state_to_latency_buffer (latency_adder.py:36-47)
            float *slot = ring + (size_t)((M.pushed + i) % M.ring_len) * 6;
#pragma unroll
            for (int c = 0; c < 6; ++c) slot[c] = s[c];
        }
    }
    float *so = a.s + (size_t)e * 8;
#pragma unroll
    for (int c = 0; c < 6; ++c) so[c] = s[c];
    if (M.on) {   // what the controller will see at the start of the next period: the observation made on the last tick
        float n[4];
        const float *np_ = nullptr;
        if (M.meas_on) {
            if (M.meas_draws) {
#pragma unroll
                for (int q = 0; q < 4; ++q) n[q] = M.meas_draws[((size_t)(P.n_sim - 1) * gridDim.y + e) * 4 + q];
            } else {
                const uint4 r = plant_philox(a.seed, a.period, a.e_offset + (unsigned)e, (unsigned)(P.n_sim - 1));
                box_muller(r.x, r.y, n[0], n[1]);
                box_muller(r.z, r.w, n[2], n[3]);
            }
            np_ = n;
        }
        plant_observe(M, ring, M.pushed + P.n_sim, np_, M.obs + (size_t)e * 8);
    }
}

// (Re)start of the observation chain: ring = the reference's initial buffer (zeros, cos = 1, latency_adder.py:26-28), nothing
// pushed yet; the first solve sees the TRUE state, as the controller call of set_cartpole_state_at_t0 does
// (CartPole/__init__.py:869-880: self.s, no noise, no latency).
__global__ void fleet_observe_reset_kernel(PlantModels M, const float *s, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float *ring = M.ring + (size_t)e * M.ring_len * 6;
    for (int j = 0; j < M.ring_len; ++j)
        for (int c = 0; c < 6; ++c) ring[j * 6 + c] = (c == IDX_COS) ? 1.0f : 0.0f;
    for (int c = 0; c < 8; ++c) M.obs[(size_t)e * 8 + c] = s[(size_t)e * 8 + c];
}

// ---- host side ---------------------------------------------------------------------------------------------------------
struct FleetState {
    cps_fleet_config cfg;
    int E, bpe, block, pair;   // pair: two rollouts per thread (packed FP32), K even
    size_t smem;
    int rs_off;                // MppiParams::rs_off for this geometry
    float *d_s, *d_unom, *d_uprev, *d_partials;
    unsigned *d_tickets;
    long long period;
    CostParams cost_up, cost_dn;
    PlantModels pm;            // plant-side models (cps_fleet_set_plant_models); pm.pushed counts the ring pushes
};

void cps_fleet_free(cps_handle *h) {
    FleetState *F = h->fleet;
    if (!F) return;
    cudaFree(F->d_s); cudaFree(F->d_unom); cudaFree(F->d_uprev); cudaFree(F->d_partials); cudaFree(F->d_tickets);
    cudaFree(F->pm.obs); cudaFree(F->pm.ring);
    delete F;
    h->fleet = nullptr;
}

typedef void (*fleet_fn)(const FleetArgs);
template <int INTEG, bool PHILOX, bool PAIR, int NSUB>
static fleet_fn pick_fleet2(int cost) {
    switch (cost) {
    case CPS_COST_DEFAULT: return fleet_kernel<INTEG, COST_DEFAULT, PHILOX, PAIR, NSUB>;
    case CPS_COST_QUADRATIC_BOUNDARY: return fleet_kernel<INTEG, COST_QB, PHILOX, PAIR, NSUB>;
    case CPS_COST_QB_GRAD_MINIMAL: return fleet_kernel<INTEG, COST_GRADMIN, PHILOX, PAIR, NSUB>;
    case CPS_COST_QB_GRAD: return fleet_kernel<INTEG, COST_GRAD, PHILOX, PAIR, NSUB>;
    default: return nullptr;
    }
}
template <int INTEG>
static fleet_fn pick_fleet1(int cost, bool philox, bool pair, bool n10) {
    if (pair && n10) return philox ? pick_fleet2<INTEG, true, true, 10>(cost) : pick_fleet2<INTEG, false, true, 10>(cost);
    if (pair) return philox ? pick_fleet2<INTEG, true, true, 0>(cost) : pick_fleet2<INTEG, false, true, 0>(cost);
    return philox ? pick_fleet2<INTEG, true, false, 0>(cost) : pick_fleet2<INTEG, false, false, 0>(cost);
}
static fleet_fn pick_fleet(int integ, int cost, bool philox, bool pair, bool n10) {
    return integ == CPS_EULER_V0 ? pick_fleet1<0>(cost, philox, pair, n10) : pick_fleet1<1>(cost, philox, pair, n10);
}

extern "C" int cps_fleet_create(cps_handle *h, const cps_fleet_config *cfg) {
    if (!h) return CPS_ERR_INVALID;
    if (!cfg || cfg->struct_size != (int)sizeof(cps_fleet_config))
        return fail(h, CPS_ERR_INVALID, "cps_fleet_create: cps_fleet_config size mismatch");
    if (cfg->n_experiments < 1 || cfg->sim_substeps < 1 || !(cfg->dt_simulation > 0.0))
        return fail(h, CPS_ERR_INVALID, "cps_fleet_create: need n_experiments >= 1, sim_substeps >= 1, dt_simulation > 0");
    if (cfg->noise_source != CPS_FLEET_NOISE_SUPPLIED && cfg->noise_source != CPS_FLEET_NOISE_PHILOX)
        return fail(h, CPS_ERR_INVALID, "cps_fleet_create: unknown noise source %d", cfg->noise_source);
    if (h->cfg.integrator == CPS_PREDICTOR_NEURAL)
        return fail(h, CPS_ERR_UNSUPPORTED, "cps_fleet_create: fleets run with the ODE predictors only");
    if (h->cfg.cost_id == CPS_COST_NONE) return fail(h, CPS_ERR_NOT_CONFIGURED, "cps_fleet_create: the handle has no cost function");
    if (h->cfg.noise_mode != CPS_NOISE_INDUCING)
        return fail(h, CPS_ERR_UNSUPPORTED, "cps_fleet_create: fleets use inducing-point noise (CPS_NOISE_INDUCING)");
    if (cfg->n_experiments > 65535) return fail(h, CPS_ERR_INVALID, "cps_fleet_create: at most 65535 experiments per handle");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    cps_fleet_free(h);
    FleetState *F = new (std::nothrow) FleetState();
    if (!F) return fail(h, CPS_ERR_INVALID, "cps_fleet_create: out of host memory");
    memset(F, 0, sizeof(*F));
    F->cfg = *cfg;
    F->E = cfg->n_experiments;
    const int K = h->cfg.num_rollouts, T = h->cfg.horizon;
    // large fleets are throughput work: two rollouts per thread in packed FP32 (mppi_solve_block2) from 65536 rollouts
    // per launch on, as for a single solve (tools/ab_fleet_pairs.py: 14 % faster at E = 256 and 1024, 30 % slower at E <= 16)
    F->pair = (K % 2 == 0 && K >= 64 && (long long)cfg->n_experiments * K >= CPS_MPPI_PAIR_MIN_ROLLOUTS &&
               !(h->cfg.flags & CPS_FLAG_NO_PAIRS)) ? 1 : 0;
    const int threads = F->pair ? K / 2 : K;   // threads per experiment
    F->block = threads >= 256 ? 256 : ((threads + 31) / 32) * 32;
    F->bpe = (threads + F->block - 1) / F->block;
    MppiParams mp_geo = h->mp;
    const size_t draws = (cfg->noise_source == CPS_FLEET_NOISE_PHILOX) ? (size_t)h->n_ind * F->block * (F->pair ? 2 : 1) : 0;
    F->smem = sizeof(float) * mppi_smem_floats(mp_geo, h->cfg.cost_id, F->block, F->pair ? 2 : 1, draws);
    F->rs_off = mp_geo.rs_off;
    if (F->smem > 200 * 1024) { delete F; return fail(h, CPS_ERR_UNSUPPORTED, "cps_fleet_create: horizon too large for shared memory"); }
    h->fleet = F;
    const size_t E = F->E;
    CUDA_TRY(h, cudaMalloc(&F->d_s, sizeof(float) * E * 8));
    CUDA_TRY(h, cudaMalloc(&F->d_unom, sizeof(float) * E * T));
    CUDA_TRY(h, cudaMalloc(&F->d_uprev, sizeof(float) * E));
    CUDA_TRY(h, cudaMalloc(&F->d_partials, sizeof(float) * E * F->bpe * (h->n_red + 2)));
    CUDA_TRY(h, cudaMalloc(&F->d_tickets, sizeof(unsigned) * E));
    CUDA_TRY(h, cudaMemset(F->d_s, 0, sizeof(float) * E * 8));
    CUDA_TRY(h, cudaMemset(F->d_unom, 0, sizeof(float) * E * T));
    CUDA_TRY(h, cudaMemset(F->d_uprev, 0, sizeof(float) * E));
    CUDA_TRY(h, cudaMemset(F->d_tickets, 0, sizeof(unsigned) * E));
    CUDA_TRY(h, cudaDeviceSynchronize());
    return CPS_OK;
}

extern "C" int cps_fleet_set_states(cps_handle *h, const float *s_host, long long period) {
    if (!h) return CPS_ERR_INVALID;
    FleetState *F = h->fleet;
    if (!F) return fail(h, CPS_ERR_NOT_CONFIGURED, "cps_fleet_set_states: no fleet (cps_fleet_create)");
    if (!s_host || period < 0) return fail(h, CPS_ERR_INVALID, "cps_fleet_set_states: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    float *tmp = new (std::nothrow) float[(size_t)F->E * 8];
    if (!tmp) return fail(h, CPS_ERR_INVALID, "cps_fleet_set_states: out of host memory");
    for (int e = 0; e < F->E; ++e) {
        for (int c = 0; c < 6; ++c) tmp[(size_t)e * 8 + c] = s_host[(size_t)e * 6 + c];
        tmp[(size_t)e * 8 + 6] = tmp[(size_t)e * 8 + 7] = 0.0f;
    }
    cudaError_t er = cudaMemcpyAsync(F->d_s, tmp, sizeof(float) * (size_t)F->E * 8, cudaMemcpyHostToDevice, h->stream);
    if (er == cudaSuccess) er = cudaMemsetAsync(F->d_unom, 0, sizeof(float) * (size_t)F->E * h->cfg.horizon, h->stream);
    if (er == cudaSuccess) er = cudaMemsetAsync(F->d_uprev, 0, sizeof(float) * (size_t)F->E, h->stream);
    if (er == cudaSuccess) er = cudaStreamSynchronize(h->stream);
    delete[] tmp;
    CUDA_TRY(h, er);
    F->period = period;
    if (F->pm.on) {   // ring = the reference's initial buffer; the first solve sees the true state
        F->pm.pushed = 0;
        fleet_observe_reset_kernel<<<(F->E + 127) / 128, 128, 0, h->stream>>>(F->pm, F->d_s, F->E);
        h->launches += 1;
        CUDA_TRY(h, cudaGetLastError());
    }
    return CPS_OK;
}

extern "C" int cps_fleet_set_plant_models(cps_handle *h, const cps_fleet_plant_models *m) {
    if (!h) return CPS_ERR_INVALID;
    FleetState *F = h->fleet;
    if (!F) return fail(h, CPS_ERR_NOT_CONFIGURED, "cps_fleet_set_plant_models: no fleet (cps_fleet_create)");
    if (!m || m->struct_size != (int)sizeof(cps_fleet_plant_models))
        return fail(h, CPS_ERR_INVALID, "cps_fleet_set_plant_models: cps_fleet_plant_models size mismatch");
    if (m->control_noise_mode < 0 || m->control_noise_mode > 2) return fail(h, CPS_ERR_INVALID, "cps_fleet_set_plant_models: control_noise_mode must be 0 (OFF), 1 (additive) or 2 (truncnorm)");
    if (m->control_noise_mode == 2 && !(m->control_noise_mult > 0.0f)) return fail(h, CPS_ERR_INVALID, "cps_fleet_set_plant_models: truncnorm needs a positive scale");
    const double lat_len = m->latency / F->cfg.dt_simulation;
    if (!(m->latency >= 0.0) || lat_len > 200.0)   // MAX_LATENCY_LEN (latency_adder.py:8,74-75)
        return fail(h, CPS_ERR_INVALID, "cps_fleet_set_plant_models: latency must lie in [0, 200 plant ticks]");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    PlantModels &M = F->pm;
    cudaFree(M.obs); cudaFree(M.ring);
    memset(&M, 0, sizeof(M));
    M.ctrl_mode = m->control_noise_mode; M.ctrl_mult = m->control_noise_mult; M.ctrl_add = m->control_noise_add;
    M.meas_on = m->measurement_noise ? 1 : 0;
    M.sig_a = m->sigma_angle; M.sig_p = m->sigma_position; M.sig_aD = m->sigma_angleD; M.sig_pD = m->sigma_positionD;
    M.lat_int = (int)lat_len; M.lat_frac = lat_len - (double)M.lat_int;
    M.ring_len = M.lat_int + 2;
    M.on = (M.ctrl_mode || M.meas_on || m->latency > 0.0) ? 1 : 0;
    if (!M.on) return CPS_OK;
    CUDA_TRY(h, cudaMalloc(&M.obs, sizeof(float) * (size_t)F->E * 8));
    CUDA_TRY(h, cudaMalloc(&M.ring, sizeof(float) * (size_t)F->E * M.ring_len * 6));
    M.pushed = 0;
    fleet_observe_reset_kernel<<<(F->E + 127) / 128, 128, 0, h->stream>>>(M, F->d_s, F->E);
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return CPS_OK;
}

extern "C" int cps_fleet_get_observed(cps_handle *h, float *obs_host) {
    if (!h) return CPS_ERR_INVALID;
    FleetState *F = h->fleet;
    if (!F || !F->pm.on) return fail(h, CPS_ERR_NOT_CONFIGURED, "cps_fleet_get_observed: no plant-side models (cps_fleet_set_plant_models)");
    if (!obs_host) return fail(h, CPS_ERR_INVALID, "cps_fleet_get_observed: null pointer");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    float *tmp = new (std::nothrow) float[(size_t)F->E * 8];
    if (!tmp) return fail(h, CPS_ERR_INVALID, "cps_fleet_get_observed: out of host memory");
    cudaError_t er = cudaMemcpyAsync(tmp, F->pm.obs, sizeof(float) * (size_t)F->E * 8, cudaMemcpyDeviceToHost, h->stream);
    if (er == cudaSuccess) er = cudaStreamSynchronize(h->stream);
    if (er == cudaSuccess)
        for (int e = 0; e < F->E; ++e)
            for (int c = 0; c < 6; ++c) obs_host[(size_t)e * 6 + c] = tmp[(size_t)e * 8 + c];
    delete[] tmp;
    CUDA_TRY(h, er);
    return CPS_OK;
}

extern "C" int cps_fleet_reset(cps_handle *h, long long period) {
    if (!h) return CPS_ERR_INVALID;
    FleetState *F = h->fleet;
    if (!F) return fail(h, CPS_ERR_NOT_CONFIGURED, "cps_fleet_reset: no fleet (cps_fleet_create)");
    if (period < 0) return fail(h, CPS_ERR_INVALID, "cps_fleet_reset: period < 0");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaMemsetAsync(F->d_unom, 0, sizeof(float) * (size_t)F->E * h->cfg.horizon, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(F->d_uprev, 0, sizeof(float) * (size_t)F->E, h->stream));
    F->period = period;
    return CPS_OK;
}

extern "C" int cps_fleet_get_states(cps_handle *h, float *s_host, float *u_nom_host, float *u_prev_host) {
    if (!h) return CPS_ERR_INVALID;
    FleetState *F = h->fleet;
    if (!F) return fail(h, CPS_ERR_NOT_CONFIGURED, "cps_fleet_get_states: no fleet (cps_fleet_create)");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (s_host) {
        float *tmp = new (std::nothrow) float[(size_t)F->E * 8];
        if (!tmp) return fail(h, CPS_ERR_INVALID, "cps_fleet_get_states: out of host memory");
        cudaError_t er = cudaMemcpyAsync(tmp, F->d_s, sizeof(float) * (size_t)F->E * 8, cudaMemcpyDeviceToHost, h->stream);
        if (er == cudaSuccess) er = cudaStreamSynchronize(h->stream);
        if (er == cudaSuccess)
            for (int e = 0; e < F->E; ++e)
                for (int c = 0; c < 6; ++c) s_host[(size_t)e * 6 + c] = tmp[(size_t)e * 8 + c];
        delete[] tmp;
        CUDA_TRY(h, er);
    }
    if (u_nom_host)
        CUDA_TRY(h, cudaMemcpyAsync(u_nom_host, F->d_unom, sizeof(float) * (size_t)F->E * h->cfg.horizon, cudaMemcpyDeviceToHost, h->stream));
    if (u_prev_host)
        CUDA_TRY(h, cudaMemcpyAsync(u_prev_host, F->d_uprev, sizeof(float) * (size_t)F->E, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return CPS_OK;
}

extern "C" long long cps_fleet_period(const cps_handle *h) { return (h && h->fleet) ? h->fleet->period : -1; }

// CostParams for a given target equilibrium: fold_cost() depends on it for quadratic_boundary_grad (weight set, w[9]).
int cps_fold_cost_for(cps_handle *h, float target_equilibrium, CostParams *out);

// Shared by cps_fleet_step (closed loop: replay_dev == nullptr) and cps_fleet_relabel (states from a recording).
static int fleet_launch(cps_handle *h, const char *who, int n_periods, const float *tp_dev, const float *te_dev,
                        const float *noise_dev, float *record_dev, float *J_out_dev, const float *replay_dev,
                        const float *L_dev, const float *mp_dev, float *Q_out_dev, const int *active_dev = nullptr,
                        const float *ctrl_draws_dev = nullptr, const float *meas_draws_dev = nullptr) {
    FleetState *F = h->fleet;
    if (!F) return fail(h, CPS_ERR_NOT_CONFIGURED, "%s: no fleet (cps_fleet_create)", who);
    if (n_periods < 0) return fail(h, CPS_ERR_INVALID, "%s: negative number of periods / rows", who);
    const bool philox = F->cfg.noise_source == CPS_FLEET_NOISE_PHILOX;
    if (!philox && !noise_dev) return fail(h, CPS_ERR_INVALID, "%s: this fleet expects supplied noise", who);
    if (philox && noise_dev) return fail(h, CPS_ERR_INVALID, "%s: this fleet generates its noise (Philox); pass NULL", who);
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    FleetArgs a;
    a.ode = h->ode; a.mp = h->mp;
    a.mp.rs_off = F->rs_off;
    a.pm = F->pm;
    if (replay_dev) a.pm.on = 0;   // relabelling feeds recorded states: no plant, no observation chain
    if (a.pm.on && !philox && ((a.pm.ctrl_mode && !ctrl_draws_dev) || (a.pm.meas_on && !meas_draws_dev)))
        return fail(h, CPS_ERR_INVALID, "%s: a fleet with supplied noise needs the plant-side draws as well (cps_fleet_step_noisy)", who);
    int rc;
    if ((rc = cps_fold_cost_for(h, 1.0f, &a.cost_up)) != CPS_OK) return rc;
    if ((rc = cps_fold_cost_for(h, -1.0f, &a.cost_dn)) != CPS_OK) return rc;
    PlantParams &P = a.plant;
    P.k = h->phys[CPS_PH_K]; P.m_cart = h->phys[CPS_PH_M_CART]; P.g = h->phys[CPS_PH_G]; P.J_fric = h->phys[CPS_PH_J_FRIC];
    P.M_fric = h->phys[CPS_PH_M_FRIC]; P.u_max = h->phys[CPS_PH_U_MAX]; P.thl = h->phys[CPS_PH_TRACK_HALF_LENGTH];
    P.L = (double)h->phys[CPS_PH_L]; P.m_pole = (double)h->phys[CPS_PH_M_POLE];
    P.dt = F->cfg.dt_simulation; P.n_sim = F->cfg.sim_substeps;
    a.bpe = F->bpe;
    a.s = F->d_s; a.u_nom = F->d_unom; a.u_prev = F->d_uprev;
    a.seed = F->cfg.seed; a.e_offset = (unsigned)F->cfg.experiment_offset;
    a.partials = F->d_partials; a.tickets = F->d_tickets; a.nonfinite = h->d_nonfinite;
    a.k = P.k; a.m_cart = P.m_cart; a.g = P.g; a.J_fric = P.J_fric; a.M_fric = P.M_fric; a.u_max = P.u_max;
    a.m_pole_fixed = h->phys[CPS_PH_M_POLE];
    a.L_default = h->L_var; a.mp_default = h->m_pole_var;
    a.h_step = (double)h->cfg.dt / (double)h->cfg.substeps;
    if (F->pair && noise_dev && ((uintptr_t)noise_dev % 8) != 0)
        return fail(h, CPS_ERR_INVALID, "%s: supplied noise must be 8-byte aligned", who);
    fleet_fn fn = pick_fleet(h->cfg.integrator, h->cfg.cost_id, philox, F->pair != 0, h->ode.n == 10);
    if (!fn) return fail(h, CPS_ERR_NOT_CONFIGURED, "%s: no kernel for this configuration", who);
    if (F->smem > 48 * 1024) CUDA_TRY(h, cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F->smem));
    const size_t E = F->E, K = h->cfg.num_rollouts;
    const double dt_control = F->cfg.dt_simulation * F->cfg.sim_substeps;
    const dim3 grid(F->bpe, F->E);
    for (int j = 0; j < n_periods; ++j) {
        a.tp = tp_dev ? tp_dev + (size_t)j * E : nullptr;
        a.te = te_dev ? te_dev + (size_t)j * E : nullptr;
        a.noise = noise_dev ? noise_dev + (size_t)j * E * h->n_ind * K : nullptr;
        a.record = record_dev ? record_dev + (size_t)j * E * CPS_FLEET_RECORD : nullptr;
        a.J_out = J_out_dev ? J_out_dev + (size_t)j * E * K : nullptr;
        a.replay_s = replay_dev ? replay_dev + (size_t)j * E * 6 : nullptr;
        a.L_row = L_dev ? L_dev + (size_t)j * E : nullptr;
        a.mp_row = mp_dev ? mp_dev + (size_t)j * E : nullptr;
        a.Q_out = Q_out_dev ? Q_out_dev + (size_t)j * E : nullptr;
        a.active = active_dev ? active_dev + (size_t)j * E : nullptr;
        if (a.pm.on) {
            a.pm.pushed = F->pm.pushed;
            a.pm.ctrl_draws = ctrl_draws_dev ? ctrl_draws_dev + (size_t)j * E : nullptr;
            a.pm.meas_draws = meas_draws_dev ? meas_draws_dev + (size_t)j * F->cfg.sim_substeps * E * 4 : nullptr;
            F->pm.pushed += F->cfg.sim_substeps;
        }
        a.period = (unsigned)F->period;
        a.time = (double)F->period * dt_control;
        fn<<<grid, F->block, F->smem, h->stream>>>(a);
        F->period += 1;
        h->launches += 1;
    }
    CUDA_TRY(h, cudaGetLastError());
    return CPS_OK;
}

extern "C" int cps_fleet_step(cps_handle *h, int n_periods, const float *tp_dev, const float *te_dev, const float *noise_dev,
                              float *record_dev, float *J_out_dev) {
    if (!h) return CPS_ERR_INVALID;
    return fleet_launch(h, "cps_fleet_step", n_periods, tp_dev, te_dev, noise_dev, record_dev, J_out_dev, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int cps_fleet_step_noisy(cps_handle *h, int n_periods, const float *tp_dev, const float *te_dev, const float *noise_dev,
                                    float *record_dev, float *J_out_dev, const float *ctrl_draws_dev, const float *meas_draws_dev) {
    if (!h) return CPS_ERR_INVALID;
    if (!h->fleet || !h->fleet->pm.on)
        return fail(h, CPS_ERR_NOT_CONFIGURED, "cps_fleet_step_noisy: no plant-side models (cps_fleet_set_plant_models)");
    return fleet_launch(h, "cps_fleet_step_noisy", n_periods, tp_dev, te_dev, noise_dev, record_dev, J_out_dev, nullptr, nullptr, nullptr,
                        nullptr, nullptr, ctrl_draws_dev, meas_draws_dev);
}

extern "C" int cps_fleet_relabel(cps_handle *h, int n_rows, const float *states_dev, const float *tp_dev, const float *te_dev,
                                 const float *L_dev, const float *m_pole_dev, const float *noise_dev, float *Q_out_dev,
                                 float *J_out_dev) {
    if (!h) return CPS_ERR_INVALID;
    if (!states_dev || !Q_out_dev) return fail(h, CPS_ERR_INVALID, "cps_fleet_relabel: null pointer");
    return fleet_launch(h, "cps_fleet_relabel", n_rows, tp_dev, te_dev, noise_dev, nullptr, J_out_dev, states_dev, L_dev, m_pole_dev, Q_out_dev);
}

extern "C" int cps_fleet_relabel_masked(cps_handle *h, int n_rows, const float *states_dev, const float *tp_dev, const float *te_dev,
                                        const float *L_dev, const float *m_pole_dev, const float *noise_dev, float *Q_out_dev,
                                        float *J_out_dev, const int *active_dev) {
    if (!h) return CPS_ERR_INVALID;
    if (!states_dev || !Q_out_dev) return fail(h, CPS_ERR_INVALID, "cps_fleet_relabel_masked: null pointer");
    return fleet_launch(h, "cps_fleet_relabel_masked", n_rows, tp_dev, te_dev, noise_dev, nullptr, J_out_dev, states_dev, L_dev,
                        m_pole_dev, Q_out_dev, active_dev);
}

extern "C" int cps_fleet_noise(cps_handle *h, long long period, float *out_dev) {
    if (!h) return CPS_ERR_INVALID;
    FleetState *F = h->fleet;
    if (!F) return fail(h, CPS_ERR_NOT_CONFIGURED, "cps_fleet_noise: no fleet (cps_fleet_create)");
    if (!out_dev || period < 0) return fail(h, CPS_ERR_INVALID, "cps_fleet_noise: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const int K = h->cfg.num_rollouts;
    const dim3 grid((K + 255) / 256, F->E);
    fleet_noise_kernel<<<grid, 256, 0, h->stream>>>(F->cfg.seed, (unsigned)period, (unsigned)F->cfg.experiment_offset, F->E, K,
                                                    h->n_ind, out_dev);
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return CPS_OK;
}
