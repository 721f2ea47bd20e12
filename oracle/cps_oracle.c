/*
 * cps_oracle.c -- CPU restatement of the CartPoleSimulation MPPI rollout hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product
 * (cartpolesimulation_b200/) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path
 * (SURVEY.md section 4), so the pin is made against outputs of the unmodified reference
 * code run in the build container (oracle/gen_golden.py -> tests/golden/(all).npz,
 * checked by tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line (relative to the reference root) whose
 * arithmetic it restates.  Compile with -ffp-contract=off and without -ffast-math: the
 * fp32 paths restate eager torch ops, which round after every operation.
 *
 * State layout (CartPole/state_utilities.py:5-23, sorted names):
 *   0 angle, 1 angleD, 2 angle_cos, 3 angle_sin, 4 position, 5 positionD
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define CPS_PI_D 3.14159265358979323846

enum { IDX_ANGLE = 0, IDX_ANGLED = 1, IDX_COS = 2, IDX_SIN = 3, IDX_POS = 4, IDX_POSD = 5 };

/* physical parameter vector, all already rounded to fp32 by the caller, as
 * CartPole/cartpole_parameters.py:27-29 does (lib.to_tensor(value, float32)). */
enum { PH_K = 0, PH_MCART, PH_MPOLE, PH_G, PH_JFRIC, PH_MFRIC, PH_L, PH_UMAX, PH_TRACK_HALF, PH_N };

/* cost plugin ids (Control_Toolkit_ASF/Cost_Functions/CartPole/<name>.py) */
enum { COST_DEFAULT = 0, COST_QUADRATIC_BOUNDARY = 1, COST_QB_GRAD_MINIMAL = 2, COST_QB_GRAD = 3 };

int cps_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void cps_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------------------------------------------------------------------------------
 * ODE_v0: explicit Euler + edge bounce + fmod wrap, fp64 inside a control step.
 * ------------------------------------------------------------------------------------- */

/* _cartpole_ode (CartPole/cartpole_equations.py:71-99) as numba types it when called from
 * cartpole_fine_integration_numba (CartPole/cartpole_numba.py:62-63):
 *   - constants are 0-d float32 arrays; `k + 1`, `L/2.0`, `x ** 2` promote to float64;
 *   - products of two float32 operands stay float32 (numba has no value-based casting).
 * first != 0: (ca, sa, angleD, positionD) are still the float32 state columns
 * (first substep of a control step, cartpole_numba.py:30-31), so a few sub-products are
 * rounded to float32; afterwards those arrays are float64 (state + stateD * t_step with a
 * Python-float t_step promotes, cartpole_equations.py:130-131). */
static inline void ode_v0(double ca, double sa, double angleD, double positionD, float u, int first,
                          const float *ph, double *angleDD, double *positionDD) {
    const float k = ph[PH_K], m_cart = ph[PH_MCART], m_pole = ph[PH_MPOLE], g = ph[PH_G];
    const float J_fric = ph[PH_JFRIC], M_fric = ph[PH_MFRIC], L = ph[PH_L];
    const double kp1 = (double)k + 1.0;                 /* float32 0-d + int64 -> float64 */
    const float msum = m_cart + m_pole;                 /* float32 + float32 */
    const double A = kp1 * (double)msum - (double)m_pole * (ca * ca);
    const double L_half = (double)L / 2.0;
    double F_fric, T_fric, grav, gsa;
    if (first) {
        const float caf = (float)ca, saf = (float)sa, aDf = (float)angleD, pDf = (float)positionD;
        F_fric = (double)((-M_fric) * pDf);             /* float32 product */
        T_fric = (double)((-J_fric) * aDf);
        grav = (double)(((m_pole * g) * saf) * caf);
        gsa = (double)(g * saf);
    } else {
        F_fric = (double)(-M_fric) * positionD;
        T_fric = (double)(-J_fric) * angleD;
        grav = ((double)(m_pole * g) * sa) * ca;         /* m_pole*g is float32*float32 */
        gsa = (double)g * sa;
    }
    const double pDD = (grav + (T_fric * ca) / L_half
                        + kp1 * (-((((double)m_pole * L_half) * (angleD * angleD)) * sa) + F_fric + (double)u)) / A;
    *positionDD = pDD;
    *angleDD = (gsa + pDD * ca + T_fric / ((double)m_pole * L_half)) / (kp1 * L_half);
}

/* wrap_angle_rad_inplace (CartPole/_CartPole_mathematical_helpers.py:24-29) */
static inline double wrap_fmod(double angle) {
    const double m = fmod(angle, 2.0 * CPS_PI_D);
    if (m < -CPS_PI_D) return m + 2.0 * CPS_PI_D;
    if (m > CPS_PI_D) return m - 2.0 * CPS_PI_D;
    return m;
}

/* One control step = `n` substeps (cartpole_fine_integration_numba, CartPole/cartpole_numba.py:56-78)
 * then rounding to float32 on assignment into s_next (cartpole_numba.py:21-24). */
static inline void control_step_v0(float *s, float Q, int n, double t_step, const float *ph, float L_var) {
    float php[PH_N];
    memcpy(php, ph, sizeof(php));
    php[PH_L] = L_var; /* only L comes from variable_parameters (predictors_customization_v0.py:47-54) */
    const float u = ph[PH_UMAX] * Q; /* Q2u in float32 (cartpole_equations.py:119-127,160-164) */
    double angle = s[IDX_ANGLE], angleD = s[IDX_ANGLED], ca = s[IDX_COS], sa = s[IDX_SIN];
    double position = s[IDX_POS], positionD = s[IDX_POSD];
    const double thl = (double)ph[PH_TRACK_HALF];
    for (int i = 0; i < n; ++i) {
        double aDD, pDD;
        ode_v0(ca, sa, angleD, positionD, u, i == 0, php, &aDD, &pDD);
        /* cartpole_integration_numba: explicit Euler, all with the old derivatives (cartpole_equations.py:356-364) */
        const double angle_n = angle + angleD * t_step;
        const double angleD_n = angleD + aDD * t_step;
        const double position_n = position + positionD * t_step;
        const double positionD_n = positionD + pDD * t_step;
        angle = angle_n; angleD = angleD_n; position = position_n; positionD = positionD_n;
        ca = cos(angle);
        /* edge_bounce (cartpole_equations.py:341-347) */
        if (position >= thl || -position >= thl) {
            angleD -= 2.0 * (positionD * ca) / (0.5 * (double)L_var);
            angle += angleD * t_step;
            positionD = -positionD;
            position += positionD * t_step;
        }
        angle = wrap_fmod(angle);
        ca = cos(angle);
        sa = sin(angle);
    }
    s[IDX_ANGLE] = (float)angle; s[IDX_ANGLED] = (float)angleD; s[IDX_COS] = (float)ca;
    s[IDX_SIN] = (float)sa; s[IDX_POS] = (float)position; s[IDX_POSD] = (float)positionD;
}

/* predictor_ODE_v0.predict (SI_Toolkit/src/SI_Toolkit/Predictors/predictor_ODE_v0.py:42-74).
 * s0: [B][6] if s0_batched else [6] (tiled, :59-60).  Q: [B][T].  traj: [B][T+1][6] or NULL;
 * s_final: [B][6] or NULL. */
void cps_oracle_rollout_v0(const float *s0, int s0_batched, const float *Q, int B, int T, int n, double dt,
                           const float *ph, float L_var, float *traj, float *s_final) {
    const double t_step = dt / (double)n; /* predictors_customization_v0.py:38 */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        float s[6];
        memcpy(s, s0 + (s0_batched ? (size_t)b * 6 : 0), sizeof(s));
        if (traj) memcpy(traj + (size_t)b * (T + 1) * 6, s, sizeof(s));
        for (int t = 0; t < T; ++t) {
            control_step_v0(s, Q[(size_t)b * T + t], n, t_step, ph, L_var);
            if (traj) memcpy(traj + ((size_t)b * (T + 1) + t + 1) * 6, s, sizeof(s));
        }
        if (s_final) memcpy(s_final + (size_t)b * 6, s, sizeof(s));
    }
}

/* ---------------------------------------------------------------------------------------
 * ODE: Euler-Cromer + atan2 wrap, float32 throughout (eager torch / TF ops).
 * ------------------------------------------------------------------------------------- */

/* _cartpole_ode in float32, operation order as written (cartpole_equations.py:71-99). */
static inline void ode_f32(float ca, float sa, float angleD, float positionD, float u, const float *ph,
                           float *angleDD, float *positionDD) {
    const float k = ph[PH_K], m_cart = ph[PH_MCART], m_pole = ph[PH_MPOLE], g = ph[PH_G];
    const float J_fric = ph[PH_JFRIC], M_fric = ph[PH_MFRIC], L = ph[PH_L];
    const float kp1 = k + 1.0f;
    const float A = kp1 * (m_cart + m_pole) - m_pole * (ca * ca);
    const float F_fric = (-M_fric) * positionD;
    const float T_fric = (-J_fric) * angleD;
    const float L_half = L / 2.0f;
    const float t1 = ((m_pole * g) * sa) * ca;
    const float t2 = (T_fric * ca) / L_half;
    const float inner = (-(((m_pole * L_half) * (angleD * angleD)) * sa) + F_fric) + u;
    const float pDD = ((t1 + t2) + kp1 * inner) / A;
    *positionDD = pDD;
    *angleDD = ((g * sa + pDD * ca) + T_fric / (m_pole * L_half)) / (kp1 * L_half);
}

/* CartPoleEquations._cartpole_fine_integration (cartpole_equations.py:214-261) with
 * _cartpole_integration_euler_cromer (:292-303) and wrap_angle_rad = atan2 (:306-308). */
static inline void control_step_cromer(float *s, float Q, int n, float t_step, const float *ph) {
    const float u = ph[PH_UMAX] * Q;
    float angle = s[IDX_ANGLE], angleD = s[IDX_ANGLED], ca = s[IDX_COS], sa = s[IDX_SIN];
    float position = s[IDX_POS], positionD = s[IDX_POSD];
    for (int i = 0; i < n; ++i) {
        float aDD, pDD;
        ode_f32(ca, sa, angleD, positionD, u, ph, &aDD, &pDD);
        angleD = angleD + aDD * t_step;
        positionD = positionD + pDD * t_step;
        angle = angle + angleD * t_step;
        position = position + positionD * t_step;
        ca = cosf(angle);
        sa = sinf(angle);
        angle = atan2f(sa, ca);
    }
    s[IDX_ANGLE] = angle; s[IDX_ANGLED] = angleD; s[IDX_COS] = ca;
    s[IDX_SIN] = sa; s[IDX_POS] = position; s[IDX_POSD] = positionD;
}

/* predictor_ODE._predict_core (SI_Toolkit/src/SI_Toolkit/Predictors/predictor_ODE.py:86-97) via
 * autoregression_loop.run (autoregression.py:33-108) and next_state_predictor_ODE._step
 * (SI_Toolkit_ASF/ToolkitCustomization/predictors_customization.py:49-66): L and m_pole are taken
 * from variable_parameters, i.e. the caller passes them inside `ph`. */
void cps_oracle_rollout_cromer(const float *s0, int s0_batched, const float *Q, int B, int T, int n, double dt,
                               const float *ph, float *traj, float *s_final) {
    const float t_step = (float)(dt / (double)n); /* python float, cast to the tensor dtype by torch */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        float s[6];
        memcpy(s, s0 + (s0_batched ? (size_t)b * 6 : 0), sizeof(s));
        if (traj) memcpy(traj + (size_t)b * (T + 1) * 6, s, sizeof(s));
        for (int t = 0; t < T; ++t) {
            control_step_cromer(s, Q[(size_t)b * T + t], n, t_step, ph);
            if (traj) memcpy(traj + ((size_t)b * (T + 1) + t + 1) * 6, s, sizeof(s));
        }
        if (s_final) memcpy(s_final + (size_t)b * 6, s, sizeof(s));
    }
}

/* ---------------------------------------------------------------------------------------
 * float64 "truth" integration of either scheme: same formulas, no float32 rounding anywhere (inputs are the
 * float32 values promoted).  Not a restatement of any reference code path -- it measures the noise floor:
 * how far the reference's own float32 outputs are from the exact iteration, so that the CUDA kernels can be
 * required to be no further away than that (tests/test_gpu_parity.py::test_fp32_noise_floor).
 * ------------------------------------------------------------------------------------- */
void cps_oracle_rollout_f64(int integrator, const float *s0, int s0_batched, const float *Q, int B, int T, int n,
                            double dt, const float *ph, double *traj /* [B][T+1][6] */) {
    const double h = dt / (double)n;
    const double k = ph[PH_K], m_cart = ph[PH_MCART], m_pole = ph[PH_MPOLE], g = ph[PH_G];
    const double J_fric = ph[PH_JFRIC], M_fric = ph[PH_MFRIC], L = ph[PH_L], u_max = ph[PH_UMAX];
    const double thl = ph[PH_TRACK_HALF], kp1 = k + 1.0, L_half = L / 2.0;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const float *s = s0 + (s0_batched ? (size_t)b * 6 : 0);
        double angle = s[IDX_ANGLE], angleD = s[IDX_ANGLED], ca = s[IDX_COS], sa = s[IDX_SIN];
        double position = s[IDX_POS], positionD = s[IDX_POSD];
        double *row = traj + (size_t)b * (T + 1) * 6;
        row[0] = angle; row[1] = angleD; row[2] = ca; row[3] = sa; row[4] = position; row[5] = positionD;
        for (int t = 0; t < T; ++t) {
            const double u = u_max * (double)Q[(size_t)b * T + t];
            for (int i = 0; i < n; ++i) {
                const double A = kp1 * (m_cart + m_pole) - m_pole * (ca * ca);
                const double F_fric = -M_fric * positionD, T_fric = -J_fric * angleD;
                const double pDD = (m_pole * g * sa * ca + (T_fric * ca) / L_half
                                    + kp1 * (-(m_pole * L_half * (angleD * angleD) * sa) + F_fric + u)) / A;
                const double aDD = (g * sa + pDD * ca + T_fric / (m_pole * L_half)) / (kp1 * L_half);
                if (integrator == 0) {
                    const double an = angle + angleD * h, pn = position + positionD * h;
                    angleD += aDD * h; positionD += pDD * h; angle = an; position = pn;
                    ca = cos(angle);
                    if (position >= thl || -position >= thl) {
                        angleD -= 2.0 * (positionD * ca) / (0.5 * L);
                        angle += angleD * h;
                        positionD = -positionD;
                        position += positionD * h;
                    }
                    angle = wrap_fmod(angle);
                    ca = cos(angle); sa = sin(angle);
                } else {
                    angleD += aDD * h; positionD += pDD * h;
                    angle += angleD * h; position += positionD * h;
                    ca = cos(angle); sa = sin(angle);
                    angle = atan2(sa, ca);
                }
            }
            double *o = row + (size_t)(t + 1) * 6;
            o[0] = angle; o[1] = angleD; o[2] = ca; o[3] = sa; o[4] = position; o[5] = positionD;
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * Cost plugins (float32, torch op order).
 * cp[] layout per plugin:
 *  DEFAULT / QUADRATIC_BOUNDARY: [dd_weight, ep_weight, cc_weight, ccrc_weight, R, MAX_COST]
 *  QB_GRAD_MINIMAL:              [dd_quadratic_w, db_w, ep_w, ekp_w, cc_w, R, permissible_track_fraction]
 *  QB_GRAD: [dd_q, dd_lin, db, ep, ekp, cc, ccrc (the up or down set, chosen by the caller as
 *            quadratic_boundary_grad.py:182-200 does), R, permissible_track_fraction,
 *            target_angular_speed_sqr_max_correction, admissible_angle(rad)]
 * ------------------------------------------------------------------------------------- */

static inline float sqf(float x) { return x * x; }

/* un-shifted stage cost of one (state, input) pair */
static float stage_cost_one(int cost_id, const float *cp, const float *s, float u, float u_prev, float thl,
                            float target_position, float target_equilibrium) {
    const float position = s[IDX_POS], angle = s[IDX_ANGLE], angleD = s[IDX_ANGLED];
    switch (cost_id) {
    case COST_DEFAULT: {
        /* default.py:23-88: dd + ep + cc (+ ccrc = 0, :84-86) */
        const float ddc = sqf((position - target_position) / (2.0f * thl))
                          + ((fabsf(position) > 0.90f * thl) ? 1.0f : 0.0f) * 1.0e7f;
        const float dd = cp[0] * ddc;
        const float ep = cp[1] * ((target_equilibrium * 0.25f) * sqf(1.0f - cosf(angle)));
        const float cc = cp[2] * (cp[4] * (u * u));
        return (dd + ep) + cc;
    }
    case COST_QUADRATIC_BOUNDARY: {
        /* quadratic_boundary.py:26-87 */
        const float mask = (fabsf(position) > 0.95f * thl) ? 1.0f : 0.0f;
        const float ddc = sqf((position - target_position) / (2.0f * thl))
                          + (mask * 1e9f) * sqf((fabsf(position) - 0.95f * thl) / (0.05f * thl));
        const float dd = cp[0] * ddc;
        const float ep = cp[1] * ((target_equilibrium * 0.25f) * sqf(1.0f - cosf(angle)));
        const float cc = cp[2] * (cp[4] * (u * u));
        const float ccrc = cp[3] * sqf(u - u_prev);
        return ((dd + ep) + cc) + ccrc;
    }
    case COST_QB_GRAD_MINIMAL: {
        /* quadratic_boundary_grad_minimal.py:64-126 */
        const float f = cp[6];
        const float ddq = cp[0] * sqf((position - target_position) / (2.0f * thl));
        const float mask = (fabsf(position) > f * thl) ? 1.0f : 0.0f;
        const float db = cp[1] * (mask * sqf((fabsf(position) - f * thl) / ((1.0f - f) * thl)));
        const float ep = cp[2] * sqf(1.0f - target_equilibrium * cosf(angle));
        const float ekp = cp[3] * (angleD * angleD);
        const float cc = cp[4] * (cp[5] * (u * u));
        return (((ddq + db) + ep) + ekp) + cc;
    }
    case COST_QB_GRAD: {
        /* quadratic_boundary_grad.py:64-241 (stop_gradient is the identity in the forward pass) */
        const float f = cp[8];
        const float e = target_equilibrium;
        const float dist = (position - target_position) / (2.0f * thl);
        const float ddq = cp[0] * sqf(dist);
        const float ddl = cp[1] * fabsf(dist);
        const float mask = (fabsf(position) > f * thl) ? 1.0f : 0.0f;
        const float db = cp[2] * (mask * sqf((fabsf(position) - f * thl) / ((1.0f - f) * thl)));
        const float ca = cosf(angle);
        const float ep = cp[3] * (sqf(2.0f - e * ca) - 1.0f);
        /* _E_kin_cost (:115-142) */
        const float dflt = (120.0f * (1.0f + e)) / 2.0f;
        const float tmax = fabsf(dflt + cp[9]);
        const float basic = (1.0f - e * ca) / 2.0f;
        const float cond = e * (ca - cosf(cp[10]));
        const float scaling = (cond > 0.0f) ? 0.0f : basic;
        const float ekp = cp[4] * fabsf(angleD * angleD - tmax * scaling);
        const float cc = cp[5] * (cp[7] * (u * u));
        const float ccrc = cp[6] * sqf(u - u_prev);
        /* stage_cost = dd_linear + dd_quadratic + db + ep + ekp + cc + ccrc (:219) */
        return (((((ddl + ddq) + db) + ep) + ekp) + cc) + ccrc;
    }
    default:
        return NAN;
    }
}

static float terminal_cost_one(int cost_id, const float *s, float thl, float target_position) {
    if (cost_id == COST_DEFAULT || cost_id == COST_QUADRATIC_BOUNDARY) {
        /* default.py:59-68 / quadratic_boundary.py:57-66 */
        const int bad = (fabsf(s[IDX_ANGLE]) > 0.2f) || (fabsf(s[IDX_POS] - target_position) > 0.1f * thl);
        return 10000.0f * (bad ? 1.0f : 0.0f);
    }
    return 0.0f; /* *_grad*.py get_terminal_cost: zeros */
}

static inline int cost_is_shifted(int cost_id) {
    /* plugins that implement _get_stage_cost get `- MAX_COST` from the base class
     * (Control_Toolkit/Cost_Functions/__init__.py:63-64); the *_grad* ones override get_stage_cost. */
    return cost_id == COST_DEFAULT || cost_id == COST_QUADRATIC_BOUNDARY;
}

/* get_stage_cost semantics (shifted where the reference shifts) -> stage[K][T];
 * with unshifted != 0 returns _get_stage_cost instead. */
void cps_oracle_stage_cost(int cost_id, const float *cp, const float *traj, const float *Q, float u_prev, int K,
                           int T, float thl, float target_position, float target_equilibrium, int unshifted,
                           float *stage) {
    const float max_cost = (cost_is_shifted(cost_id) && !unshifted) ? cp[5] : 0.0f;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k) {
        for (int t = 0; t < T; ++t) {
            const float up = (t == 0) ? u_prev : Q[(size_t)k * T + t - 1];
            const float c = stage_cost_one(cost_id, cp, traj + ((size_t)k * (T + 1) + t) * 6, Q[(size_t)k * T + t], up,
                                           thl, target_position, target_equilibrium);
            stage[(size_t)k * T + t] = c - max_cost;
        }
    }
}

/* Summation order of torch's CPU `sum` over a contiguous row of n float32 values -- the order `lib.mean(..., 1)` of
 * get_trajectory_cost (Control_Toolkit/Cost_Functions/__init__.py:90-93) runs in with the torch library, since
 * torch.mean on CPU is sum followed by a TRUE division by n.  It is a library detail (ATen/native/cpu/SumKernel.cpp,
 * cascade_sum; torch 2.11), established empirically with exact-arithmetic probes (which pairs of addends meet before a
 * third one) and checked bit for bit against torch.sum for n = 1..5000 (tests/test_oracle_golden.py::test_torch_row_sum):
 *   - vectors of 8 lanes (the kernel is built for AVX2 even on AVX-512 hosts), 4 vector accumulators ("ilp"):
 *     element e < 8*(n/8) sits in vector v = e/8, lane e%8; vectors of the complete groups of 4 accumulate into
 *     acc[v%4], every 16 groups the accumulators cascade into a second (third, fourth) level; left-over vectors go to
 *     accumulator 0; then acc0 += acc1, acc2, acc3;
 *   - the n%8 tail elements are summed sequentially from zero, then the 8 lanes of acc0 are added in lane order;
 *   - n < 8: scalar path, 4 interleaved accumulators, left-overs to accumulator 0, then a0 + a1 + a2 + a3.
 * With the MAX_COST plugins every addend is ~ -6e9 (ulp 512) and the row sum ~ -3e11 (ulp 32768): the order decides
 * the bucket a rollout lands in, so it has to be THIS order for the controls to agree with the reference. */
#define TRS_LEVELS 4
float cps_oracle_torch_row_sum(const float *x, int n) {
    if (n < 8) {
        float a[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const int g = n / 4;
        for (int i = 0; i < g; ++i)
            for (int k = 0; k < 4; ++k) a[k] += x[i * 4 + k];
        for (int i = g * 4; i < n; ++i) a[0] += x[i];
        for (int k = 1; k < 4; ++k) a[0] += a[k];
        return a[0];
    }
    const int vec_size = n / 8, size_ilp = vec_size / 4;
    int level_power = 4;
    {   /* max(4, CeilLog2(size_ilp) / 4) */
        int cl = 0;
        while ((1 << cl) < size_ilp) ++cl;
        if (cl / TRS_LEVELS > level_power) level_power = cl / TRS_LEVELS;
    }
    const int level_step = 1 << level_power, level_mask = level_step - 1;
    float acc[TRS_LEVELS][4][8];
    memset(acc, 0, sizeof(acc));
    int i = 0;
    while (i + level_step <= size_ilp) {
        for (int j = 0; j < level_step; ++j, ++i)
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 8; ++l) acc[0][k][l] += x[(i * 4 + k) * 8 + l];
        for (int j = 1; j < TRS_LEVELS; ++j) {
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 8; ++l) { acc[j][k][l] += acc[j - 1][k][l]; acc[j - 1][k][l] = 0.0f; }
            if ((i & (level_mask << (j * level_power))) != 0) break;
        }
    }
    for (; i < size_ilp; ++i)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l) acc[0][k][l] += x[(i * 4 + k) * 8 + l];
    for (int j = 1; j < TRS_LEVELS; ++j)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l) acc[0][k][l] += acc[j][k][l];
    for (int v = size_ilp * 4; v < vec_size; ++v)
        for (int l = 0; l < 8; ++l) acc[0][0][l] += x[v * 8 + l];
    for (int k = 1; k < 4; ++k)
        for (int l = 0; l < 8; ++l) acc[0][0][l] += acc[0][k][l];
    float fin = 0.0f;
    for (int e = vec_size * 8; e < n; ++e) fin += x[e];
    for (int l = 0; l < 8; ++l) fin += acc[0][0][l];
    return fin;
}

/* cost_function_base.get_trajectory_cost (Control_Toolkit/Cost_Functions/__init__.py:74-93):
 * mean over the T+1 entries [stage_0 - MAX .. stage_{T-1} - MAX, terminal], in the summation order of the torch
 * library (cps_oracle_torch_row_sum), then divided by T+1. */
void cps_oracle_trajectory_cost(int cost_id, const float *cp, const float *traj, const float *Q, float u_prev,
                                int K, int T, float thl, float target_position, float target_equilibrium,
                                float *J) {
    const float max_cost = cost_is_shifted(cost_id) ? cp[5] : 0.0f;
#pragma omp parallel
    {
        float *row = (float *)malloc(sizeof(float) * (size_t)(T + 1));
#pragma omp for schedule(static)
        for (int k = 0; k < K; ++k) {
            for (int t = 0; t < T; ++t) {
                const float up = (t == 0) ? u_prev : Q[(size_t)k * T + t - 1];
                const float c = stage_cost_one(cost_id, cp, traj + ((size_t)k * (T + 1) + t) * 6, Q[(size_t)k * T + t], up,
                                               thl, target_position, target_equilibrium);
                row[t] = c - max_cost;
            }
            row[T] = terminal_cost_one(cost_id, traj + ((size_t)k * (T + 1) + T) * 6, thl, target_position);
            J[k] = cps_oracle_torch_row_sum(row, T + 1) / (float)(T + 1);
        }
        free(row);
    }
}

/* ---------------------------------------------------------------------------------------
 * Noise interpolation (Control_Toolkit/others/Interpolator.py:53-84,97-106)
 * ------------------------------------------------------------------------------------- */

int cps_oracle_num_inducing(int T, int p) { return (int)ceil((double)(T - 1) / (double)p) + 1; }

/* interp_mat[n_ind][T] exactly as calculate_interpolation_matrix builds it (Interpolator.py:53-77):
 * rows r < (n_ind-1)*p get the tent weights (p-j)/p and j/p (float32 division by `step`); the very last
 * row r == (n_ind-1)*p is set to 1 and THEN divided by step like everything else (:73-74), i.e. it is 1/p.
 * That row only survives the truncation to T rows when (T-1) % p == 0. */
void cps_oracle_interp_matrix(int T, int p, float *W /* [n_ind][T] */) {
    const int n_ind = cps_oracle_num_inducing(T, p);
    const int last = (n_ind - 1) * p;
    memset(W, 0, sizeof(float) * (size_t)n_ind * T);
    for (int r = 0; r < T; ++r) {
        if (r < last) {
            const int i = r / p, j = r % p;
            W[(size_t)i * T + r] = (float)(p - j) / (float)p;
            W[(size_t)(i + 1) * T + r] = (float)j / (float)p;
        } else if (r == last) {
            W[(size_t)(n_ind - 1) * T + r] = 1.0f / (float)p;
        }
    }
}

/* delta_u[K][T] = (eps[K][n_ind] * stdev) @ W  (optimizer_mppi.py:169-178). */
void cps_oracle_interpolate(const float *eps, int K, int T, int p, float stdev, float *delta_u) {
    const int n_ind = cps_oracle_num_inducing(T, p);
    float *W = (float *)malloc(sizeof(float) * (size_t)n_ind * T);
    cps_oracle_interp_matrix(T, p, W);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k) {
        for (int t = 0; t < T; ++t) {
            float acc = 0.0f;
            for (int i = 0; i < n_ind; ++i) {
                const float w = W[(size_t)i * T + t];
                if (w != 0.0f) acc += (eps[(size_t)k * n_ind + i] * stdev) * w;
            }
            delta_u[(size_t)k * T + t] = acc;
        }
    }
    free(W);
}

/* ---------------------------------------------------------------------------------------
 * One MPPI solve (Control_Toolkit/Optimizers/optimizer_mppi.py:153-192)
 * ------------------------------------------------------------------------------------- */

static inline float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* mp[] = [cc_weight, R, LBD, NU, SQRTRHODTINV, action_low, action_high]
 * integrator: 0 = ODE_v0 (explicit Euler), 1 = ODE (Euler-Cromer).
 * u_nom[T] is updated in place exactly like self.u_nom (shift at the START of the solve, :183).
 * delta_u_in: if non-NULL, [K][T] perturbations are taken as given (post-interpolation injection),
 * else they are interpolated from eps[K][n_ind].
 * Optional outputs (may be NULL): J[K], traj[K][T+1][6], u_run[K][T], delta_u_out[K][T]. Returns u. */
/* Everything of _predict_and_cost that follows the rollout (:188-190): trajectory cost, MPPI correction cost,
 * exp-weighted average, clip.  u_nom is the SHIFTED nominal sequence on entry and the updated one on exit. */
static float mppi_tail(int cost_id, const float *cp, const float *mp, const float *ph, float *u_nom,
                       const float *delta_u, const float *u_run, const float *traj, float u_prev,
                       float target_position, float target_equilibrium, int K, int T, float *J) {
    const float cc_weight = mp[0], R = mp[1], LBD = mp[2], NU = mp[3], lo = mp[5], hi = mp[6];
    cps_oracle_trajectory_cost(cost_id, cp, traj, u_run, u_prev, K, T, ph[PH_TRACK_HALF], target_position,
                               target_equilibrium, J);
    /* mppi_correction_cost (:153-154), summed over the horizon */
    const float c1 = (0.5f * (1.0f - 1.0f / NU)) * R;
    float *row = (float *)malloc(sizeof(float) * (size_t)T);
    for (int k = 0; k < K; ++k) {
        for (int t = 0; t < T; ++t) {
            const float du = delta_u[(size_t)k * T + t], u = u_run[(size_t)k * T + t];
            row[t] = cc_weight * ((c1 * (du * du) + (R * u) * du) + (0.5f * R) * (u * u));
        }
        J[k] = J[k] + cps_oracle_torch_row_sum(row, T); /* lib.sum(..., (1, 2)) of a contiguous [K, T, 1] (:159) */
    }
    free(row);
    /* reward_weighted_average (:162-167) */
    float rho = J[0];
    for (int k = 1; k < K; ++k) rho = (J[k] < rho) ? J[k] : rho;
    const float neg_inv_lbd = (float)(-1.0 / (double)LBD);
    double a = 0.0; /* accumulate in double: the torch sum order is a library detail, double removes order sensitivity */
    double *b = (double *)calloc((size_t)T, sizeof(double));
    for (int k = 0; k < K; ++k) {
        const float w = expf(neg_inv_lbd * (J[k] - rho));
        a += (double)w;
        for (int t = 0; t < T; ++t) b[t] += (double)(w * delta_u[(size_t)k * T + t]);
    }
    for (int t = 0; t < T; ++t) u_nom[t] = clipf(u_nom[t] + (float)b[t] / (float)a, lo, hi); /* :189 */
    free(b);
    return u_nom[0];
}

void cps_oracle_net_rollout(int net_type, int n_state_in, int n_layers, const int *hsz, int n_out,
                            const float *weights, const int *in_idx, const int *out_idx, const float *norm_a,
                            const float *norm_b, const float *denorm_A, const float *denorm_B, const float *s0,
                            int s0_batched, const float *Q, const float *h0, int h0_batched, int B, int T,
                            float *traj, float *h_final);

/* integrator: 0 = ODE_v0, 1 = ODE, 2 = neural (net_* arguments used; h0 = the predictor's stored hidden state,
 * shared by all rollouts, predictor_autoregressive_neural.py:291). */
static float mppi_step_impl(int integrator, int cost_id, const float *cp, const float *mp, const float *ph,
                            const float *s, float *u_nom, const float *eps, const float *delta_u_in, float u_prev,
                            float target_position, float target_equilibrium, int K, int T, int n, double dt, int p,
                            float *J_out, float *traj_out, float *u_run_out, float *delta_u_out,
                            int net_type, int n_state_in, int n_layers, const int *hsz, int n_out,
                            const float *weights, const int *in_idx, const int *out_idx, const float *norm_a,
                            const float *norm_b, const float *denorm_A, const float *denorm_B, const float *h0) {
    const float stdev = mp[4], lo = mp[5], hi = mp[6];
    float *delta_u = (float *)malloc(sizeof(float) * (size_t)K * T);
    float *u_run = (float *)malloc(sizeof(float) * (size_t)K * T);
    float *traj = (float *)malloc(sizeof(float) * (size_t)K * (T + 1) * 6);
    float *J = (float *)malloc(sizeof(float) * (size_t)K);
    /* u_nom = concat([u_nom[1:], u_nom[-1:]]) (:183) */
    for (int t = 0; t + 1 < T; ++t) u_nom[t] = u_nom[t + 1];
    if (delta_u_in) memcpy(delta_u, delta_u_in, sizeof(float) * (size_t)K * T);
    else cps_oracle_interpolate(eps, K, T, p, stdev, delta_u);
    for (size_t i = 0; i < (size_t)K * T; ++i) u_run[i] = clipf(u_nom[i % T] + delta_u[i], lo, hi); /* :185-186 */
    if (integrator == 0) cps_oracle_rollout_v0(s, 0, u_run, K, T, n, dt, ph, ph[PH_L], traj, NULL);
    else if (integrator == 1) cps_oracle_rollout_cromer(s, 0, u_run, K, T, n, dt, ph, traj, NULL);
    else cps_oracle_net_rollout(net_type, n_state_in, n_layers, hsz, n_out, weights, in_idx, out_idx, norm_a, norm_b,
                                denorm_A, denorm_B, s, 0, u_run, h0, 0, K, T, traj, NULL);
    const float u = mppi_tail(cost_id, cp, mp, ph, u_nom, delta_u, u_run, traj, u_prev, target_position,
                              target_equilibrium, K, T, J);
    if (J_out) memcpy(J_out, J, sizeof(float) * (size_t)K);
    if (traj_out) memcpy(traj_out, traj, sizeof(float) * (size_t)K * (T + 1) * 6);
    if (u_run_out) memcpy(u_run_out, u_run, sizeof(float) * (size_t)K * T);
    if (delta_u_out) memcpy(delta_u_out, delta_u, sizeof(float) * (size_t)K * T);
    free(J); free(traj); free(u_run); free(delta_u);
    return u;
}

float cps_oracle_mppi_step(int integrator, int cost_id, const float *cp, const float *mp, const float *ph,
                           const float *s, float *u_nom, const float *eps, const float *delta_u_in, float u_prev,
                           float target_position, float target_equilibrium, int K, int T, int n, double dt, int p,
                           float *J_out, float *traj_out, float *u_run_out, float *delta_u_out) {
    return mppi_step_impl(integrator, cost_id, cp, mp, ph, s, u_nom, eps, delta_u_in, u_prev, target_position,
                          target_equilibrium, K, T, n, dt, p, J_out, traj_out, u_run_out, delta_u_out, 0, 0, 0, NULL, 0,
                          NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
}

/* optimizer_mppi with predictor_autoregressive_neural (optimizer_mppi.py:180-192 with predict_core of
 * predictor_autoregressive_neural.py:266-313). */
float cps_oracle_mppi_step_net(int cost_id, const float *cp, const float *mp, const float *ph, const float *s,
                               float *u_nom, const float *eps, float u_prev, float target_position,
                               float target_equilibrium, int K, int T, int p, float *J_out, float *traj_out,
                               float *u_run_out, int net_type, int n_state_in, int n_layers, const int *hsz, int n_out,
                               const float *weights, const int *in_idx, const int *out_idx, const float *norm_a,
                               const float *norm_b, const float *denorm_A, const float *denorm_B, const float *h0) {
    return mppi_step_impl(2, cost_id, cp, mp, ph, s, u_nom, eps, NULL, u_prev, target_position, target_equilibrium, K, T,
                          1, 0.0, p, J_out, traj_out, u_run_out, NULL, net_type, n_state_in, n_layers, hsz, n_out,
                          weights, in_idx, out_idx, norm_a, norm_b, denorm_A, denorm_B, h0);
}

/* ---------------------------------------------------------------------------------------
 * Autoregressive neural predictor (GRU / Dense), float32.
 * SI_Toolkit/src/SI_Toolkit/Predictors/predictor_autoregressive_neural.py:266-313,332-352
 * SI_Toolkit/src/SI_Toolkit/Functions/Pytorch/Network.py:239-287 (Sequence.forward, 'with cells')
 * ------------------------------------------------------------------------------------- */

typedef struct {
    int net_type;   /* 0 = GRU, 1 = Dense */
    int n_in;       /* net inputs = 1 control + n_state_in */
    int n_layers;   /* hidden layers */
    int n_out;
    const int *h;   /* hidden sizes [n_layers] */
    /* GRU layer l: w_ih [3H][in], w_hh [3H][H], b_ih [3H], b_hh [3H]  (torch GRUCell, gate order r,z,n)
     * Dense layer l: w [H][in], b [H].  Output layer: w_out [n_out][H_last], b_out [n_out]. */
    const float *const *w_ih;
    const float *const *w_hh;
    const float *const *b_ih;
    const float *const *b_hh;
    const float *w_out;
    const float *b_out;
    int h_max;
} net_t;

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* one network step on one sample.  x: [n_in]; hstate: concatenated hidden vectors (GRU only, updated in place);
 * y: [n_out]; scratch: >= 10*h_max floats */
static void net_step(const net_t *N, const float *x, float *hstate, float *y, float *scratch) {
    const float *in = x;
    int in_len = N->n_in;
    float *hp = hstate;
    float *gi = scratch;
    for (int l = 0; l < N->n_layers; ++l) {
        const int H = N->h[l];
        if (N->net_type == 0) {
            /* torch.nn.GRUCell: r = s(Wir x + bir + Whr h + bhr); z likewise; n = tanh(Win x + bin + r*(Whn h + bhn));
             * h' = (h - n) * z + n  (ATen/native/RNN.cpp GRUCell, non-fused CPU path) */
            float *gh = gi + 3 * H;
            for (int j = 0; j < 3 * H; ++j) {
                float a = N->b_ih[l][j];
                const float *w = N->w_ih[l] + (size_t)j * in_len;
                for (int i = 0; i < in_len; ++i) a += w[i] * in[i];
                gi[j] = a;
                float c = N->b_hh[l][j];
                const float *wh = N->w_hh[l] + (size_t)j * H;
                for (int i = 0; i < H; ++i) c += wh[i] * hp[i];
                gh[j] = c;
            }
            float *hn = gh + 3 * H;
            for (int j = 0; j < H; ++j) {
                const float r = sigmoidf_(gi[j] + gh[j]);
                const float z = sigmoidf_(gi[H + j] + gh[H + j]);
                const float nn = tanhf(gi[2 * H + j] + r * gh[2 * H + j]);
                hn[j] = (hp[j] - nn) * z + nn;
            }
            memcpy(hp, hn, sizeof(float) * H);
            in = hp;
        } else {
            /* Dense: h = tanh(W x + b) (Network.py:255-260) */
            float *hn = scratch + (size_t)(8 + (l & 1)) * N->h_max; /* ping-pong buffers */
            for (int j = 0; j < H; ++j) {
                float a = N->b_ih[l][j];
                const float *w = N->w_ih[l] + (size_t)j * in_len;
                for (int i = 0; i < in_len; ++i) a += w[i] * in[i];
                hn[j] = tanhf(a);
            }
            in = hn;
        }
        in_len = H;
        if (N->net_type == 0) hp += H;
    }
    for (int j = 0; j < N->n_out; ++j) {
        float a = N->b_out[j];
        const float *w = N->w_out + (size_t)j * in_len;
        for (int i = 0; i < in_len; ++i) a += w[i] * in[i];
        y[j] = a;
    }
}

/* Differential networks (outputs named D_*): SI_Toolkit/src/SI_Toolkit/Predictors/autoregression.py:118-158
 * (differential_model_autoregression_helper) + Functions/General/Normalising.py:111-186
 * (get_scaling_function_for_output_of_differential_network).  The network predicts normalised derivatives; the helper
 * keeps the normalised state of the output features, starting_point <- starting_point + (p1 * y + p2) with
 * p1 = a * C * dt, p2 = a * D * dt (a: normalisation of the integrated variables, C / D: de-normalisation of the
 * derivatives), hands it out as the step's output and gathers the next network input from it.
 * The configuration is a process-wide switch consulted by cps_oracle_net_rollout (and through it by
 * cps_oracle_mppi_step_net): p1 == NULL turns it off.  Test infrastructure: not thread safe by design. */
static struct {
    int on;
    float p1[8], p2[8], on_a[8], on_b[8];
    int out_to_in[8];
} g_diff = {0, {0}, {0}, {0}, {0}, {0}};

void cps_oracle_net_differential(const float *p1, const float *p2, const float *on_a, const float *on_b,
                                 const int *out_to_in, int n_out, int n_state_in) {
    g_diff.on = (p1 != NULL);
    if (!p1) return;
    for (int o = 0; o < n_out && o < 8; ++o) { g_diff.p1[o] = p1[o]; g_diff.p2[o] = p2[o]; g_diff.on_a[o] = on_a[o]; g_diff.on_b[o] = on_b[o]; }
    for (int i = 0; i < n_state_in && i < 8; ++i) g_diff.out_to_in[i] = out_to_in[i];
}

/* Flat-argument entry (ctypes friendly).
 * weights: concatenation, per hidden layer l: GRU: w_ih, w_hh, b_ih, b_hh ; Dense: w, b ; then w_out, b_out.
 * Net inputs are [Q, state features in_idx[0..n_state_in)] (state indices into the 6-vector);
 * net outputs are state features out_idx[0..n_out) (predictor_autoregressive_neural.py:279-282);
 * the net output feeds back as the next net state input (autoregression.py:94-98), which requires the
 * output features to equal the input state features (true for GRU-6IN-...-5OUT nets).
 * norm_a/b: [1 + n_state_in] for (Q, inputs); denorm_A/B: [n_out] (Functions/General/Normalising.py:15-108).
 * h0: [Htot] shared initial hidden state (memory_states_ref rows are identical across the batch) or
 *     [B][Htot] if h0_batched.  traj: [B][T+1][6]; angle = atan2(sin, cos) augmentation
 *     (predictors_customization.py:120-139); features the net does not output are zero (missing_outputs).
 * h_final: optional [B][Htot]. */
void cps_oracle_net_rollout(int net_type, int n_state_in, int n_layers, const int *hsz, int n_out,
                            const float *weights, const int *in_idx, const int *out_idx, const float *norm_a,
                            const float *norm_b, const float *denorm_A, const float *denorm_B, const float *s0,
                            int s0_batched, const float *Q, const float *h0, int h0_batched, int B, int T,
                            float *traj, float *h_final) {
    net_t N;
    const float *wih[8], *whh[8], *bih[8], *bhh[8];
    N.net_type = net_type; N.n_in = 1 + n_state_in; N.n_layers = n_layers; N.n_out = n_out; N.h = hsz;
    const float *p = weights;
    int in_len = N.n_in, Htot = 0, Hmax = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int H = hsz[l];
        if (net_type == 0) {
            wih[l] = p; p += (size_t)3 * H * in_len;
            whh[l] = p; p += (size_t)3 * H * H;
            bih[l] = p; p += 3 * H;
            bhh[l] = p; p += 3 * H;
        } else {
            wih[l] = p; p += (size_t)H * in_len;
            bih[l] = p; p += H;
            whh[l] = NULL; bhh[l] = NULL;
        }
        in_len = H; Htot += H; if (H > Hmax) Hmax = H;
    }
    N.w_ih = wih; N.w_hh = whh; N.b_ih = bih; N.b_hh = bhh;
    N.w_out = p; p += (size_t)n_out * in_len; N.b_out = p; N.h_max = Hmax;
#pragma omp parallel
    {
        float *scratch = (float *)malloc(sizeof(float) * (size_t)(16 * (Hmax > 64 ? Hmax : 64) + 64));
        float *h = (float *)malloc(sizeof(float) * (size_t)(Htot > 0 ? Htot : 1));
        float x[64], y[64];
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b) {
            const float *s = s0 + (s0_batched ? (size_t)b * 6 : 0);
            float *row = traj + (size_t)b * (T + 1) * 6;
            memcpy(row, s, 6 * sizeof(float));
            if (net_type == 0) memcpy(h, h0 + (h0_batched ? (size_t)b * Htot : 0), sizeof(float) * Htot);
            /* normalise the initial state features (:271,279) */
            for (int i = 0; i < n_state_in; ++i) x[1 + i] = norm_a[1 + i] * s[in_idx[i]] + norm_b[1 + i];
            /* autoregression_loop.run takes the helper only in its general loop: with horizon == 1 the "0th iteration"
             * branch hands out the raw network output (autoregression.py:49-70) */
            const int diff = g_diff.on && T > 1;
            float sp[8];
            if (diff) /* dmah.set_starting_point(normalize_state(s)) (predictor_autoregressive_neural.py:276-277) */
                for (int j = 0; j < n_out; ++j) sp[j] = g_diff.on_a[j] * s[out_idx[j]] + g_diff.on_b[j];
            int has_angle = 0, has_sin = 0, has_cos = 0;
            for (int j = 0; j < n_out; ++j) {
                if (out_idx[j] == IDX_ANGLE) has_angle = 1;
                if (out_idx[j] == IDX_SIN) has_sin = 1;
                if (out_idx[j] == IDX_COS) has_cos = 1;
            }
            for (int t = 0; t < T; ++t) {
                x[0] = norm_a[0] * Q[(size_t)b * T + t] + norm_b[0];
                net_step(&N, x, h, y, scratch);
                if (diff) { /* autoregression.py:149-154 */
                    for (int j = 0; j < n_out; ++j) { sp[j] = sp[j] + (g_diff.p1[j] * y[j] + g_diff.p2[j]); y[j] = sp[j]; }
                }
                float *o = row + (size_t)(t + 1) * 6;
                for (int j = 0; j < 6; ++j) o[j] = 0.0f;
                for (int j = 0; j < n_out; ++j) o[out_idx[j]] = denorm_A[j] * y[j] + denorm_B[j];
                /* predictor_output_augmentation._augment (predictors_customization.py:120-139) */
                if (!has_angle && has_sin && has_cos) o[IDX_ANGLE] = atan2f(o[IDX_SIN], o[IDX_COS]);
                if (has_angle && !has_sin) o[IDX_SIN] = sinf(o[IDX_ANGLE]);
                if (has_angle && !has_cos) o[IDX_COS] = cosf(o[IDX_ANGLE]);
                if (diff) for (int i = 0; i < n_state_in; ++i) x[1 + i] = sp[g_diff.out_to_in[i]];
                else for (int i = 0; i < n_state_in && i < n_out; ++i) x[1 + i] = y[i]; /* feed back normalised output */
            }
            if (h_final && net_type == 0) memcpy(h_final + (size_t)b * Htot, h, sizeof(float) * Htot);
        }
        free(h); free(scratch);
    }
}

/* ---------------------------------------------------------------------------------------
 * The plant: CartPole.update_state (CartPole/__init__.py:283-324), one tick of dt_simulation.
 * State lives in a float32 array (self.s); every scalar read from it is float32, the time step is a
 * Python float, second derivatives are float64 results of the numba-compiled _cartpole_ode.
 * ------------------------------------------------------------------------------------- */

/* _cartpole_ode as numba types it for CartPole.cartpole_ode (CartPole/__init__.py:342-343 ->
 * cartpole_equations.py:165-179 -> :44-105): (ca, sa, angleD, positionD) float32 scalars from self.s,
 * u float32 (Q2u, :160-164), k / m_cart / g / J_fric / M_fric 0-d float32 arrays, L and m_pole Python
 * floats (float(L), float(m_pole)).  float32 x float32 products stay float32; anything touching a
 * float64 operand, an integer literal or `**` promotes to float64. */
static inline void ode_plant(float ca, float sa, float angleD, float positionD, float u, const float *ph, double L,
                             double m_pole, double *angleDD, double *positionDD) {
    const float k = ph[PH_K], m_cart = ph[PH_MCART], g = ph[PH_G], J_fric = ph[PH_JFRIC], M_fric = ph[PH_MFRIC];
    const double kp1 = (double)(k + 1.0f);                    /* float32 (pinned against the live plant) */
    const float ca2 = ca * ca, aD2 = angleD * angleD;        /* x ** 2 of a float32 scalar stays float32 */
    const double A = kp1 * ((double)m_cart + m_pole) - m_pole * (double)ca2;
    const float F_fric = (-M_fric) * positionD;
    const float T_fric = (-J_fric) * angleD;
    const double L_half = L / 2.0;
    const double pDD = (((m_pole * (double)g) * (double)sa) * (double)ca + (double)(T_fric * ca) / L_half
                        + kp1 * (-(((m_pole * L_half) * (double)aD2) * (double)sa) + (double)F_fric + (double)u)) / A;
    *positionDD = pDD;
    *angleDD = ((double)(g * sa) + pDD * (double)ca + (double)T_fric / (m_pole * L_half)) / (kp1 * L_half);
}

/* python math.fmod form (CartPole/_CartPole_mathematical_helpers.py:13-21) on a float32 angle */
static inline double wrap_angle_rad_py(double angle) { return wrap_fmod(angle); }

/* One plant tick with the stored second derivatives (computed at the end of the previous tick with the
 * control that was active then): cartpole_integration (Euler-Cromer, cartpole_equations.py:368-378),
 * edge_bounce (:341-347 via CartPole/__init__.py:460-470), cos/sin of the UNWRAPPED angle (:329-331),
 * wrap (:333-334).  Each assignment into self.s rounds to float32. */
static inline void plant_integrate(float *s, double angleDD, double positionDD, double dt, const float *ph, double L) {
    const float angle = s[IDX_ANGLE], angleD = s[IDX_ANGLED], position = s[IDX_POS], positionD = s[IDX_POSD];
    const double angleD_next = (double)angleD + angleDD * dt;
    const double positionD_next = (double)positionD + positionDD * dt;
    const double angle_next = (double)angle + angleD_next * dt;
    const double position_next = (double)position + positionD_next * dt;
    s[IDX_ANGLE] = (float)angle_next; s[IDX_ANGLED] = (float)angleD_next;
    s[IDX_POS] = (float)position_next; s[IDX_POSD] = (float)positionD_next;
    /* edge_bounce_numba(angle, cos(angle), angleD, position, positionD, dt, L=float(L)) on the float32 values */
    const float thl = ph[PH_TRACK_HALF];
    if (s[IDX_POS] >= thl || -s[IDX_POS] >= thl) {
        const float a = s[IDX_ANGLE], aD = s[IDX_ANGLED], p = s[IDX_POS], pD = s[IDX_POSD];
        const float ca = cosf(a);
        const double aD2 = (double)aD - 2.0 * (double)(pD * ca) / (0.5 * L);
        const double a2 = (double)a + aD2 * dt;
        const float pD2 = -pD;
        const double p2 = (double)p + (double)pD2 * dt;
        s[IDX_ANGLE] = (float)a2; s[IDX_ANGLED] = (float)aD2; s[IDX_POS] = (float)p2; s[IDX_POSD] = pD2;
    }
    s[IDX_COS] = cosf(s[IDX_ANGLE]);
    s[IDX_SIN] = sinf(s[IDX_ANGLE]);
    s[IDX_ANGLE] = (float)wrap_angle_rad_py((double)s[IDX_ANGLE]);
}

/* One controller period of the plant: second derivatives from (s, Q) (set_cartpole_state_at_t0 / Update_Q ->
 * Q2u -> cartpole_ode, CartPole/__init__.py:317-321,883-885), then n_sim ticks; the derivatives are refreshed
 * at the end of every tick with the same Q (:317-321).  ticks_out: [n_sim][6] state after each tick or NULL;
 * dd_out: [n_sim + 1][2] second derivatives before tick 0 and after each tick, or NULL. */
void cps_oracle_plant_period(float *s, float Q, int n_sim, double dt_sim, const float *ph, double L, double m_pole,
                             float *ticks_out, double *dd_out) {
    const float u = ph[PH_UMAX] * Q;
    double aDD, pDD;
    ode_plant(s[IDX_COS], s[IDX_SIN], s[IDX_ANGLED], s[IDX_POSD], u, ph, L, m_pole, &aDD, &pDD);
    if (dd_out) { dd_out[0] = aDD; dd_out[1] = pDD; }
    for (int i = 0; i < n_sim; ++i) {
        plant_integrate(s, aDD, pDD, dt_sim, ph, L);
        ode_plant(s[IDX_COS], s[IDX_SIN], s[IDX_ANGLED], s[IDX_POSD], u, ph, L, m_pole, &aDD, &pDD);
        if (ticks_out) memcpy(ticks_out + (size_t)i * 6, s, 6 * sizeof(float));
        if (dd_out) { dd_out[2 * (i + 1)] = aDD; dd_out[2 * (i + 1) + 1] = pDD; }
    }
}
