/* fw_oracle.c — CPU restatement of the reference's fixed-wing env step, in plain C.
 *
 * TEST INFRASTRUCTURE.  This file is the parity ORACLE for libfwb200.so: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  It is never linked into, imported by, or
 * used as a fallback for the product path.
 *
 * What it restates (all paths relative to /root/reference/magpie/libs/):
 *   pyfly/pyfly/pyfly.py            PyFly.step :1358-1420, _dynamics :1450-1482, _forces :1484-1643,
 *                                   _f_*_dot :1645-1747, _rot_b_v :1749-1803, _calculate_airspeed_factors :1830-1850,
 *                                   _set_states_from_ode_solution :1852-1881, Variable.apply_conditions :114-133,
 *                                   ControlVariable :264-382, Actuation :453-655, AttitudeQuaternion :658-748,
 *                                   Wind :751-876, PyFly.reset :1262-1311
 *   pyfly/pyfly/dryden.py           DrydenGustModel.simulate :193-261 + Filter.simulate :22-39
 *   fixed-wing-gym/gym_fixed_wing/fixed_wing.py   reset :414-481, step :483-628, linear_action_scaling :630-652,
 *                                   sample_target :654-746, get_reward :941-1111, get_observation :1113-1262,
 *                                   _get_error/_get_angle_dist/_get_goal_status :1318-1361, _get_next_target :1363-1471,
 *                                   get_metric :1644-1736
 *   stable-baselines3/stable_baselines3/common/buffers.py   compute_returns_and_advantage :304-333
 * Third-party arithmetic that is NOT under /root/reference (requirements-magpie.txt pins scipy==1.6.0,
 * numpy==1.19.5; this container has scipy 1.18.1 / numpy 2.3.5): scipy.integrate.solve_ivp(method RK45) as
 * called at pyfly.py:1393-1395 — restated from scipy/integrate/_ivp/rk.py (rk_step, RungeKutta._step_impl,
 * RK45 tableau), common.py (norm, select_initial_step), base.py (OdeSolver.step); and scipy.signal.lsim's
 * linear-interpolation recurrence (_ltisys.py) whose discretised matrices arrive precomputed in FwConfig.filt.
 *
 * Pinning: tests/test_oracle_golden.py checks this file against fixtures recorded from the live reference
 * (tests/golden/make_golden.py) and against the reference's own golden PID evaluation
 * (examples/evaluations/eval_res_PID_none.npy).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/fwb200.h"

#define NY FW_NY

typedef struct FwoEnv {
    FwConfig cfg;
    double gam[9];   /* pyfly.py:1099-1116 */
    double ar;       /* pyfly.py:1119 */

    /* ---- PyFly state objects (.value) ---- */
    double quat[4];        /* state["attitude"].value (normalised at every write, pyfly.py:1859) */
    double omega[3], pos[3], vel[3];
    double act_val[3], act_dot[3];  /* elevon_right, elevon_left, throttle (dynamics order, pyfly_config.json) */
    double elev, ail;               /* derived elevator / aileron (pyfly.py:485-492) */
    double roll, pitch, yaw, Va, alpha, beta;
    double cmd_dyn[3];     /* constrained commands of the dynamics states */
    double cmd_in[3];      /* constrained elevator, aileron, throttle commands */
    double wind[3];        /* Wind.steady */
    int cur_sim_step;
    /* .history[-1] of the variables the observation reads (differs from .value after a failed step) */
    double h_roll, h_pitch, h_Va, h_omega[3], h_alpha, h_beta;

    /* ---- turbulence: whole episode pre-generated at reset like dryden.simulate ---- */
    double* turb;          /* [6][turb_len]: lin u,v,w then ang p,q,r */
    int turb_len;

    /* ---- gym env ---- */
    int steps_count, steps_for_target;
    double target[3];
    int tcls[3];                  /* _target_props[k]["class"] (an injected target forces constant, fixed_wing.py:446-450) */
    double t_slope[3], t_amp[3], t_period[3], t_phase[3], t_bias[3];
    double* act_hist;  int n_act;  int act_is_f32;   /* history["action"], raw */
    double* cmd_hist;  int n_cmd;                    /* elevator/aileron/throttle .history["command"], [t][3] */
    double* err_hist;  int n_err;                    /* history["error"], [t][3] */
    int obs_hist_n;    /* length of the history["error"] the observation sees: the running episode's in step(), the ENDED
                          episode's inside reset() (self.history is replaced after the reset observation, fixed_wing.py:453-460) */
    int has_history;   /* self.history is not None: false until the first reset has completed */
    uint8_t* goal_hist; int n_goal;                  /* history["goal"], [t][4] roll pitch Va all */
    double* st_hist;   int n_st;                     /* state .history of roll pitch Va p q r alpha beta, [t][8] */
    double* tgt_hist;  int n_tgt;                    /* history["target"], [t][3] */
    /* target class attitude_angular: omega_p/q/r as target states 3..5 (same row counts as the base histories) */
    double atarget[3];
    double* aerr_hist; double* atgt_hist; uint8_t* agoal_hist;   /* [t][3] each */
    double ametrics[FW_NMETRIC_ANG];
    double prev_shaping[3]; int has_prev_shaping[3];   /* self.prev_shaping per function class (None after reset) */
    int goal_achieved;                                /* self.goal_achieved: set once, never cleared by reset */
    /* ---- waypoint head (simple_train.py:197-702) ---- */
    const double* wp_tasks; int wp_n_tasks, wp_len, wp_task, wp_pos;   /* [n_tasks][wp_len][FW_WP_ROW], not owned */
    double wp_goal[3];
    double ep_return;
    double metrics[FW_NMETRIC];
    int term_code;
    int last_nfev, last_natt;
    uint64_t episode;
    uint64_t ep_seed;      /* Philox key of the running episode: cfg.seed when it was reset (fwo_set_config) */
    double Jy0;            /* self.I[1, 1], fixed at construction (pyfly.py:1086-1096): parameter randomisation never reaches it */
    int64_t env_id;
} FwoEnv;

/* ------------------------------------------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11, "Parallel random numbers: as easy as 1, 2, 3"); constants from Random123. */
void fwo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout shared with the CUDA path (DESIGN.md "RNG"): counter = (env_id lo32, episode lo32, purpose|env hi,
 * block); two 53-bit uniforms per block. */
enum { FWO_RNG_RESET = 0, FWO_RNG_NOISE = 1, FWO_RNG_RESAMPLE = 2, FWO_RNG_ACTION = 3, FWO_RNG_OBS = 4,
       FWO_RNG_OBS_INIT = 5, FWO_RNG_MODEL = 6 };

static void rng_block(uint64_t seed, int64_t env_id, uint64_t episode, uint32_t purpose, uint32_t block,
                      uint32_t out[4]) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)env_id, (uint32_t)episode,
                       (purpose << 28) | ((uint32_t)((uint64_t)env_id >> 32) & 0x0FFFFFFFu), block};
    fwo_philox4x32(ctr, key, out);
}

static double u53(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi) << 21) ^ (((uint64_t)lo) >> 11)) * (1.0 / 9007199254740992.0);
}

static double rng_uniform(const FwoEnv* e, uint32_t purpose, int idx) {
    uint32_t r[4];
    rng_block(e->ep_seed, e->env_id, e->episode, purpose, (uint32_t)(idx >> 1), r);
    return (idx & 1) ? u53(r[2], r[3]) : u53(r[0], r[1]);
}

/* four unit normals for turbulence sample k (Box-Muller on 32-bit uniforms) */
void fwo_noise4(uint64_t seed, int64_t env_id, uint64_t episode, uint32_t k, double out[4]) {
    uint32_t r[4];
    rng_block(seed, env_id, episode, FWO_RNG_NOISE, k, r);
    for (int i = 0; i < 2; ++i) {
        double u1 = ((double)r[2 * i] + 0.5) * (1.0 / 4294967296.0);
        double u2 = ((double)r[2 * i + 1] + 0.5) * (1.0 / 4294967296.0);
        double rad = sqrt(-2.0 * log(u1));
        out[2 * i] = rad * cos(6.283185307179586476925 * u2);
        out[2 * i + 1] = rad * sin(6.283185307179586476925 * u2);
    }
}

/* ------------------------------------------------------------------------------------------------------------- */
static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static double sgn(double v) { return (v > 0) - (v < 0); }   /* np.sign: sign(0) == 0 */

static void gammas(FwoEnv* e) {
    const FwConfig* c = &e->cfg;
    /* I = [[Jx,0,-Jxz],[0,Jy,0],[-Jxz,0,Jz]]  (pyfly.py:1086-1096) */
    double I00 = c->Jx, I11 = c->Jy, I22 = c->Jz, I02 = -c->Jxz;
    double* g = e->gam;
    g[0] = I00 * I22 - I02 * I02;
    g[1] = (fabs(I02) * (I00 - I11 + I22)) / g[0];
    g[2] = (I22 * (I22 - I11) + I02 * I02) / g[0];
    g[3] = I22 / g[0];
    g[4] = fabs(I02) / g[0];
    g[5] = (I22 - I00) / I11;
    g[6] = fabs(I02) / I11;
    g[7] = ((I00 - I11) * I00 + I02 * I02) / g[0];
    g[8] = I00 / g[0];
    e->ar = c->b * c->b / c->S_wing;
}

/* _rot_b_v, quaternion branch (pyfly.py:1780-1800) */
static void rot_quat(const double q[4], double R[3][3]) {
    double e0 = q[0], e1 = q[1], e2 = q[2], e3 = q[3];
    R[0][0] = -1 + 2 * (e0 * e0 + e1 * e1); R[0][1] = 2 * (e1 * e2 + e3 * e0); R[0][2] = 2 * (e1 * e3 - e2 * e0);
    R[1][0] = 2 * (e1 * e2 - e3 * e0); R[1][1] = -1 + 2 * (e0 * e0 + e2 * e2); R[1][2] = 2 * (e2 * e3 + e1 * e0);
    R[2][0] = 2 * (e1 * e3 + e2 * e0); R[2][1] = 2 * (e2 * e3 - e1 * e0); R[2][2] = -1 + 2 * (e0 * e0 + e3 * e3);
}

/* _rot_b_v, Euler branch (pyfly.py:1757-1777) */
static void rot_euler(double phi, double th, double psi, double R[3][3]) {
    double cph = cos(phi), sph = sin(phi), cth = cos(th), sth = sin(th), cps = cos(psi), sps = sin(psi);
    R[0][0] = cth * cps; R[0][1] = cth * sps; R[0][2] = -sth;
    R[1][0] = sph * sth * cps - cph * sps; R[1][1] = sph * sth * sps + cph * cps; R[1][2] = sph * cth;
    R[2][0] = cph * sth * cps + sph * sps; R[2][1] = cph * sth * sps - sph * cps; R[2][2] = cph * cth;
}

static const double* turb_col(const FwoEnv* e, int which /*0 lin, 1 ang*/, int k, double buf[3]) {
    if (!e->cfg.turbulence) { buf[0] = buf[1] = buf[2] = 0; return buf; }
    if (k >= e->turb_len) k = e->turb_len - 1;  /* the reference would re-simulate; never reached with L = steps_max */
    for (int i = 0; i < 3; ++i) buf[i] = e->turb[(which * 3 + i) * e->turb_len + k];
    return buf;
}

/* _calculate_airspeed_factors (pyfly.py:1830-1850). R is the body<-vehicle rotation of either branch. */
static void airspeed(const FwoEnv* e, double R[3][3], const double vel[3], double* Va, double* alpha, double* beta) {
    double tl[3];
    turb_col(e, 0, e->cur_sim_step, tl);
    double a[3];
    for (int i = 0; i < 3; ++i) {
        double w = R[i][0] * e->wind[0] + R[i][1] * e->wind[1] + R[i][2] * e->wind[2] + tl[i];
        a[i] = vel[i] - w;
    }
    *Va = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    *alpha = atan2(a[2], a[0]);
    *beta = asin(a[1] / *Va);
}

/* Actuation.set_states (pyfly.py:471-492) + ControlVariable.apply_conditions (:312-328) */
static void set_actuators(FwoEnv* e, const double* v6) {
    const FwConfig* c = &e->cfg;
    e->act_val[0] = clipd(v6[0], c->elevon_min, c->elevon_max);
    e->act_val[1] = clipd(v6[1], c->elevon_min, c->elevon_max);
    e->act_val[2] = clipd(v6[2], c->throttle_min, c->throttle_max);
    e->act_dot[0] = clipd(v6[3], -c->elevon_dot_max, c->elevon_dot_max);
    e->act_dot[1] = clipd(v6[4], -c->elevon_dot_max, c->elevon_dot_max);
    e->act_dot[2] = v6[5];   /* throttle: dot_max is None */
    /* _map_elevon_to_elevail (pyfly.py:651-655) */
    e->ail = (-e->act_val[0] + e->act_val[1]) / 2;
    e->elev = (e->act_val[0] + e->act_val[1]) / 2;
}

/* _set_states_from_ode_solution(save=False) (pyfly.py:1852-1881). Returns a FwTermCode on ConstraintException. */
static int write_back(FwoEnv* e, const double* y) {
    const FwConfig* c = &e->cfg;
    double n = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3]);
    for (int i = 0; i < 4; ++i) e->quat[i] = y[i] / n;
    for (int i = 0; i < 3; ++i) {
        if (y[4 + i] < c->omega_con_min[i] || y[4 + i] > c->omega_con_max[i]) return FW_TERM_OMEGA_P + i;
        e->omega[i] = y[4 + i];
    }
    for (int i = 0; i < 3; ++i) e->pos[i] = y[7 + i];
    for (int i = 0; i < 3; ++i) e->vel[i] = y[10 + i];
    set_actuators(e, y + 13);
    return 0;
}

/* PyFly._dynamics (pyfly.py:1450-1482) with _forces (:1484-1643) inlined. Returns FwTermCode (0 = ok). */
static int dynamics(FwoEnv* e, int t_positive, const double* y, double* dy) {
    const FwConfig* c = &e->cfg;
    e->last_nfev++;
    if (t_positive) {
        int rc = write_back(e, y);
        if (rc) return rc;
    }
    const double* att = y;                   /* un-normalised (pyfly.py:1464) */
    const double* om = e->omega;             /* read back from the state objects (pyfly.py:1466-1470) */
    const double* vel = e->vel;
    double elevator = e->elev, aileron = e->ail, rudder = 0.0, throttle = e->act_val[2];

    /* ---- _forces ---- */
    double p = om[0], q = om[1], r = om[2];
    if (c->turbulence) {
        double ta[3];
        turb_col(e, 1, e->cur_sim_step, ta);
        p -= ta[0]; q -= ta[1]; r -= ta[2];
    }
    double R[3][3], Va, alpha, beta;
    rot_quat(att, R);
    airspeed(e, R, vel, &Va, &alpha, &beta);
    /* state["Va"].apply_conditions: constraint first, then clip (pyfly.py:121-128) */
    if (c->va_con_max > 0 && Va > c->va_con_max) return FW_TERM_VA;
    if (Va < c->va_value_min) Va = c->va_value_min;

    double pre_fac = 0.5 * c->rho * (Va * Va) * c->S_wing;
    double e0 = att[0], e1 = att[1], e2 = att[2], e3 = att[3];
    double mg = c->mass * c->g;
    double fg[3] = {mg * (2 * (e1 * e3 - e2 * e0)), mg * (2 * (e2 * e3 + e1 * e0)),
                    mg * (e3 * e3 + e0 * e0 - e1 * e1 - e2 * e2)};

    double C_L_alpha_lin = c->C_L_0 + c->C_L_alpha * alpha;
    double ex1 = exp(-c->M * (alpha - c->a_0)), ex2 = exp(c->M * (alpha + c->a_0));
    double sigma = (1 + ex1 + ex2) / ((1 + ex1) * (1 + ex2));
    double sa = sin(alpha), ca = cos(alpha), sg = sgn(alpha);
    double C_L = (1 - sigma) * C_L_alpha_lin + sigma * (2 * sg * (sa * sa) * ca);
    double f_lift = pre_fac * (C_L + c->C_L_q * c->c / (2 * Va) * q + c->C_L_delta_e * elevator);
    double C_D_alpha = c->C_D_p + (1 - sigma) * (C_L_alpha_lin * C_L_alpha_lin) / (M_PI * c->e_oswald * e->ar)
                       + sigma * (2 * sg * (sa * sa * sa));
    double C_D_beta = c->C_D_beta1 * beta + c->C_D_beta2 * (beta * beta);
    double f_drag = pre_fac * (C_D_alpha + C_D_beta + c->C_D_q * c->c / (2 * Va) * q
                               + c->C_D_delta_e * (elevator * elevator));
    double C_m = (1 - sigma) * (c->C_m_0 + c->C_m_alpha * alpha) + sigma * (c->C_m_fp * sg * (sa * sa));
    /* sic: span b, not chord c, in the damping term (pyfly.py:1579) */
    double m = pre_fac * c->c * (C_m + c->C_m_q * c->b / (2 * Va) * q + c->C_m_delta_e * elevator);
    double f_y = pre_fac * (c->C_Y_0 + c->C_Y_beta * beta + c->C_Y_p * c->b / (2 * Va) * p
                            + c->C_Y_r * c->b / (2 * Va) * r + c->C_Y_delta_a * aileron + c->C_Y_delta_r * rudder);
    double l = pre_fac * c->b * (c->C_l_0 + c->C_l_beta * beta + c->C_l_p * c->b / (2 * Va) * p
                                 + c->C_l_r * c->b / (2 * Va) * r + c->C_l_delta_a * aileron
                                 + c->C_l_delta_r * rudder);
    double n = pre_fac * c->b * (c->C_n_0 + c->C_n_beta * beta + c->C_n_p * c->b / (2 * Va) * p
                                 + c->C_n_r * c->b / (2 * Va) * r + c->C_n_delta_a * aileron
                                 + c->C_n_delta_r * rudder);
    double Rs[3][3];
    rot_euler(0.0, alpha, beta, Rs);      /* pyfly.py:1617-1620 */
    double fs[3] = {-f_drag, f_y, -f_lift}, f_aero[3];
    for (int i = 0; i < 3; ++i) f_aero[i] = Rs[i][0] * fs[0] + Rs[i][1] * fs[1] + Rs[i][2] * fs[2];
    double Vd = Va + throttle * (c->k_motor - Va);
    double f_prop = 0.5 * c->rho * c->S_prop * c->C_prop * Vd * (Vd - Va);
    double kt = c->k_Omega * throttle;
    double tau_prop = -c->k_T_P * (kt * kt);
    double f[3] = {f_prop + fg[0] + f_aero[0], 0 + fg[1] + f_aero[1], 0 + fg[2] + f_aero[2]};
    double tau[3] = {l + tau_prop, m + 0, n + 0};

    /* ---- _f_attitude_dot (pyfly.py:1645-1657): state omega, NOT the turbulence-corrected one ---- */
    double P = om[0], Q = om[1], Rr = om[2];
    dy[0] = 0.5 * (0 * att[0] + -P * att[1] + -Q * att[2] + -Rr * att[3]);
    dy[1] = 0.5 * (P * att[0] + 0 * att[1] + Rr * att[2] + -Q * att[3]);
    dy[2] = 0.5 * (Q * att[0] + -Rr * att[1] + 0 * att[2] + P * att[3]);
    dy[3] = 0.5 * (Rr * att[0] + Q * att[1] + -P * att[2] + 0 * att[3]);
    /* ---- _f_omega_dot (pyfly.py:1659-1683) ---- */
    const double* g = e->gam;
    dy[4] = g[1] * P * Q - g[2] * Q * Rr + g[3] * tau[0] + g[4] * tau[2];
    dy[5] = g[5] * P * Rr - g[6] * (P * P - Rr * Rr) + tau[1] / e->Jy0;
    dy[6] = g[7] * P * Q - g[1] * Q * Rr + g[4] * tau[0] + g[8] * tau[2];
    /* ---- _f_p_dot (pyfly.py:1706-1737) ---- */
    double T[3][3] = {
        {e1 * e1 + e0 * e0 - e2 * e2 - e3 * e3, 2 * (e1 * e2 - e3 * e0), 2 * (e1 * e3 + e2 * e0)},
        {2 * (e1 * e2 + e3 * e0), e2 * e2 + e0 * e0 - e1 * e1 - e3 * e3, 2 * (e2 * e3 - e1 * e0)},
        {2 * (e1 * e3 - e2 * e0), 2 * (e2 * e3 + e1 * e0), e3 * e3 + e0 * e0 - e1 * e1 - e2 * e2}};
    for (int i = 0; i < 3; ++i) dy[7 + i] = T[i][0] * vel[0] + T[i][1] * vel[1] + T[i][2] * vel[2];
    /* ---- _f_v_dot (pyfly.py:1685-1704) ---- */
    dy[10] = Rr * vel[1] - Q * vel[2] + f[0] / c->mass;
    dy[11] = P * vel[2] - Rr * vel[0] + f[1] / c->mass;
    dy[12] = Q * vel[0] - P * vel[1] + f[2] / c->mass;
    /* ---- Actuation.rhs (pyfly.py:519-543), coefficient rows from ControlVariable.coefs (:296-304) ---- */
    double w0 = c->elevon_omega0, z = c->elevon_zeta, it = 1 / c->throttle_tau;
    double c00[3] = {0, 0, -it}, c01[3] = {1, 1, 0}, c02[3] = {0, 0, it};
    double c10[3] = {-(w0 * w0), -(w0 * w0), 0}, c11[3] = {-2 * z * w0, -2 * z * w0, 0}, c12[3] = {w0 * w0, w0 * w0, 0};
    for (int i = 0; i < 3; ++i) {
        dy[13 + i] = e->act_val[i] * c00[i] + e->cmd_dyn[i] * c02[i] + e->act_dot[i] * c01[i];
        dy[16 + i] = e->act_val[i] * c10[i] + e->cmd_dyn[i] * c12[i] + e->act_dot[i] * c11[i];
    }
    return 0;
}

/* RMS norm (scipy/integrate/_ivp/common.py:63-65) */
static double rms(const double* x) {
    double s = 0;
    for (int i = 0; i < NY; ++i) s += x[i] * x[i];
    return sqrt(s) / sqrt((double)NY);
}

/* Dormand-Prince tableau (scipy/integrate/_ivp/rk.py, class RK45) */
/* C = [0, 1/5, 3/10, 4/5, 8/9, 1]: only enters through t>0 (pyfly.py:1461), so it is not stored */
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

/* scipy.integrate.solve_ivp(fun, (0, dt), y0) with RK45 defaults, as called at pyfly.py:1393-1395.
 * y is updated in place to sol.y[:, -1].  Returns FwTermCode (0 = ok). */
static int solve_ivp_rk45(FwoEnv* e, double* y) {
    const double rtol = e->cfg.rtol, atol = e->cfg.atol, t_bound = e->cfg.dt;
    double f[NY], K[7][NY], ynew[NY], tmp[NY], scale[NY];
    int rc;
    /* RungeKutta.__init__: self.f = fun(t0, y0) — t == 0, so no state write-back (pyfly.py:1461) */
    if ((rc = dynamics(e, 0, y, f))) return rc;
    /* select_initial_step (common.py:68-134), order = error_estimator_order = 4 */
    double h_abs;
    {
        double d0, d1, d2, h0, h1, y1[NY], f1[NY];
        for (int i = 0; i < NY; ++i) scale[i] = atol + fabs(y[i]) * rtol;
        for (int i = 0; i < NY; ++i) tmp[i] = y[i] / scale[i];
        d0 = rms(tmp);
        for (int i = 0; i < NY; ++i) tmp[i] = f[i] / scale[i];
        d1 = rms(tmp);
        h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        if (h0 > t_bound) h0 = t_bound;
        for (int i = 0; i < NY; ++i) y1[i] = y[i] + h0 * 1.0 * f[i];
        if ((rc = dynamics(e, 1, y1, f1))) return rc;
        for (int i = 0; i < NY; ++i) tmp[i] = (f1[i] - f[i]) / scale[i];
        d2 = rms(tmp) / h0;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(100 * h0, h1), t_bound);
    }
    double t = 0;
    while (t < t_bound) {          /* solve_ivp loop + OdeSolver.step (base.py:179-210) */
        /* RungeKutta._step_impl (rk.py:111-173) */
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        int accepted = 0, rejected = 0;
        double t_new = t;
        while (!accepted) {
            if (h_abs < min_step) return 0;   /* TOO_SMALL_STEP: solve_ivp returns status -1, pyfly ignores it */
            double h = h_abs;
            t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            e->last_natt++;
            /* rk_step (rk.py:14-73) */
            memcpy(K[0], f, sizeof(f));
            for (int s = 1; s < 6; ++s) {
                for (int i = 0; i < NY; ++i) {
                    double acc = 0;
                    for (int j = 0; j < s; ++j) acc += K[j][i] * RK_A[s][j];
                    tmp[i] = y[i] + acc * h;
                }
                if ((rc = dynamics(e, 1, tmp, K[s]))) return rc;
            }
            for (int i = 0; i < NY; ++i) {
                double acc = 0;
                for (int j = 0; j < 6; ++j) acc += K[j][i] * RK_B[j];
                ynew[i] = y[i] + h * acc;
            }
            if ((rc = dynamics(e, 1, ynew, K[6]))) return rc;
            for (int i = 0; i < NY; ++i) {
                double sc = atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol, acc = 0;
                for (int j = 0; j < 7; ++j) acc += K[j][i] * RK_E[j];
                tmp[i] = acc * h / sc;
            }
            double err = rms(tmp);
            if (err < 1) {
                double factor = (err == 0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                accepted = 1;
            } else {
                h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
                rejected = 1;
            }
        }
        t = t_new;
        memcpy(y, ynew, sizeof(ynew));
        memcpy(f, K[6], sizeof(f));
    }
    return 0;
}

/* Classical RK4 with n fixed substeps (throughput mode; NOT the reference integrator). Uses the same RHS with its
 * clip / constraint side effects. */
static int solve_rk4_fixed(FwoEnv* e, double* y) {
    int n = e->cfg.rk4_substeps > 0 ? e->cfg.rk4_substeps : 1, rc;
    double h = e->cfg.dt / n, k1[NY], k2[NY], k3[NY], k4[NY], tmp[NY];
    for (int s = 0; s < n; ++s) {
        if ((rc = dynamics(e, s > 0, y, k1))) return rc;
        for (int i = 0; i < NY; ++i) tmp[i] = y[i] + 0.5 * h * k1[i];
        if ((rc = dynamics(e, 1, tmp, k2))) return rc;
        for (int i = 0; i < NY; ++i) tmp[i] = y[i] + 0.5 * h * k2[i];
        if ((rc = dynamics(e, 1, tmp, k3))) return rc;
        for (int i = 0; i < NY; ++i) tmp[i] = y[i] + h * k3[i];
        if ((rc = dynamics(e, 1, tmp, k4))) return rc;
        for (int i = 0; i < NY; ++i) y[i] = y[i] + h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    }
    return 0;
}

/* Actuation.set_and_constrain_commands (pyfly.py:545-582) for inputs (elevator, aileron, throttle) and elevon
 * dynamics.  Appends to the command histories of elevator/aileron/throttle. */
static void set_commands(FwoEnv* e, const double a[3]) {
    const FwConfig* c = &e->cfg;
    double er = -1 * a[1] + a[0], el = a[1] + a[0];              /* _map_elevail_to_elevon :645-649 */
    e->cmd_dyn[0] = clipd(er, c->elevon_min, c->elevon_max);     /* set_command -> Variable.apply_conditions */
    e->cmd_dyn[1] = clipd(el, c->elevon_min, c->elevon_max);
    e->cmd_dyn[2] = clipd(a[2], c->throttle_min, c->throttle_max);
    double elev_c = (e->cmd_dyn[0] + e->cmd_dyn[1]) / 2, ail_c = (-e->cmd_dyn[0] + e->cmd_dyn[1]) / 2;
    /* elevator/aileron limits set by Actuation.finalize (pyfly.py:599-623) == act_lo/act_hi of the gym env */
    e->cmd_in[0] = clipd(elev_c, c->act_lo[0], c->act_hi[0]);
    e->cmd_in[1] = clipd(ail_c, c->act_lo[1], c->act_hi[1]);
    e->cmd_in[2] = e->cmd_dyn[2];
    if (e->n_cmd < c->steps_max + 1) {
        for (int i = 0; i < 3; ++i) e->cmd_hist[e->n_cmd * 3 + i] = e->cmd_in[i];
        e->n_cmd++;
    }
}

/* PyFly.step (pyfly.py:1358-1420). Returns 0 on success or the FwTermCode of the violated constraint. */
static int sim_step(FwoEnv* e, const double commands[3]) {
    const FwConfig* c = &e->cfg;
    int rc = 0;
    set_commands(e, commands);
    double y[NY];
    for (int i = 0; i < 4; ++i) y[i] = e->quat[i];
    for (int i = 0; i < 3; ++i) { y[4 + i] = e->omega[i]; y[7 + i] = e->pos[i]; y[10 + i] = e->vel[i]; }
    for (int i = 0; i < 3; ++i) { y[13 + i] = e->act_val[i]; y[16 + i] = e->act_dot[i]; }
    e->last_nfev = 0; e->last_natt = 0;
    rc = (c->integrator == FW_INT_RK45_SCIPY) ? solve_ivp_rk45(e, y) : solve_rk4_fixed(e, y);
    if (!rc) {
        /* _set_states_from_ode_solution(sol.y[:, -1], save=True): attitude, roll, pitch, yaw, omega_*, pos, vel,
         * actuators in this order; a raise leaves later variables without a new history entry */
        double n = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3]);
        for (int i = 0; i < 4; ++i) e->quat[i] = y[i] / n;
        double e0 = e->quat[0], e1 = e->quat[1], e2 = e->quat[2], e3 = e->quat[3];
        /* AttitudeQuaternion.as_euler_angle (pyfly.py:684-708) */
        e->roll = atan2(2 * (e0 * e1 + e2 * e3), e0 * e0 + e3 * e3 - e1 * e1 - e2 * e2);
        e->pitch = asin(2 * (e0 * e2 - e1 * e3));
        e->yaw = atan2(2 * (e0 * e3 + e1 * e2), e0 * e0 + e1 * e1 - e2 * e2 - e3 * e3);
        e->h_roll = e->roll; e->h_pitch = e->pitch;
        for (int i = 0; i < 3 && !rc; ++i) {
            if (y[4 + i] < c->omega_con_min[i] || y[4 + i] > c->omega_con_max[i]) { rc = FW_TERM_OMEGA_P + i; break; }
            e->omega[i] = y[4 + i];
            e->h_omega[i] = y[4 + i];
        }
        if (!rc) {
            for (int i = 0; i < 3; ++i) { e->pos[i] = y[7 + i]; e->vel[i] = y[10 + i]; }
            set_actuators(e, y + 13);
            /* pyfly.py:1398-1406: Euler branch of _rot_b_v */
            double R[3][3], Va, alpha, beta;
            rot_euler(e->roll, e->pitch, e->yaw, R);
            airspeed(e, R, e->vel, &Va, &alpha, &beta);
            if (c->va_con_max > 0 && Va > c->va_con_max) rc = FW_TERM_VA;
            else {
                if (Va < c->va_value_min) Va = c->va_value_min;
                e->Va = Va; e->alpha = alpha; e->beta = beta;
                e->h_Va = Va; e->h_alpha = alpha; e->h_beta = beta;
            }
        }
    }
    e->cur_sim_step += 1;
    return rc;
}

/* DrydenGustModel.simulate + Filter.simulate -> scipy lsim recurrence (dryden.py:193-261).  noise: [4][L] unit
 * normals; out: [6][L] (vel_lin rows then vel_ang rows). */
void fwo_dryden(const FwConfig* c, const double* noise, int L, double* out) {
    for (int fi = 0; fi < 6; ++fi) {
        const FwFilter* F = &c->filt[fi];
        const double* u = noise + (size_t)F->noise_row * L;
        double x[3] = {0, 0, 0}, xn[3];
        int n = F->order;
        for (int i = 0; i < L; ++i) {
            double ui = u[i] * c->turb_noise_scale;
            /* pyfly simulates blocks of turbulence_sim_length samples; every new block calls lsim with T[0] > 0, which
             * first steps the carried state over [0, T[0]] with zero input: x <- x expm(A^T T[0]) = x Ablk^m
             * (pyfly.py:870-871, dryden.py:30-36, scipy lsim) */
            const int block_start = c->turb_block_len > 0 && (i % c->turb_block_len) == 0;
            if (i > 0 && block_start) {
                for (int m = 0; m < i / c->turb_block_len; ++m) {
                    for (int a = 0; a < n; ++a) {
                        double sacc = 0;
                        for (int b = 0; b < n; ++b) sacc += x[b] * F->Ablk[b * n + a];
                        xn[a] = sacc;
                    }
                    for (int a = 0; a < n; ++a) x[a] = xn[a];
                }
            }
            if (i > 0 && !block_start) {
                double up = u[i - 1] * c->turb_noise_scale;
                /* xout[i] = xout[i-1] @ Ad + U[i-1] @ Bd0 + U[i] @ Bd1   (row-vector convention) */
                for (int a = 0; a < n; ++a) {
                    double s = 0;
                    for (int b = 0; b < n; ++b) s += x[b] * F->Ad[b * n + a];
                    xn[a] = s + up * F->Bd0[a] + ui * F->Bd1[a];
                }
                for (int a = 0; a < n; ++a) x[a] = xn[a];
            }
            double yv = 0;
            for (int a = 0; a < n; ++a) yv += x[a] * F->C[a];
            out[(size_t)fi * L + i] = yv + ui * F->D;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------- */
/* gym env */

/* numpy's pairwise summation for n <= 128 contiguous elements (numpy/core/src/umath/loops_utils.h, pairwise_sum):
 * n < 8 sequential, otherwise 8 running accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) + tail. */
static double np_sum_f64(const double* a, int n) {
    if (n < 8) { double s = 0; for (int i = 0; i < n; ++i) s += a[i]; return n ? s : 0.0; }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}
static float np_sum_f32(const float* a, int n) {
    if (n < 8) { float s = 0; for (int i = 0; i < n; ++i) s += a[i]; return n ? s : 0.0f; }
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

/* _get_error (fixed_wing.py:1318-1344): roll has wrap=True (pyfly_config.json) -> _get_angle_dist(target, value) */
static double py_mod(double a, double b) { double m = fmod(a, b); if (m != 0 && ((m < 0) != (b < 0))) m += b; return m; }
static double get_error(const FwoEnv* e, int k) {
    if (k >= 3) return e->atarget[k - 3] - e->omega[k - 3];      /* omega states do not wrap */
    if (k == 0) {
        double dist = py_mod(e->roll - e->target[0] + M_PI, 2 * M_PI) - M_PI;
        if (dist < -M_PI) dist += 2 * M_PI;
        return dist;
    }
    return e->target[k] - (k == 1 ? e->pitch : e->Va);
}

/* per angular state (attitude_angular targets with a bound take part in "all", fixed_wing.py:1346-1361) */
static void goal_status_ang(const FwoEnv* e, uint8_t ga[3]) {
    for (int a = 0; a < 3; ++a) ga[a] = e->cfg.ang_on ? (fabs(get_error(e, 3 + a)) <= e->cfg.ang_bound[a]) : 1;
}
static void goal_status(const FwoEnv* e, uint8_t g[4]) {
    g[3] = 1;
    for (int k = 0; k < 3; ++k) { g[k] = fabs(get_error(e, k)) <= e->cfg.tgt_bound[k]; g[3] &= g[k]; }
    if (e->cfg.ang_on) { uint8_t ga[3]; goal_status_ang(e, ga); g[3] &= ga[0] & ga[1] & ga[2]; }
}

/* _attitude_to_angular_rates (fixed_wing.py:1558-1642) for omega state a (0 p, 1 q, 2 r): reads the CURRENT targets and
 * simulator values; the `damping = 0.05` branches are overwritten unconditionally in the reference (kept so). */
static double attitude_to_angular_rate(const FwoEnv* e, int a) {
    const FwConfig* c = &e->cfg;
    const double max_vel = c->ang_max_vel[a];
    const double roll_angle = e->roll, pitch_angle = e->pitch;
    const double roll_error = get_error(e, 0), pitch_error = get_error(e, 1);
    const double q_w = cos(roll_angle), r_w = sin(roll_angle);
    const double max_pitch_change = max_vel * c->dt * (q_w + r_w);
    double res, damping;
    if (a == 0) {
        damping = fabs(roll_error / (0.5 * M_PI));
        const double q_roll = sin(roll_angle) * tan(pitch_angle) * e->atarget[1] * c->dt;
        const double r_roll = cos(roll_angle) * tan(pitch_angle) * e->atarget[2] * c->dt;
        res = clipd(-(roll_error - q_roll - r_roll) / c->dt, -max_vel, max_vel);
    } else if (a == 1) {
        damping = fabs(pitch_error / (0.5 * M_PI));
        if (max_pitch_change > fabs(pitch_error)) res = -pitch_error / (2 * q_w);
        else res = sgn(q_w) * max_vel * sgn(pitch_error);
    } else {
        damping = fabs(pitch_error / (0.5 * M_PI));
        if (max_pitch_change > fabs(pitch_error)) res = pitch_error / r_w;
        else res = -sgn(r_w) * max_vel * sgn(pitch_error);
    }
    if (isnan(damping)) damping = 0.05; else damping = fmin(1.0, damping);
    return clipd(e->atarget[a] + (res * damping - e->atarget[a]) * 1 / 20, -max_vel, max_vel);
}

/* sample_target (fixed_wing.py:654-746).  u12: four U[0,1) draws per target state in the order the reference consumes
 * them (initial value; slope, sign | amplitude, period, phase). */
static void sample_target(FwoEnv* e, const double u12[12]) {
    const FwConfig* c = &e->cfg;
    double val[3] = {e->roll, e->pitch, e->Va};
    e->steps_for_target = 0;
    for (int k = 0; k < 3; ++k) {
        const double* u = u12 + 4 * k;
        double low = c->tgt_low[k], high = c->tgt_high[k];
        if (!isnan(c->tgt_delta[k])) {
            low = fmax(low, val[k] - c->tgt_delta[k]);
            high = fmax(fmin(high, val[k] + c->tgt_delta[k]), low);
        }
        const double initial = low + (high - low) * u[0];   /* RandomState.uniform: low + (high-low)*random_sample() */
        e->tcls[k] = c->tgt_class[k];
        if (c->tgt_class[k] == FW_TGT_LINEAR) {
            double slope = c->tgt_slope_low[k] + (c->tgt_slope_high[k] - c->tgt_slope_low[k]) * u[1];
            if (u[2] < 0.5) slope *= -1;
            if (c->tgt_radians[k]) slope = slope * (M_PI / 180.0);
            e->t_slope[k] = slope;
        } else if (c->tgt_class[k] == FW_TGT_SINUSOIDAL) {
            double amp = c->tgt_amp_low[k] + (c->tgt_amp_high[k] - c->tgt_amp_low[k]) * u[1];
            if (c->tgt_radians[k]) amp = amp * (M_PI / 180.0);
            const double period = c->tgt_period_low[k] + (c->tgt_period_high[k] - c->tgt_period_low[k]) * u[2];
            const double phase = (0 + (2 * M_PI - 0) * u[3]) / (2 * M_PI / period);
            e->t_amp[k] = amp; e->t_period[k] = period; e->t_phase[k] = phase;
            e->t_bias[k] = initial - amp * sin(2 * M_PI / period * (e->steps_count + phase));
        }
        e->target[k] = initial;
    }
    if (c->ang_on) {
        /* fixed_wing.py:671-675, 741-746: the three are set to 0, then ALL re-derived from those zeros (the dict
         * comprehension is evaluated before the update) */
        double v[3];
        e->atarget[0] = e->atarget[1] = e->atarget[2] = 0;
        for (int a = 0; a < 3; ++a) v[a] = attitude_to_angular_rate(e, a);
        for (int a = 0; a < 3; ++a) e->atarget[a] = v[a];
    }
}

/* the twelve draws of one target sampling: Philox (purpose, block base) or the test override */
static void target_draws(const FwoEnv* e, uint32_t purpose, uint32_t block0, double u12[12]) {
    const FwConfig* c = &e->cfg;
    for (int i = 0; i < 12; ++i) {
        if (!isnan(c->rng_u_override)) { u12[i] = c->rng_u_override; continue; }
        uint32_t r[4];
        rng_block(e->ep_seed, e->env_id, e->episode, purpose, block0 + (uint32_t)(i >> 1), r);
        u12[i] = (i & 1) ? u53(r[2], r[3]) : u53(r[0], r[1]);
    }
}

/* _get_next_target (fixed_wing.py:1363-1471) */
static void next_target(FwoEnv* e) {
    const FwConfig* c = &e->cfg;
    double res[3] = {e->target[0], e->target[1], e->target[2]};
    double ares[3] = {0, 0, 0};
    if (c->ang_on) for (int a = 0; a < 3; ++a) ares[a] = attitude_to_angular_rate(e, a);   /* every value from the OLD targets */
    for (int k = 0; k < 3; ++k) {
        if (e->tcls[k] == FW_TGT_LINEAR) res[k] = e->target[k] + e->t_slope[k] * c->dt;
        else if (e->tcls[k] == FW_TGT_SINUSOIDAL)
            res[k] = e->t_amp[k] * sin(2 * M_PI / e->t_period[k] * (e->steps_count + e->t_phase[k])) + e->t_bias[k];
    }
    if (e->tcls[2] == FW_TGT_COMPENSATE) {
        /* pitch target seen by the Va law: the target itself, or the bias of a sinusoidal one (:1381-1384) */
        double pitch_tar = (e->tcls[1] == FW_TGT_SINUSOIDAL) ? e->t_bias[1] : e->target[1], va_target = e->target[2];
        if (pitch_tar <= -2.5 * (M_PI / 180.0)) {
            double va_end = 28.434 - 40.0841 * pitch_tar, slope;
            if (va_target <= va_end) {
                double s = (va_target < va_end * 0.95) ? 1.0 : 1 - va_target / (va_end * 1.5);
                slope = 7 * fmax(0.0, s);
            } else slope = 0;
            res[2] = va_target + (slope * (-e->target[1]) - 0.25) * c->dt;
        } else if (pitch_tar >= 5 * (M_PI / 180.0)) {
            double va_end = 26.27 - 41.2529 * pitch_tar;
            if (va_target > va_end) {
                if (e->steps_for_target < 750) res[2] = va_target + (va_end - va_target) * 1 / 150;
                else res[2] = va_end;
            } else res[2] = va_target;
        }
    }
    /* wrap of roll targets beyond pi (fixed_wing.py:1465-1469) */
    if (fabs(res[0]) > M_PI) res[0] = sgn(res[0]) * (py_mod(fabs(res[0]), M_PI) - M_PI);
    for (int k = 0; k < 3; ++k) e->target[k] = res[k];
    if (c->ang_on) for (int a = 0; a < 3; ++a) e->atarget[a] = ares[a];
}

/* sum(|diff|) of one column of the trailing `window` rows of hist[n][3], accumulated in float32
 * (fixed_wing.py:1198-1228: np.sum(np.abs(np.diff(...)), dtype=np.float32)) */
static double delta_feature_f32(const double* hist, int n, int col, int window, int is_f32) {
    int lo = n - window < 0 ? 0 : n - window;
    float s = 0;
    for (int t = lo + 1; t < n; ++t) {
        float d;
        if (is_f32) d = fabsf((float)hist[t * 3 + col] - (float)hist[(t - 1) * 3 + col]);
        else d = (float)fabs(hist[t * 3 + col] - hist[(t - 1) * 3 + col]);
        s += d;
    }
    return (double)s;
}

/* observation.noise (fixed_wing.py:1246-1247): every entry += N(mean, var).  MT19937 cannot be reproduced: the draws
 * come from the Philox stream (purpose OBS, block = 4 * steps_count + b), 14 Box-Muller normals per observation. */
static void add_obs_noise(const FwoEnv* e, double* obs, int dim) {
    const FwConfig* c = &e->cfg;
    if (!(c->obs_noise_std > 0) && c->obs_noise_mean == 0) return;
    for (int b = 0; b * 4 < dim; ++b) {
        uint32_t r[4];
        rng_block(e->ep_seed, e->env_id, e->episode, FWO_RNG_OBS, (uint32_t)(e->steps_count * 32 + b), r);
        for (int i = 0; i < 2; ++i) {
            double u1 = ((double)r[2 * i] + 0.5) * (1.0 / 4294967296.0);
            double u2 = ((double)r[2 * i + 1] + 0.5) * (1.0 / 4294967296.0);
            double rad = sqrt(-2.0 * log(u1));
            double z0 = rad * cos(6.283185307179586476925 * u2), z1 = rad * sin(6.283185307179586476925 * u2);
            int j = b * 4 + i * 2;
            if (j < dim) obs[j] += c->obs_noise_mean + c->obs_noise_std * z0;
            if (j + 1 < dim) obs[j + 1] += c->obs_noise_mean + c->obs_noise_std * z1;
        }
    }
}

static void push_state_history(FwoEnv* e) {
    double* p = e->st_hist + (size_t)e->n_st * 8;
    p[0] = e->roll; p[1] = e->pitch; p[2] = e->Va; p[3] = e->omega[0]; p[4] = e->omega[1]; p[5] = e->omega[2];
    p[6] = e->alpha; p[7] = e->beta;
    e->n_st++;
}

/* sum |diff| over hist[lo..hi) of one column, accumulated in float32 (np.sum(..., dtype=np.float32)) */
static double window_feature_f32(const double* hist, int lo, int hi, int col, int is_f32) {
    float s = 0;
    for (int t = lo + 1; t < hi; ++t) {
        float d;
        if (is_f32) d = fabsf((float)hist[t * 3 + col] - (float)hist[(t - 1) * 3 + col]);
        else d = (float)fabs(hist[t * 3 + col] - hist[(t - 1) * 3 + col]);
        s += d;
    }
    return (double)s;
}

/* np.sum(history["error"][name][start:stop]), absolute indices clamped to [0, n] like a python slice; n = len(history) */
static double err_slice_sum(const FwoEnv* e, int n, int k, int start, int stop) {
    if (start < 0) start = 0;
    if (stop < 0) stop = 0;
    if (start > n) start = n;
    if (stop > n) stop = n;
    if (stop <= start) return 0.0;
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(stop - start));
    for (int t = start; t < stop; ++t) tmp[t - start] = k < 3 ? e->err_hist[(size_t)t * 3 + k] : e->aerr_hist[(size_t)t * 3 + k - 3];
    const double s = np_sum_f64(tmp, stop - start);
    free(tmp);
    return s;
}

/* get_observation (fixed_wing.py:1113-1262), general layout: obs_len rows (newest first) of obs_n entries, entry kinds
 * state / target absolute / target relative / action, history rows clamped to the episode start with the
 * `init_noise` offset, optional (val - mean) / var normalisation. */
static void get_observation_generic(const FwoEnv* e, double* obs) {
    const FwConfig* c = &e->cfg;
    const int L = c->obs_len, n = c->obs_n;
    const double cur_state[8] = {e->h_roll, e->h_pitch, e->h_Va, e->h_omega[0], e->h_omega[1], e->h_omega[2],
                                 e->h_alpha, e->h_beta};
    const double actval[3] = {e->elev, e->ail, e->act_val[2]};
    const int W = c->integration_window, ostep = c->obs_step > 0 ? c->obs_step : 1;
    for (int row = 0; row < L; ++row) {
        const int i = 1 + row * ostep;        /* range(1, length * step, step) (fixed_wing.py:1129-1138) */
        int ie = i;
        double init_noise = 0;
        if (i > e->steps_count) {
            ie = e->steps_count + 1;
            if (L > 1) {
                double u;
                if (!isnan(c->obs_init_noise)) u = c->obs_init_noise;
                else {
                    uint32_t r[4];
                    rng_block(e->ep_seed, e->env_id, e->episode, FWO_RNG_OBS_INIT, (uint32_t)(e->steps_count * 8 + (i - 1)), r);
                    u = 2.0 * u53(r[0], r[1]) - 1.0;
                }
                init_noise = u * c->dt;
            }
        }
        for (int k = 0; k < n; ++k) {
            double val;
            const int idx = c->obs_idx[k];
            switch (c->obs_kind[k]) {
                case FW_OBS_STATE:
                    val = (ie == 1) ? cur_state[idx] : e->st_hist[(size_t)(e->n_st - ie) * 8 + idx];
                    break;
                case FW_OBS_TARGET_ABS:
                    if (idx >= 3) val = (ie == 1) ? e->atarget[idx - 3] : e->atgt_hist[(size_t)(e->n_tgt - ie) * 3 + idx - 3];
                    else val = (ie == 1) ? e->target[idx] : e->tgt_hist[(size_t)(e->n_tgt - ie) * 3 + idx];
                    break;
                case FW_OBS_TARGET_REL:
                    if (idx >= 3) val = (ie == 1) ? get_error(e, idx) : e->aerr_hist[(size_t)(e->n_err - ie) * 3 + idx - 3];
                    else val = (ie == 1) ? get_error(e, idx) : e->err_hist[(size_t)(e->n_err - ie) * 3 + idx];
                    break;
                case FW_OBS_TARGET_INT:       /* fixed_wing.py:1165-1180 */
                    if (!e->has_history) val = get_error(e, idx) * W;
                    else {
                        const int n = e->obs_hist_n;
                        /* history[-W - i : -i]: a python slice (W + i > n clamps to the start) */
                        val = err_slice_sum(e, n, idx, n - W - ie < 0 ? 0 : n - W - ie, n - ie);
                        if (e->steps_count - ie < W)
                            val += (W - (e->steps_count - ie)) * (idx < 3 ? e->err_hist[idx] : e->aerr_hist[idx - 3]);
                    }
                    break;
                default: {
                    if (e->steps_count - ie < 0) {
                        val = actval[idx];
                        if (c->scale_actions)
                            val = (c->scale_high - c->scale_low) * (val - c->act_lo[idx]) / (c->act_hi[idx] - c->act_lo[idx]) + c->scale_low;
                    } else {
                        const double* hist = c->scale_actions ? e->act_hist : e->cmd_hist;
                        const int N = c->scale_actions ? e->n_act : e->n_cmd;
                        int hi = N - (ie - 1), lo = N - c->obs_window[k] - ie + 1;
                        if (lo < 0) lo = 0;
                        val = window_feature_f32(hist, lo, hi, idx, c->scale_actions ? e->act_is_f32 : 0);
                    }
                }
            }
            val += init_noise;
            if (c->obs_normalize && c->obs_norm_flag[k]) { val -= c->obs_mean[k]; val /= c->obs_var[k]; }
            obs[row * n + k] = val;
        }
    }
    add_obs_noise(e, obs, L * n);
}

/* get_observation (fixed_wing.py:1113-1262) for observation.length == 1, no normalisation */
static void get_observation(const FwoEnv* e, double* obs) {
    const FwConfig* c = &e->cfg;
    if (c->obs_generic) { get_observation_generic(e, obs); return; }
    obs[0] = e->h_roll; obs[1] = e->h_pitch; obs[2] = e->h_Va;
    obs[3] = e->h_omega[0]; obs[4] = e->h_omega[1]; obs[5] = e->h_omega[2];
    obs[6] = e->target[0]; obs[7] = e->target[1]; obs[8] = e->target[2];
    obs[9] = e->h_alpha; obs[10] = e->h_beta;
    double actval[3] = {e->elev, e->ail, e->act_val[2]};
    for (int j = 0; j < 3; ++j) {
        if (e->steps_count - 1 < 0) {
            double v = actval[j];
            if (c->scale_actions)   /* linear_action_scaling(direction="backward") (fixed_wing.py:643-652) */
                v = (c->scale_high - c->scale_low) * (v - c->act_lo[j]) / (c->act_hi[j] - c->act_lo[j]) + c->scale_low;
            obs[11 + j] = v;
        } else if (c->scale_actions) {
            obs[11 + j] = delta_feature_f32(e->act_hist, e->n_act, j, c->obs_act_window, e->act_is_f32);
        } else {
            obs[11 + j] = delta_feature_f32(e->cmd_hist, e->n_cmd, j, c->obs_act_window, 0);
        }
    }
    add_obs_noise(e, obs, FW_NOBS);
}

/* get_reward (fixed_wing.py:941-1111) for the default factor family: three linear error factors, a linear
 * action-delta factor and a linear action-bound factor, one "linear" term of weight 1, form "absolute". */
static double get_reward(const FwoEnv* e, const double action[3]) {
    const FwConfig* c = &e->cfg;
    double val_shaping = 0, val_plain = 0;
    for (int k = 0; k < 3; ++k) {
        if (c->rew_err_scaling[k] <= 0) continue;
        double v = fabs(get_error(e, k)) / c->rew_err_scaling[k];
        v = clipd(v, 0, c->rew_err_max[k]);
        val_shaping += v * -1.0;
    }
    if (c->rew_delta_scaling > 0) {
        double v = 0;
        if (e->steps_count > 1) {
            int w = c->rew_delta_window, lo = e->n_act - w < 0 ? 0 : e->n_act - w, rows = e->n_act - lo - 1;
            if (e->act_is_f32) {
                float d[3 * FW_ACT_WINDOW_MAX];
                for (int t = 0; t < rows; ++t)
                    for (int j = 0; j < 3; ++j)
                        d[t * 3 + j] = fabsf((float)e->act_hist[(lo + t + 1) * 3 + j] - (float)e->act_hist[(lo + t) * 3 + j]);
                float s = np_sum_f32(d, rows * 3);
                v = (double)fminf(fmaxf(s / (float)c->rew_delta_scaling, 0.0f), (float)c->rew_delta_max);
            } else {
                double d[3 * FW_ACT_WINDOW_MAX];
                for (int t = 0; t < rows; ++t)
                    for (int j = 0; j < 3; ++j)
                        d[t * 3 + j] = fabs(e->act_hist[(lo + t + 1) * 3 + j] - e->act_hist[(lo + t) * 3 + j]);
                v = clipd(np_sum_f64(d, rows * 3) / c->rew_delta_scaling, 0, c->rew_delta_max);
            }
        }
        val_plain += v * -1.0;
    }
    if (c->rew_bound_scaling > 0 && c->has_action_bounds) {
        double hi = 0, lo = 0;
        for (int j = 0; j < 3; ++j) {
            if (action[j] > c->action_bounds_max[j]) hi += fabs(action[j] - c->action_bounds_max[j]);
            if (action[j] < c->action_bounds_min[j]) lo += fabs(action[j] - c->action_bounds_min[j]);
        }
        double v = clipd(fabs(hi + lo) / c->rew_bound_scaling, 0, c->rew_bound_max);
        val_plain += v * -1.0;
    }
    return 1.0 * (val_plain + val_shaping);
}

/* get_reward (fixed_wing.py:941-1111), the general engine: any list of factors (state error / value, action value /
 * delta / bound, success, step, goal per_state / all) with linear / exponential / quadratic function classes, shaping
 * and non-shaping parts per term, absolute or potential form. */
static double get_reward_generic(FwoEnv* e, const double action[3], int success) {
    const FwConfig* c = &e->cfg;
    double val_t[3] = {0, 0, 0}, shp_t[3] = {0, 0, 0};
    const double states[8] = {e->roll, e->pitch, e->Va, e->omega[0], e->omega[1], e->omega[2], e->alpha, e->beta};
    uint8_t g[4];
    goal_status(e, g);
    for (int i = 0; i < c->rew_n; ++i) {
        double val = 0;
        switch (c->rew_class[i]) {
            case FW_RF_STATE_ERROR: val = get_error(e, c->rew_idx[i]); break;
            case FW_RF_STATE_VALUE: val = states[c->rew_idx[i]]; break;
            case FW_RF_STATE_INT_ERROR: {     /* fixed_wing.py:1003-1012; [-0:] is the whole list */
                const int W = c->integration_window, n = e->n_err;
                val = err_slice_sum(e, n, c->rew_idx[i], (W == 0 || W > n) ? 0 : n - W, n);
                if (e->steps_count < W)
                    val += (W - e->steps_count) * (c->rew_idx[i] < 3 ? e->err_hist[c->rew_idx[i]] : e->aerr_hist[c->rew_idx[i] - 3]);
                break;
            }
            case FW_RF_ACTION_VALUE:
                if (e->act_is_f32) { float sacc = 0; for (int j = 0; j < 3; ++j) sacc += fabsf((float)action[j]); val = sacc; }
                else for (int j = 0; j < 3; ++j) val += fabs(action[j]);
                break;
            case FW_RF_ACTION_DELTA:
                if (e->steps_count > 1) {
                    int w = c->rew_window[i], lo = e->n_act - w < 0 ? 0 : e->n_act - w, rows = e->n_act - lo - 1;
                    if (e->act_is_f32) {
                        float d[3 * FW_ACT_WINDOW_MAX];
                        for (int t = 0; t < rows; ++t) for (int j = 0; j < 3; ++j)
                            d[t * 3 + j] = fabsf((float)e->act_hist[(lo + t + 1) * 3 + j] - (float)e->act_hist[(lo + t) * 3 + j]);
                        val = (double)np_sum_f32(d, rows * 3);
                    } else {
                        double d[3 * FW_ACT_WINDOW_MAX];
                        for (int t = 0; t < rows; ++t) for (int j = 0; j < 3; ++j)
                            d[t * 3 + j] = fabs(e->act_hist[(lo + t + 1) * 3 + j] - e->act_hist[(lo + t) * 3 + j]);
                        val = np_sum_f64(d, rows * 3);
                    }
                }
                break;
            case FW_RF_ACTION_BOUND: {
                double hi = 0, lo = 0;
                for (int j = 0; j < 3; ++j) {
                    if (action[j] > c->action_bounds_max[j]) hi += fabs(action[j] - c->action_bounds_max[j]);
                    if (action[j] < c->action_bounds_min[j]) lo += fabs(action[j] - c->action_bounds_min[j]);
                }
                val = hi + lo;
                break;
            }
            case FW_RF_SUCCESS:
                val = success ? (c->rew_value_timesteps[i] ? (double)(c->steps_max - e->steps_count) : c->rew_value[i]) : 0;
                break;
            case FW_RF_STEP: val = c->rew_value[i]; break;
            case FW_RF_GOAL_PER_STATE: {      /* value / len(self.target) per achieved state (fixed_wing.py:1038-1044) */
                const int nt = c->ang_on ? 6 : 3;
                for (int k = 0; k < 3; ++k) val += g[k] ? c->rew_value[i] / nt : 0;
                if (c->ang_on) { uint8_t ga[3]; goal_status_ang(e, ga); for (int a = 0; a < 3; ++a) val += ga[a] ? c->rew_value[i] / nt : 0; }
                break;
            }
            case FW_RF_GOAL_ALL: val += g[3] ? c->rew_value[i] : 0; break;
        }
        /* values derived from a float32 action array stay float32 through the function class (numpy keeps the array
         * dtype against python scalars); everything else is float64 */
        const int f32v = e->act_is_f32 && (c->rew_class[i] == FW_RF_ACTION_DELTA || c->rew_class[i] == FW_RF_ACTION_VALUE);
        if (c->rew_fclass[i] == FW_FN_LINEAR) {
            if (f32v) val = (double)fminf(fmaxf(fabsf((float)val) / (float)c->rew_scaling[i], 0.0f), (float)c->rew_maxv[i]);
            else val = clipd(fabs(val) / c->rew_scaling[i], 0, c->rew_maxv[i]);
        } else if (f32v) val = (double)(((float)val * (float)val) / (float)c->rew_scaling[i]);
        else val = val * val / c->rew_scaling[i];
        if (c->rew_shaping[i]) shp_t[c->rew_fclass[i]] += val * c->rew_sign[i];
        else val_t[c->rew_fclass[i]] += val * c->rew_sign[i];
    }
    double reward = 0;
    for (int t = 0; t < c->rew_nterms; ++t) {
        const int fc = c->term_fclass[t];
        double v;
        if (fc == FW_FN_EXPONENTIAL) {
            if (c->rew_potential) v = e->has_prev_shaping[fc] ? -1 + exp(val_t[fc] + (shp_t[fc] - e->prev_shaping[fc])) : -1 + exp(val_t[fc]);
            else v = -1 + exp(val_t[fc] + shp_t[fc]);
        } else {
            v = val_t[fc];
            if (c->rew_potential) { if (e->has_prev_shaping[fc]) v += shp_t[fc] - e->prev_shaping[fc]; }
            else v += shp_t[fc];
        }
        e->prev_shaping[fc] = shp_t[fc];
        e->has_prev_shaping[fc] = 1;
        reward += c->term_weight[t] * v;
    }
    return reward;
}

/* get_metric x9 (fixed_wing.py:1644-1736) computed from the full histories exactly as the reference does */
/* the five error metrics of ONE target state from its error history hist[t * 3 + col] (fixed_wing.py:1655-1730) */
static void state_error_metrics(const FwoEnv* e, const double* hist, int col, double out5[5]) {
    const FwConfig* c = &e->cfg;
    const int ne = e->n_err;
    double e0 = hist[col], sum = 0, sabs = 0, vmin = INFINITY, vmax = -INFINITY;
    for (int t = 0; t < ne; ++t) {
        double v = hist[t * 3 + col];
        sum += v; sabs += fabs(v); vmin = fmin(vmin, v); vmax = fmax(vmax, v);
    }
    out5[0] = (fabs(e0) >= 0.01) ? fabs((sum / ne) / e0) : NAN;
    out5[1] = sabs;
    int lo = ne - FW_END_ERR_WINDOW < 0 ? 0 : ne - FW_END_ERR_WINDOW;
    double s50 = 0;
    for (int t = lo; t < ne; ++t) s50 += hist[t * 3 + col];
    out5[2] = fabs(s50 / (ne - lo));
    /* rise_time: reverse scan (fixed_wing.py:1702-1719) */
    double rise_end = NAN, rise_start = NAN, low_lim = fabs(c->rise_low * e0), high_lim = fabs(c->rise_high * e0);
    for (int j = 1; j < ne; ++j) {
        double er = fabs(hist[(ne - 1 - j) * 3 + col]), prev = fabs(hist[(ne - j) * 3 + col]);
        if (er >= low_lim && prev < low_lim) rise_end = e->steps_count - j;
        if (er >= high_lim && prev < high_lim) rise_start = e->steps_count - j;
    }
    out5[3] = rise_end - rise_start;
    /* overshoot (fixed_wing.py:1722-1730) */
    double opp = (e0 > 0) ? vmin : vmax;
    out5[4] = (sgn(opp) == sgn(e0)) ? NAN : fabs(opp / e0);
}

/* success / settling_time (fixed_wing.py:1684-1699), success_time_frac (:1733-1734) of one goal column */
static void goal_metrics(const FwoEnv* e, const uint8_t* gh, int stride, int col, double out3[3]) {
    const FwConfig* c = &e->cfg;
    const int ng = e->n_goal;
    int cnt = 0, total = 0, settle = -1;
    for (int t = 0; t < ng; ++t) {
        cnt += gh[t * stride + col];
        total += gh[t * stride + col];
        if (t >= c->streak_req) cnt -= gh[(t - c->streak_req) * stride + col];
        if (settle < 0 && t + 1 >= c->streak_req && (double)cnt / c->streak_req >= c->streak_fraction) settle = t;
    }
    out3[0] = settle >= 0;
    out3[1] = settle >= 0 ? (double)settle : NAN;
    out3[2] = (double)total / ng;
}

static void compute_metrics(FwoEnv* e) {
    const FwConfig* c = &e->cfg;
    double* m = e->metrics;
    for (int k = 0; k < 3; ++k) {
        double o[5];
        state_error_metrics(e, e->err_hist, k, o);
        m[FW_M_AVG_ERROR + k] = o[0]; m[FW_M_TOTAL_ERROR + k] = o[1]; m[FW_M_END_ERROR + k] = o[2];
        m[FW_M_RISE_TIME + k] = o[3]; m[FW_M_OVERSHOOT + k] = o[4];
    }
    /* control_variation (fixed_wing.py:1670-1680) */
    {
        double s = 0;
        for (int t = 1; t < e->n_cmd; ++t)
            for (int j = 0; j < 3; ++j) s += fabs(e->cmd_hist[t * 3 + j] - e->cmd_hist[(t - 1) * 3 + j]);
        m[FW_M_CONTROL_VARIATION] = s / (3 * c->dt * (e->n_cmd - 1));
    }
    for (int k = 0; k < 4; ++k) {
        double o[3];
        goal_metrics(e, e->goal_hist, 4, k, o);
        m[FW_M_SUCCESS + k] = o[0]; m[FW_M_SETTLING_TIME + k] = o[1]; m[FW_M_SUCCESS_TIME_FRAC + k] = o[2];
    }
    /* attitude_angular: the same metrics for omega_p/q/r, metric-major (FW_NMETRIC_ANG) */
    if (c->ang_on) for (int a = 0; a < 3; ++a) {
        double o[5], g3[3];
        state_error_metrics(e, e->aerr_hist, a, o);
        goal_metrics(e, e->agoal_hist, 3, a, g3);
        for (int q = 0; q < 5; ++q) e->ametrics[q * 3 + a] = o[q];
        for (int q = 0; q < 3; ++q) e->ametrics[15 + q * 3 + a] = g3[q];
    }
}

/* ------------------------------------------------------------------------------------------------------------- */
/* public oracle API (ctypes) */

FwoEnv* fwo_create(const FwConfig* cfg, int64_t env_id) {
    FwoEnv* e = (FwoEnv*)calloc(1, sizeof(FwoEnv));
    e->cfg = *cfg;
    gammas(e);
    e->Jy0 = e->cfg.Jy;
    int L = cfg->steps_max + 2;
    e->turb_len = cfg->steps_max > 0 ? cfg->steps_max : 1;
    e->turb = (double*)calloc((size_t)6 * e->turb_len, sizeof(double));
    e->act_hist = (double*)calloc((size_t)3 * L, sizeof(double));
    e->cmd_hist = (double*)calloc((size_t)3 * L, sizeof(double));
    e->err_hist = (double*)calloc((size_t)3 * L, sizeof(double));
    e->goal_hist = (uint8_t*)calloc((size_t)4 * L, 1);
    e->st_hist = (double*)calloc((size_t)8 * L, sizeof(double));
    e->tgt_hist = (double*)calloc((size_t)3 * L, sizeof(double));
    e->aerr_hist = (double*)calloc((size_t)3 * L, sizeof(double));
    e->atgt_hist = (double*)calloc((size_t)3 * L, sizeof(double));
    e->agoal_hist = (uint8_t*)calloc((size_t)3 * L, 1);
    e->env_id = env_id;
    e->episode = 0;
    return e;
}

void fwo_destroy(FwoEnv* e) {
    if (!e) return;
    free(e->turb); free(e->act_hist); free(e->cmd_hist); free(e->err_hist); free(e->goal_hist);
    free(e->st_hist); free(e->tgt_hist); free(e->aerr_hist); free(e->atgt_hist); free(e->agoal_hist); free(e);
}

/* sample_simulator_parameters, the "model" block (fixed_wing.py:758-800): every enabled aircraft parameter whose
 * original value is not 0 is re-drawn around the original at every reset.  The parameters are the 48 doubles of
 * FwConfig from `mass` on (this env's private copy of the config is what dynamics() reads; pyfly's _forces reads
 * self.params on every call).  Draw i of the episode: Philox block i of purpose MODEL — two uniforms, Box-Muller cosine
 * branch for the gaussian, the first uniform for the uniform distribution. */
static void sample_model_params(FwoEnv* e) {
    FwConfig* c = &e->cfg;
    if (!c->model_on) return;
    double* par = &c->mass;
    for (int i = 0; i < FW_NPARAM; ++i) {
        if (!c->par_enabled[i]) continue;
        const double orig = c->par_orig[i];
        if (orig == 0) continue;
        uint32_t r[4];
        rng_block(e->ep_seed, e->env_id, e->episode, FWO_RNG_MODEL, (uint32_t)i, r);
        const double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
        double v;
        if (c->model_uniform) {
            const double lo = orig - c->par_var[i], hi = orig + c->par_var[i];
            v = lo + (hi - lo) * u1;
        } else {
            const double z = sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586476925 * u2);
            v = orig + c->par_var[i] * z;
            if (!isnan(c->par_clip[i])) {       /* np.clip(v, orig - clip, orig + clip) = minimum(maximum(v, lo), hi) */
                const double lo = orig - c->par_clip[i], hi = orig + c->par_clip[i];
                v = fmin(fmax(v, lo), hi);
            }
        }
        par[i] = v;
    }
}
/* the aircraft parameters of the running episode (tests: inject the reference's draws / read ours back) */
void fwo_set_params(FwoEnv* e, const double* par48) { memcpy(&e->cfg.mass, par48, sizeof(double) * FW_NPARAM); }
void fwo_get_params(const FwoEnv* e, double* par48) { memcpy(par48, &e->cfg.mass, sizeof(double) * FW_NPARAM); }

/* turbulence for the whole episode (pyfly.py:870-871 -> dryden.simulate) */
static void gen_turbulence(FwoEnv* e, const double* noise, int noise_len) {
    const FwConfig* c = &e->cfg;
    if (!c->turbulence) return;
    int L = e->turb_len;
    double* nz = (double*)malloc(sizeof(double) * 4 * L);
    for (int k = 0; k < L; ++k) {
        double n4[4];
        if (noise) for (int r = 0; r < 4; ++r) n4[r] = noise[(size_t)r * noise_len + (k < noise_len ? k : noise_len - 1)];
        else fwo_noise4(e->ep_seed, e->env_id, e->episode, (uint32_t)k, n4);
        for (int r = 0; r < 4; ++r) nz[(size_t)r * L + k] = n4[r];
    }
    fwo_dryden(c, nz, L, e->turb);
    free(nz);
}

/* ---------------- waypoint head: FixedWingAircraft_simple (magpie/magpy/simple_train.py:197-702) ---------------- */

/* simulator.reset(state=waypoint) (simple_train.py:357-363 -> pyfly.py:1262-1311): position, attitude, velocity and wind
 * from the waypoint; omega from the waypoint row or, if NaN, uniform in the pyfly init range; actuators at their (0, 0)
 * init range; turbulence restarted with a fresh noise stream. */
static void wp_sim_reset(FwoEnv* e, const double* row) {
    const FwConfig* c = &e->cfg;
    e->episode += 1;                     /* a fresh Philox segment for omega and the turbulence noise */
    e->ep_seed = e->cfg.seed;
    e->cur_sim_step = 0;
    for (int i = 0; i < 3; ++i) { e->pos[i] = row[i]; e->vel[i] = row[6 + i]; e->wind[i] = row[9 + i]; }
    e->roll = row[3]; e->pitch = row[4]; e->yaw = row[5];
    for (int i = 0; i < 3; ++i)
        e->omega[i] = isnan(row[12 + i]) ? c->init_lo[3 + i] + (c->init_hi[3 + i] - c->init_lo[3 + i]) * rng_uniform(e, FWO_RNG_RESET, 3 + i)
                                         : row[12 + i];
    double a6[6] = {0, 0, 0, 0, 0, 0};
    set_actuators(e, a6);
    e->elev = 0; e->ail = 0;
    gen_turbulence(e, NULL, 0);
    double R[3][3];
    rot_euler(e->roll, e->pitch, e->yaw, R);
    airspeed(e, R, e->vel, &e->Va, &e->alpha, &e->beta);
    if (e->Va < c->va_value_min) e->Va = c->va_value_min;
    double phi = e->roll, theta = e->pitch, psi = e->yaw;
    double cps = cos(psi / 2), sps = sin(psi / 2), cth = cos(theta / 2), sth = sin(theta / 2);
    double cph = cos(phi / 2), sph = sin(phi / 2);
    e->quat[0] = cps * cth * cph + sps * sth * sph;
    e->quat[1] = cps * cth * sph - sps * sth * cph;
    e->quat[2] = cps * sth * cph + sps * cth * sph;
    e->quat[3] = sps * cth * cph - cps * sth * sph;
    e->n_cmd = 0;
}

static void wp_set_leg(FwoEnv* e) {      /* sample_target(id, pos): start = tasks[id][pos], goal = tasks[id][pos+1] */
    const double* start = e->wp_tasks + ((size_t)e->wp_task * e->wp_len + e->wp_pos) * FW_WP_ROW;
    wp_sim_reset(e, start);
    for (int k = 0; k < 3; ++k) e->wp_goal[k] = start[FW_WP_ROW + k];
}

static void wp_observation(const FwoEnv* e, double* obs) {   /* simple_train.py:692-702: the states' .value */
    obs[0] = e->roll; obs[1] = e->pitch; obs[2] = e->Va;
    obs[3] = e->omega[0]; obs[4] = e->omega[1]; obs[5] = e->omega[2];
    obs[6] = e->act_val[1]; obs[7] = e->act_val[0]; obs[8] = e->act_val[2];      /* elevon_left, elevon_right, throttle */
    obs[9] = e->pos[0]; obs[10] = e->pos[1]; obs[11] = e->pos[2];
}

void fwo_set_waypoint_tasks(FwoEnv* e, const double* tasks, int n_tasks, int wp_len, int task) {
    e->wp_tasks = tasks; e->wp_n_tasks = n_tasks; e->wp_len = wp_len; e->wp_task = task; e->wp_pos = 0;
}

static void wp_reset(FwoEnv* e, double* obs) {               /* reset (simple_train.py:385-408): reset_task(idx) */
    e->steps_count = 0;
    e->wp_pos = 0;
    e->ep_return = 0; e->term_code = 0;
    wp_set_leg(e);
    wp_observation(e, obs);
}

static void wp_step(FwoEnv* e, const double action[3], double* obs, double* reward, int* done, int* term) {
    const FwConfig* c = &e->cfg;
    int fail = sim_step(e, action);                          /* commands are passed straight through (:441-444) */
    e->steps_count += 1;
    int d = 0, tc = FW_TERM_NONE;
    double rew;
    if (c->steps_max > 0 && e->steps_count >= c->steps_max) { d = 1; tc = FW_TERM_STEPS; }
    if (!fail) {
        int all = 1;
        for (int k = 0; k < 3; ++k) all &= fabs(e->wp_goal[k] - e->pos[k]) <= c->wp_goal_bound[k];
        if (all) {                                           /* sample_task(idx) (simple_train.py:346-355, 487-491) */
            if (e->wp_pos < e->wp_len - 2) e->wp_pos += 1; else e->wp_pos = 0;
            wp_set_leg(e);
        }
        double s = 0;                                        /* get_reward (simple_train.py:673-690), after the teleport */
        for (int k = 0; k < 3; ++k) s += 1.0 * (fabs(e->wp_goal[k] - e->pos[k]) / c->wp_rew_range[k]);
        rew = 1 / exp(s);
    } else {
        d = 1;
        rew = (double)(e->steps_count - c->steps_max);
        tc = fail;
    }
    wp_observation(e, obs);
    e->ep_return += rew;
    if (d) { e->term_code = tc; for (int k = 0; k < FW_NMETRIC; ++k) e->metrics[k] = NAN; }
    *reward = rew; *done = d; *term = tc;
}

/* FixedWingAircraft.reset (fixed_wing.py:414-481) -> PyFly.reset (pyfly.py:1262-1311).
 * state: FW_NSTATE_INJECT doubles or NULL, NaN entries are sampled; target: 3 doubles or NULL;
 * noise: [4][noise_len] unit normals or NULL (Philox). */
void fwo_reset(FwoEnv* e, const double* state, const double* target, const double* noise, int noise_len,
               double obs[FW_NOBS]) {
    const FwConfig* c = &e->cfg;
    if (c->env_kind == FW_ENV_WAYPOINT) { wp_reset(e, obs); return; }
    e->episode += 1;
    e->ep_seed = c->seed;
    e->steps_count = 0;
    e->cur_sim_step = 0;
    double s12[12];
    for (int i = 0; i < 12; ++i) {
        if (state && !isnan(state[i])) s12[i] = state[i];
        else s12[i] = c->init_lo[i] + (c->init_hi[i] - c->init_lo[i]) * rng_uniform(e, FWO_RNG_RESET, i);
    }
    e->roll = s12[0]; e->pitch = s12[1]; e->yaw = s12[2];
    for (int i = 0; i < 3; ++i) { e->omega[i] = s12[3 + i]; e->pos[i] = s12[6 + i]; e->vel[i] = s12[9 + i]; }
    double a6[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; ++i) if (state && !isnan(state[12 + i])) a6[i] = state[12 + i];
    set_actuators(e, a6);    /* ControlVariable.reset -> apply_conditions (pyfly.py:340-366) */
    /* elevator / aileron are `disabled` ControlVariables: their reset() forces value 0 (pyfly.py:359-363) whatever the
     * elevons are; they are only re-derived by the first state write-back (t > 0).  The first RHS call of the first
     * step (t == 0) therefore sees elevator = aileron = 0. */
    e->elev = 0; e->ail = 0;
    /* Wind.reset (pyfly.py:799-830) */
    if (state && !isnan(state[18]) && !isnan(state[19]) && !isnan(state[20])) {
        for (int i = 0; i < 3; ++i) e->wind[i] = state[18 + i];
    } else {
        double mag = c->wind_mag_min + (c->wind_mag_max - c->wind_mag_min) * rng_uniform(e, FWO_RNG_RESET, 12);
        double w_n = -mag + (mag - -mag) * rng_uniform(e, FWO_RNG_RESET, 13);
        double w_e_max = sqrt(mag * mag - w_n * w_n);
        double w_e = -w_e_max + (w_e_max - -w_e_max) * rng_uniform(e, FWO_RNG_RESET, 14);
        double w_d = sqrt(mag * mag - w_n * w_n - w_e * w_e);
        e->wind[0] = w_n; e->wind[1] = w_e; e->wind[2] = w_d;
    }
    gen_turbulence(e, noise, noise_len);
    /* Va, alpha, beta from Euler + vel (pyfly.py:1297-1306) */
    double R[3][3];
    rot_euler(e->roll, e->pitch, e->yaw, R);
    airspeed(e, R, e->vel, &e->Va, &e->alpha, &e->beta);
    if (e->Va < c->va_value_min) e->Va = c->va_value_min;
    /* AttitudeQuaternion._from_euler_angles (pyfly.py:714-737) */
    {
        double phi = e->roll, theta = e->pitch, psi = e->yaw;
        double cps = cos(psi / 2), sps = sin(psi / 2), cth = cos(theta / 2), sth = sin(theta / 2);
        double cph = cos(phi / 2), sph = sin(phi / 2);
        e->quat[0] = cps * cth * cph + sps * sth * sph;
        e->quat[1] = cps * cth * sph - sps * sth * cph;
        e->quat[2] = cps * sth * cph + sps * cth * sph;
        e->quat[3] = sps * cth * cph - cps * sth * sph;
    }
    e->h_roll = e->roll; e->h_pitch = e->pitch; e->h_Va = e->Va; e->h_alpha = e->alpha; e->h_beta = e->beta;
    for (int i = 0; i < 3; ++i) e->h_omega[i] = e->omega[i];
    e->obs_hist_n = e->n_err;       /* the reset observation still sees the ended episode's error history */
    e->n_act = 0; e->n_cmd = 0; e->n_err = 0; e->n_goal = 0; e->act_is_f32 = 0;
    e->n_st = 0; e->n_tgt = 0;
    for (int k = 0; k < 3; ++k) { e->prev_shaping[k] = 0; e->has_prev_shaping[k] = 0; }
    push_state_history(e);          /* Variable.reset: history = [value] (pyfly.py:89-104) */
    e->ep_return = 0; e->term_code = 0;
    sample_model_params(e);         /* fixed_wing.py:437: after simulator.reset, before sample_target */
    /* sample_target, then injected targets override (fixed_wing.py:443-450) */
    double u12[12];
    target_draws(e, FWO_RNG_RESET, 8, u12);          /* blocks 8..13 of the reset stream (0..7: state and wind) */
    sample_target(e, u12);
    if (target) for (int k = 0; k < 3; ++k) if (!isnan(target[k])) {
        if (e->tcls[k] != FW_TGT_CONSTANT && e->tcls[k] != FW_TGT_COMPENSATE) e->tcls[k] = FW_TGT_CONSTANT;
        e->target[k] = target[k];
    }
    get_observation(e, obs);
    for (int k = 0; k < 3; ++k) { e->err_hist[k] = get_error(e, k); e->tgt_hist[k] = e->target[k]; }
    if (c->ang_on) for (int a = 0; a < 3; ++a) { e->aerr_hist[a] = get_error(e, 3 + a); e->atgt_hist[a] = e->atarget[a]; }
    e->n_err = 1; e->n_tgt = 1;
    e->has_history = 1; e->obs_hist_n = 1;
    goal_status(e, e->goal_hist);
    goal_status_ang(e, e->agoal_hist);
    e->n_goal = 1;
}

/* FixedWingAircraft.step (fixed_wing.py:483-628). action: raw agent action; action_is_f32: the caller's action
 * array was float32 (SB3 policies) — selects numpy's float32 arithmetic for the few ops that inherit the dtype. */
void fwo_step(FwoEnv* e, const double action[3], int action_is_f32, double obs[FW_NOBS], double* reward,
              int* done, int* term) {
    const FwConfig* c = &e->cfg;
    if (c->env_kind == FW_ENV_WAYPOINT) { wp_step(e, action, obs, reward, done, term); return; }
    e->act_is_f32 = action_is_f32;
    if (e->n_act < c->steps_max + 1) { for (int j = 0; j < 3; ++j) e->act_hist[e->n_act * 3 + j] = action[j]; e->n_act++; }
    double a[3] = {action[0], action[1], action[2]};
    if (c->scale_actions) {
        for (int j = 0; j < 3; ++j) {
            double cl = clipd(a[j], c->scale_low, c->scale_high);
            double num = action_is_f32 ? (double)((float)cl - (float)c->scale_low) : (cl - c->scale_low);
            a[j] = (c->act_hi[j] - c->act_lo[j]) * num / (c->scale_high - c->scale_low) + c->act_lo[j];
        }
    }
    int fail = sim_step(e, a);
    if (!fail) push_state_history(e);
    e->steps_count += 1;
    e->steps_for_target += 1;
    int d = 0, tc = FW_TERM_NONE;
    double rew;
    if (c->steps_max > 0 && e->steps_count >= c->steps_max) { d = 1; tc = FW_TERM_STEPS; }
    if (!fail) {
        int resample = 0, success_on_step = 0;
        if (c->streak_req > 0) {
            goal_status(e, e->goal_hist + (size_t)e->n_goal * 4);
            goal_status_ang(e, e->agoal_hist + (size_t)e->n_goal * 3);
            e->n_goal++;
            if (e->steps_for_target >= c->streak_req) {
                int cnt = 0;
                for (int t = e->n_goal - c->streak_req; t < e->n_goal; ++t) cnt += e->goal_hist[t * 4 + 3];
                if ((double)cnt / c->streak_req >= c->streak_fraction) {
                    success_on_step = !e->goal_achieved;     /* fixed_wing.py:546-547 */
                    e->goal_achieved = 1;
                    if (c->on_success == FW_SUCCESS_DONE) { d = 1; tc = FW_TERM_SUCCESS; }
                    else if (c->on_success == FW_SUCCESS_NEW) resample = 1;
                }
            }
        }
        rew = c->rew_generic ? get_reward_generic(e, action, success_on_step) : get_reward(e, action);
        if (resample || (c->resample_every > 0 && e->steps_for_target >= c->resample_every)) {
            double u12[12];
            target_draws(e, FWO_RNG_RESAMPLE, (uint32_t)(e->steps_count * 8), u12);
            sample_target(e, u12);
        }
        next_target(e);
        for (int k = 0; k < 3; ++k) {
            e->err_hist[(size_t)e->n_err * 3 + k] = get_error(e, k);
            e->tgt_hist[(size_t)e->n_tgt * 3 + k] = e->target[k];
            if (c->ang_on) {
                e->aerr_hist[(size_t)e->n_err * 3 + k] = get_error(e, 3 + k);
                e->atgt_hist[(size_t)e->n_tgt * 3 + k] = e->atarget[k];
            }
        }
        e->n_err++; e->n_tgt++;
        e->obs_hist_n = e->n_err;
        get_observation(e, obs);
    } else {
        d = 1;
        rew = c->step_fail_timesteps ? (double)(e->steps_count - c->steps_max) : c->step_fail_value;
        tc = fail;
        get_observation(e, obs);
    }
    e->ep_return += rew;
    if (d) { e->term_code = tc; compute_metrics(e); }
    *reward = rew; *done = d; *term = tc;
}

/* snapshot for tests: y[19], euler[3], vab[3], cmd[3], target[3], wind[3], counters */
void fwo_get(const FwoEnv* e, double y[NY], double euler[3], double vab[3], double cmd[3], double target[3],
             double wind[3], int32_t counters[6]) {
    for (int i = 0; i < 4; ++i) y[i] = e->quat[i];
    for (int i = 0; i < 3; ++i) { y[4 + i] = e->omega[i]; y[7 + i] = e->pos[i]; y[10 + i] = e->vel[i]; }
    for (int i = 0; i < 3; ++i) { y[13 + i] = e->act_val[i]; y[16 + i] = e->act_dot[i]; }
    euler[0] = e->roll; euler[1] = e->pitch; euler[2] = e->yaw;
    vab[0] = e->Va; vab[1] = e->alpha; vab[2] = e->beta;
    for (int i = 0; i < 3; ++i) { cmd[i] = e->cmd_in[i]; target[i] = e->target[i]; wind[i] = e->wind[i]; }
    counters[0] = e->steps_count; counters[1] = e->steps_for_target; counters[2] = e->cur_sim_step;
    counters[3] = (int32_t)e->episode; counters[4] = e->last_nfev; counters[5] = e->last_natt;
}

void fwo_get_angular(const FwoEnv* e, double atarget[3], double ametrics[FW_NMETRIC_ANG]) {
    for (int a = 0; a < 3; ++a) atarget[a] = e->atarget[a];
    for (int q = 0; q < FW_NMETRIC_ANG; ++q) ametrics[q] = e->ametrics[q];
}
void fwo_get_metrics(const FwoEnv* e, double metrics[FW_NMETRIC], double* ep_return, int32_t* ep_len, int32_t* term) {
    memcpy(metrics, e->metrics, sizeof(e->metrics));
    *ep_return = e->ep_return; *ep_len = e->steps_count; *term = e->term_code;
}

const double* fwo_turbulence(const FwoEnv* e, int* len) { *len = e->turb_len; return e->turb; }

/* one bare RHS evaluation at injected (y, cmd_dyn, wind, turbulence column): returns FwTermCode */
int fwo_rhs(const FwConfig* cfg, const double y[NY], const double cmd_dyn[3], const double wind[3],
            const double turb6[6], double dy[NY]) {
    FwoEnv* e = fwo_create(cfg, 0);
    for (int i = 0; i < 3; ++i) { e->cmd_dyn[i] = cmd_dyn[i]; e->wind[i] = wind[i]; }
    for (int i = 0; i < 6; ++i) e->turb[(size_t)i * e->turb_len] = turb6[i];
    int rc = dynamics(e, 1, y, dy);
    fwo_destroy(e);
    return rc;
}

/* RolloutBuffer.compute_returns_and_advantage (buffers.py:304-333) with numpy's dtype promotion restated:
 * rewards/values/dones are float32 arrays, gamma/gae_lambda python floats (weak scalars -> float32 arithmetic),
 * `dones` at the last step is a bool array so `1.0 - dones` is float64, which makes last_gae_lam float64 for the
 * whole reverse scan; it is rounded to float32 only when stored. */
void fwo_gae(const float* rew, const float* val, const float* done, const float* last_val, const uint8_t* last_done,
             float* adv, float* ret, int T, int N, double gamma, double lam) {
    float gf = (float)gamma, glf = (float)(gamma * lam);
    for (int n = 0; n < N; ++n) {
        double last_gae = 0;
        for (int t = T - 1; t >= 0; --t) {
            double delta, coef;
            if (t == T - 1) {
                double nnt = 1.0 - (double)last_done[n];
                float gv = gf * last_val[n];                                  /* float32 * float32 */
                delta = ((double)rew[t * N + n] + (double)gv * nnt) - (double)val[t * N + n];
                coef = (double)glf * nnt;
            } else {
                float nnt = 1.0f - done[(t + 1) * N + n];
                float d = rew[t * N + n] + gf * val[(t + 1) * N + n] * nnt - val[t * N + n];
                delta = (double)d;
                coef = (double)(glf * nnt);
            }
            last_gae = delta + coef * last_gae;
            adv[t * N + n] = (float)last_gae;
        }
    }
    for (int i = 0; i < T * N; ++i) ret[i] = adv[i] + val[i];
}

/* ------------------------------------------------------------------------------------------------------------- */
/* batch driver for the CPU baseline: n envs, auto-reset on done (VecEnv contract), OpenMP over envs */
int fwo_obs_dim(const FwConfig* c);
typedef struct FwoBatch { int n; int od; FwoEnv** envs; } FwoBatch;

FwoBatch* fwo_batch_create(const FwConfig* cfg, int n) {
    FwoBatch* b = (FwoBatch*)calloc(1, sizeof(FwoBatch));
    b->n = n;
    b->od = fwo_obs_dim(cfg);
    b->envs = (FwoEnv**)calloc((size_t)n, sizeof(FwoEnv*));
    for (int i = 0; i < n; ++i) b->envs[i] = fwo_create(cfg, cfg->env_id_offset + i);
    return b;
}
void fwo_batch_destroy(FwoBatch* b) {
    for (int i = 0; i < b->n; ++i) fwo_destroy(b->envs[i]);
    free(b->envs); free(b);
}
FwoEnv* fwo_batch_env(FwoBatch* b, int i) { return b->envs[i]; }

/* set_curriculum_level / seed on a live env (fixed_wing.py:324-412): the new ranges and seed serve every later reset
 * and target resampling; the running episode keeps its state, turbulence table and Philox key (ep_seed). */
void fwo_set_config(FwoEnv* e, const FwConfig* cfg) {
    double par[FW_NPARAM];
    memcpy(par, &e->cfg.mass, sizeof(par));      /* the running episode keeps its aircraft parameters */
    e->cfg = *cfg;
    if (cfg->model_on) memcpy(&e->cfg.mass, par, sizeof(par));
}
void fwo_batch_set_config(FwoBatch* b, const FwConfig* cfg) {
    for (int i = 0; i < b->n; ++i) fwo_set_config(b->envs[i], cfg);
}

void fwo_batch_reset(FwoBatch* b, double* obs) {
    for (int i = 0; i < b->n; ++i) fwo_reset(b->envs[i], NULL, NULL, NULL, 0, obs + (size_t)i * b->od);
}

/* actions [n][3] f32 (SB3 policy dtype); k_steps > 1 repeats the same batch call with Philox actions */
void fwo_batch_step(FwoBatch* b, const float* actions, double* obs, double* rew, uint8_t* done) {
    for (int i = 0; i < b->n; ++i) {
        double a[3] = {actions[i * 3], actions[i * 3 + 1], actions[i * 3 + 2]};
        int d, tc;
        fwo_step(b->envs[i], a, 1, obs + (size_t)i * b->od, rew + i, &d, &tc);
        done[i] = (uint8_t)d;
        if (d) fwo_reset(b->envs[i], NULL, NULL, NULL, 0, obs + (size_t)i * b->od);
    }
}

/* per-env RHS-evaluation / attempt counters of the last step, for step-size-decision parity statistics */
void fwo_batch_counters(const FwoBatch* b, int32_t* nfev, int32_t* natt) {
    for (int i = 0; i < b->n; ++i) { nfev[i] = b->envs[i]->last_nfev; natt[i] = b->envs[i]->last_natt; }
}

/* U(-1,1)^3 action for (env, global step) from the Philox action stream shared with fw_step_random */
void fwo_random_action(uint64_t action_seed, int64_t env_id, uint64_t step, float a[3]) {
    uint32_t r[4];
    rng_block(action_seed, env_id, step >> 32, FWO_RNG_ACTION, (uint32_t)step, r);
    for (int j = 0; j < 3; ++j) a[j] = (float)(((double)r[j] + 0.5) * (2.0 / 4294967296.0) - 1.0);
}

void fwo_batch_step_random(FwoBatch* b, int k_steps, uint64_t action_seed, uint64_t step0, double* obs, double* rew,
                           uint8_t* done) {
    for (int i = 0; i < b->n; ++i) {
        for (int k = 0; k < k_steps; ++k) {
            float af[3];
            fwo_random_action(action_seed, b->envs[i]->env_id, step0 + (uint64_t)k, af);
            double a[3] = {af[0], af[1], af[2]};
            int d, tc;
            fwo_step(b->envs[i], a, 1, obs + (size_t)i * b->od, rew + i, &d, &tc);
            done[i] = (uint8_t)d;
            if (d) fwo_reset(b->envs[i], NULL, NULL, NULL, 0, obs + (size_t)i * b->od);
        }
    }
}

int fwo_config_size(void) { return (int)sizeof(FwConfig); }
int fwo_obs_dim(const FwConfig* c) {
    if (c->env_kind == FW_ENV_WAYPOINT) return FW_NOBS_WAYPOINT;
    return c->obs_generic ? c->obs_len * c->obs_n : FW_NOBS;
}

void fwo_batch_set_waypoint_tasks(FwoBatch* b, const double* tasks, int n_tasks, int wp_len, const int32_t* task_of_env) {
    for (int i = 0; i < b->n; ++i) fwo_set_waypoint_tasks(b->envs[i], tasks, n_tasks, wp_len, task_of_env[i]);
}
