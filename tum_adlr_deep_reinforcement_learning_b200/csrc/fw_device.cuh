// fw_device.cuh — device code of the batched fixed-wing env step (sm_100a).
//
// One thread owns one env.  All persistent env state lives structure-of-arrays in HBM (`S.r[field * n + env]`,
// `S.i[field * n + env]`), so every load/store of a warp is one fully coalesced 128/256-byte transaction.
// The step is compute bound on the FP64 (exact mode) or FP32 (fast mode) CUDA-core pipe: ~24 RHS evaluations of
// ~900 FP instructions per env-step against ~1.6 KB of state traffic (DESIGN.md "roofline").
//
// Reference semantics followed (paths under magpie/libs/ of the reference; see DESIGN.md for the full map):
//   pyfly/pyfly/pyfly.py   PyFly.step :1358-1420, _dynamics :1450-1482, _forces :1484-1643, _f_*_dot :1645-1747,
//                          _rot_b_v :1749-1803, _calculate_airspeed_factors :1830-1850,
//                          _set_states_from_ode_solution :1852-1881, Actuation :453-655, reset :1262-1311
//   pyfly/pyfly/dryden.py  simulate :193-261 (streamed here: one lsim recurrence step per env-step)
//   fixed-wing-gym/gym_fixed_wing/fixed_wing.py   reset :414-481, step :483-628, sample_target :654-746,
//                          get_reward :941-1111, get_observation :1113-1262, _get_next_target :1363-1471,
//                          get_metric :1644-1736 (streamed: rings + running accumulators)
//   scipy.integrate.solve_ivp RK45 (un-vendored dependency; rk.py / common.py) for the exact integrator.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/fwb200.h"
#include "fw_math.cuh"

namespace fw {

// ---------------------------------------------------------------------------------------------------------------
// math wrappers: one spelling for float / double
template <typename T> struct M;
template <> struct M<double> {
    static __device__ __forceinline__ double sin(double x) { return ::sin(x); }
    static __device__ __forceinline__ double cos(double x) { return ::cos(x); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log(double x) { return ::log(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double rsqrt(double x) { return ::rsqrt(x); }
    static __device__ __forceinline__ double rsqrt_hot(double x) { return rsqrt_fast(x); }
    // straight-line versions for the RHS hot loop (fw_math.cuh)
    template <bool CT = true> static __device__ __forceinline__ double atan2_hot(double y, double x) { return atan2_bf<CT>(y, x); }
    template <bool CT = true> static __device__ __forceinline__ double asin_hot(double x) { return asin_bf<CT>(x); }
    template <bool CT = true> static __device__ __forceinline__ double exp_hot(double x) { return exp_bf<CT>(x); }
    static __device__ __forceinline__ double rcp_hot(double x) { return rcp_fast(x); }
    template <bool CT = true> static __device__ __forceinline__ double pow_hot(double x, double p) { return pow_hot_bf<CT>(x, p); }
    static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
    static __device__ __forceinline__ double asin(double x) { return ::asin(x); }
    static __device__ __forceinline__ double pow(double x, double y) { return ::pow(x, y); }
    static __device__ __forceinline__ double fabs(double x) { return ::fabs(x); }
    static __device__ __forceinline__ double fmin(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double fmax(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ double fmod(double a, double b) { return ::fmod(a, b); }
    static __device__ __forceinline__ double nan() { return CUDART_NAN; }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
    static __device__ __forceinline__ bool isnan(double x) { return ::isnan(x); }
};
template <> struct M<float> {
    static __device__ __forceinline__ float sin(float x) { return ::sinf(x); }
    static __device__ __forceinline__ float cos(float x) { return ::cosf(x); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { ::sincosf(x, s, c); }
    static __device__ __forceinline__ float exp(float x) { return ::expf(x); }
    static __device__ __forceinline__ float log(float x) { return ::logf(x); }
    static __device__ __forceinline__ float sqrt(float x) { return ::sqrtf(x); }
    static __device__ __forceinline__ float rsqrt(float x) { return ::rsqrtf(x); }
    static __device__ __forceinline__ float rsqrt_hot(float x) { return ::rsqrtf(x); }
    template <bool CT = true> static __device__ __forceinline__ float atan2_hot(float y, float x) { return ::atan2f(y, x); }
    // sin(beta) = a1 * rsqrt(|a|^2) can exceed 1 by a float32 rounding when the sideslip reaches 90 degrees: clamp, as
    // asin_bf does for its own argument, instead of handing NaN to the forces
    template <bool CT = true> static __device__ __forceinline__ float asin_hot(float x) { return ::asinf(fminf(fmaxf(x, -1.0f), 1.0f)); }
    template <bool CT = true> static __device__ __forceinline__ float exp_hot(float x) { return ::expf(x); }
    static __device__ __forceinline__ float rcp_hot(float x) { return 1.0f / x; }
    template <bool CT = true> static __device__ __forceinline__ float pow_hot(float x, float p) { return ::powf(x, p); }
    static __device__ __forceinline__ float atan2(float y, float x) { return ::atan2f(y, x); }
    static __device__ __forceinline__ float asin(float x) { return ::asinf(x); }
    static __device__ __forceinline__ float pow(float x, float y) { return ::powf(x, y); }
    static __device__ __forceinline__ float fabs(float x) { return ::fabsf(x); }
    static __device__ __forceinline__ float fmin(float a, float b) { return ::fminf(a, b); }
    static __device__ __forceinline__ float fmax(float a, float b) { return ::fmaxf(a, b); }
    static __device__ __forceinline__ float fmod(float a, float b) { return ::fmodf(a, b); }
    static __device__ __forceinline__ float nan() { return CUDART_NAN_F; }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
    static __device__ __forceinline__ bool isnan(float x) { return ::isnan(x); }
};

// ---------------------------------------------------------------------------------------------------------------
// device-side config: FwConfig converted once on the host to the arithmetic type of the kernel and passed to every
// launch as a __grid_constant__ kernel parameter (constant bank, warp-uniform broadcast reads).
template <typename T> struct DFilter {
    int order, noise_row;
    T Ad[9], Bd0[3], Bd1[3], C[3], D, Ablk[9];
};

template <typename T> struct DCfg {
    int integrator, rk4_substeps, turbulence, steps_max, scale_actions, has_action_bounds;
    int tgt_class[3], on_success, streak_req, resample_every, rew_delta_window, obs_act_window, step_fail_timesteps;
    T rtol, atol;
    T mass, Jy, S_wing, b, c, k_motor, k_T_P, k_Omega, M_, a_0, ar;
    T C_L_0, C_L_alpha, C_L_q, C_L_delta_e, C_D_p, C_D_q, C_D_beta1, C_D_beta2, C_D_delta_e;
    T C_m_0, C_m_alpha, C_m_q, C_m_delta_e, C_m_fp;
    T C_Y_0, C_Y_beta, C_Y_p, C_Y_r, C_Y_delta_a, C_Y_delta_r;
    T C_l_0, C_l_beta, C_l_p, C_l_r, C_l_delta_a, C_l_delta_r;
    T C_n_0, C_n_beta, C_n_p, C_n_r, C_n_delta_a, C_n_delta_r;
    T gam[9];
    T half_rho, mg, prop_k /* 0.5 rho S_prop C_prop */, inv_pi_e_ar, inv_Jy, inv_mass, exp_M_a0;
    T g_;        // gravity (per-env mg when aircraft parameters are randomised)
    int model_on;
    T dt, elevon_min, elevon_max, elevon_dot_max, w0sq, two_zeta_w0, inv_tau, throttle_min, throttle_max;
    T omega_con_min[3], omega_con_max[3], va_value_min, va_con_max;
    T init_lo[12], init_hi[12], wind_mag_min, wind_mag_max, turb_noise_scale;
    DFilter<T> filt[6];
    T scale_low, scale_high, act_lo[3], act_hi[3], action_bounds_min[3], action_bounds_max[3];
    T tgt_low[3], tgt_high[3], tgt_delta[3], tgt_bound[3], streak_fraction;
    int tgt_radians[3], tgt_moving;     // tgt_moving: some target class is linear / sinusoidal
    T tgt_slope_low[3], tgt_slope_high[3], tgt_amp_low[3], tgt_amp_high[3], tgt_period_low[3], tgt_period_high[3];
    T rng_u_override;
    int env_kind, turb_block_len;
    T wp_goal_bound[3], wp_rew_range[3];
    T rew_err_scaling[3], rew_err_max[3], rew_delta_scaling, rew_delta_max, rew_bound_scaling, rew_bound_max;
    T step_fail_value, rise_low, rise_high, obs_noise_mean, obs_noise_std;
    int rew_generic, rew_n, rew_potential, rew_nterms;
    int rew_class[FW_REW_FACTORS_MAX], rew_idx[FW_REW_FACTORS_MAX], rew_fclass[FW_REW_FACTORS_MAX];
    int rew_shaping[FW_REW_FACTORS_MAX], rew_window[FW_REW_FACTORS_MAX], rew_value_timesteps[FW_REW_FACTORS_MAX];
    T rew_scaling[FW_REW_FACTORS_MAX], rew_maxv[FW_REW_FACTORS_MAX], rew_sign[FW_REW_FACTORS_MAX], rew_value[FW_REW_FACTORS_MAX];
    int term_fclass[4];
    T term_weight[4];
    int obs_generic, obs_len, obs_n, obs_normalize;
    int obs_kind[FW_OBS_ENTRIES_MAX], obs_idx[FW_OBS_ENTRIES_MAX], obs_window[FW_OBS_ENTRIES_MAX], obs_norm_flag[FW_OBS_ENTRIES_MAX];
    T obs_mean[FW_OBS_ENTRIES_MAX], obs_var[FW_OBS_ENTRIES_MAX], obs_init_noise;
    int integration_window, obs_step, obs_has_int;
    int ang_on;                          // target class attitude_angular: omega_p/q/r are target states 3..5
    T ang_max_vel[3], ang_bound[3];
    unsigned long long seed;
    long long env_id_offset;
};

// Reset-time configuration: the ranges FixedWingAircraft.reset / sample_target draw from and the Philox key.  Unlike
// DCfg (a kernel parameter, frozen into every captured CUDA graph) it lives in DEVICE memory behind a stable pointer,
// so fw_set_config can change it on a live handle — set_curriculum_level (fixed_wing.py:334-412) and seed (:324-332)
// touch exactly these values and apply to FUTURE resets only; running episodes continue untouched.
template <typename T> struct ResetCfg {
    T init_lo[12], init_hi[12], wind_mag_min, wind_mag_max;
    T tgt_low[3], tgt_high[3], tgt_delta[3];
    T tgt_slope_low[3], tgt_slope_high[3], tgt_amp_low[3], tgt_amp_high[3], tgt_period_low[3], tgt_period_high[3];
    unsigned long long seed;
    // simulator.model (fixed_wing.py:758-800): which aircraft parameters are re-drawn at every reset, and how
    int model_uniform, par_enabled[FW_NPARAM];
    double par_orig[FW_NPARAM], par_var[FW_NPARAM], par_clip[FW_NPARAM];
};

// ---------------------------------------------------------------------------------------------------------------
// SoA field ids
enum RField {
    RF_Y = 0,                 // 19: quat4 omega3 pos3 vel3 act_val3 act_dot3
    RF_ROLL = 19, RF_PITCH, RF_VA, RF_ALPHA, RF_BETA,   // last committed (= .history[-1]) values
    RF_WIND = 24,             // 3
    RF_TGT = 27,              // 3
    RF_FX = 30,               // 12 filter states (filter f uses slots 3*? see FX_OFF)
    RF_FU = 42,               // 4 current (scaled) noise sample u_k
    RF_ACT_RING = 46,         // 4 x 3 previous raw actions, slot (age-1)*3 + j, age 1 = most recent
    RF_CMD_RING = 58,         // 4 x 3 previous constrained commands (same layout); slot 0..2 doubles as cmd_prev
    RF_CV_SUM = 70,           // sum |delta cmd|
    RF_E0 = 71,               // 3 initial errors
    RF_ESUM = 74, RF_EABS = 77, RF_EMIN = 80, RF_EMAX = 83, RF_EPREV = 86,   // 3 each
    RF_EP_RET = 89,
    // general observation layout only (obs_generic): 4 older rows of (8 states, 3 targets, 3 errors), and 8-deep rings
    // of raw actions / constrained commands (slot (age-1)*3 + j)
    RF_HIST = 90,             // 4 x 14
    RF_GACT = 146,            // 8 x 3
    RF_GCMD = 170,            // 8 x 3
    RF_PREV_SHAPING = 194,    // 3: prev_shaping per function class of the general reward engine (NaN = None)
    RF_TPROP = 197,           // 15: slope3 amplitude3 period3 phase3 bias3 of moving targets (tgt_moving only)
    // target class attitude_angular only (ang_on): the rate targets of omega_p/q/r and their error statistics
    RF_ATGT = 212,            // 3
    RF_AE0 = 215, RF_AESUM = 218, RF_AEABS = 221, RF_AEMIN = 224, RF_AEMAX = 227, RF_AEPREV = 230,   // 3 each
    RF_AHIST = 233,           // 4 older rows of (3 rate targets, 3 rate errors) for observation rows with a lag
    RF_COUNT = 257
};
enum IField {
    IF_STEPS = 0, IF_STEPS_TGT, IF_EPISODE, IF_SIM_STEP,
    IF_GOAL_RING = 4,         // 4 states x 4 words (128-bit rings)
    IF_GOAL_CNT = 20,         // 4 window counts
    IF_GOAL_TOTAL = 24,       // 4
    IF_SETTLE = 28,           // 4 (-1 = never)
    IF_RISE_LO = 32,          // 3 first index t with |e_t| >= low_lim and |e_{t+1}| < low_lim (-1 = none)
    IF_RISE_HI = 35,          // 3
    IF_NFEV = 38, IF_NATT = 39, IF_TERM = 40, IF_EP_LEN = 41, IF_ACT_F32 = 42,
    IF_WP_POS = 47,           // waypoint head: index of the current leg's start waypoint
    IF_TCLS = 44,             // 3: per-env target class (an injected target forces constant, fixed_wing.py:446-450)
    IF_GOAL_ACHIEVED = 43,    // self.goal_achieved: set by the first success, never cleared (fixed_wing.py:81, 546-547)
    IF_SEED_LO = 48, IF_SEED_HI = 49,   // Philox key of the running episode = ResetCfg.seed when the episode was reset
    // attitude_angular only: goal rings / counters / rise indices of omega_p/q/r (same meaning as the base ones)
    IF_AGOAL_RING = 50,       // 3 states x 4 words
    IF_AGOAL_CNT = 62, IF_AGOAL_TOTAL = 65, IF_ASETTLE = 68, IF_ARISE_LO = 71, IF_ARISE_HI = 74,   // 3 each
    IF_COUNT = 77
};

// Per-env aircraft parameters (model_on handles): S.par holds, per env, the FW_NPARAM base parameters in FwConfig order
// followed by the PD_COUNT values the RHS actually reads (base parameters and what convert_cfg derives from them).
#define FW_RHS_PARAMS(X) X(S_wing) X(b) X(c) X(k_motor) X(k_T_P) X(k_Omega) X(M_) X(a_0) X(C_L_0) X(C_L_alpha) X(C_L_q) \
    X(C_L_delta_e) X(C_D_p) X(C_D_q) X(C_D_beta1) X(C_D_beta2) X(C_D_delta_e) X(C_m_0) X(C_m_alpha) X(C_m_q) X(C_m_delta_e) \
    X(C_m_fp) X(C_Y_0) X(C_Y_beta) X(C_Y_p) X(C_Y_r) X(C_Y_delta_a) X(C_l_0) X(C_l_beta) X(C_l_p) X(C_l_r) X(C_l_delta_a) \
    X(C_n_0) X(C_n_beta) X(C_n_p) X(C_n_r) X(C_n_delta_a) X(mg) X(inv_mass) X(prop_k) X(inv_pi_e_ar) X(exp_M_a0)
enum RhsParam {
#define X(name) PD_##name,
    FW_RHS_PARAMS(X)
#undef X
    PD_COUNT
};
#define FW_PAR_FIELDS (FW_NPARAM + PD_COUNT)

template <typename T> struct Soa {
    T* r;              // [RF_COUNT][n]
    int32_t* i;        // [IF_COUNT][n]
    T* err_ring;       // [FW_END_ERR_WINDOW * 3][n]
    double* metrics;   // [n][FW_NMETRIC] (written on done)
    double* ep_ret;    // [n] episode return / length / term of the episode that ended at the last step
    int32_t* ep_len;
    int32_t* ep_term;
    const double* noise;   // injected unit white noise [n][4][noise_len] or nullptr (Philox)
    int noise_len;
    int n;
    const double* wp_tasks;      // waypoint head: [n_tasks][wp_len][FW_WP_ROW]
    const int32_t* wp_task_of_env;
    int wp_n_tasks, wp_len;
    const ResetCfg<T>* rc;       // device memory (fw_set_config)
    T* par;                      // [FW_PAR_FIELDS][n] per-env aircraft parameters, or nullptr (model off)
    T* err_ring_a;               // [FW_END_ERR_WINDOW * 3][n] error ring of omega_p/q/r, or nullptr (no attitude_angular targets)
    double* metrics_a;           // [n][FW_NMETRIC_ANG] (written on done), or nullptr
};

// the Philox key of env's running episode
template <typename T> __device__ __forceinline__ unsigned long long env_seed(const Soa<T>& S, int env) {
    return (unsigned long long)(uint32_t)S.i[IF_SEED_LO * S.n + env] |
           ((unsigned long long)(uint32_t)S.i[IF_SEED_HI * S.n + env] << 32);
}

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. SC'11).  Stream layout (identical in oracle/fw_oracle.c):
//   key = seed;  counter = (env_id lo32, episode lo32, purpose << 28 | env_id hi bits, block)
enum { RNG_RESET = 0, RNG_NOISE = 1, RNG_RESAMPLE = 2, RNG_ACTION = 3, RNG_OBS = 4, RNG_OBS_INIT = 5, RNG_MODEL = 6 };

__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ uint4 rng_block(unsigned long long seed, long long env_id, unsigned long long episode,
                                           uint32_t purpose, uint32_t block) {
    uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 ctr = make_uint4((uint32_t)env_id, (uint32_t)episode,
                           (purpose << 28) | ((uint32_t)((unsigned long long)env_id >> 32) & 0x0FFFFFFFu), block);
    return philox4x32(ctr, key);
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((unsigned long long)hi) << 21) ^ (((unsigned long long)lo) >> 11)) * (1.0 / 9007199254740992.0);
}
template <typename T>
__device__ __forceinline__ T rng_uniform(unsigned long long seed, long long env_id, unsigned long long episode,
                                         uint32_t purpose, int idx) {
    uint4 r = rng_block(seed, env_id, episode, purpose, (uint32_t)(idx >> 1));
    return (T)((idx & 1) ? u53(r.z, r.w) : u53(r.x, r.y));
}
// four unit normals for turbulence sample k (Box-Muller on 32-bit uniforms; evaluated in double in both precisions:
// it is 4 transcendentals per env-step against ~200 in the integrator)
__device__ __forceinline__ void rng_noise4(unsigned long long seed, long long env_id, unsigned long long episode,
                                           uint32_t k, double out[4]) {
    uint4 r = rng_block(seed, env_id, episode, RNG_NOISE, k);
    uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double u1 = ((double)w[2 * i] + 0.5) * (1.0 / 4294967296.0);
        double u2 = ((double)w[2 * i + 1] + 0.5) * (1.0 / 4294967296.0);
        double rad = ::sqrt(-2.0 * ::log(u1)), s, c;
        ::sincos(6.283185307179586476925 * u2, &s, &c);
        out[2 * i] = rad * c;
        out[2 * i + 1] = rad * s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T clip(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
template <typename T> __device__ __forceinline__ T sgn(T v) { return (T)((v > (T)0) - (v < (T)0)); }

// out-of-line math for code that runs once per step (post-step commit, reset): one shared copy each instead of an
// inlined expansion per call site keeps the kernel's cold instruction footprint small
template <typename T> __device__ __noinline__ T atan2_ni(T a, T b) { return M<T>::atan2(a, b); }
template <typename T> __device__ __noinline__ T asin_ni(T a) { return M<T>::asin(a); }
template <typename T> __device__ __noinline__ void sincos_ni(T a, T* sn, T* cs) { M<T>::sincos(a, sn, cs); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// per-step dynamics context that is constant during one integration
template <typename T> struct DynCtx {
    T cmd[3];      // constrained commands of elevon_right, elevon_left, throttle
    T wind[3];     // steady wind NED
    T tl[3];       // linear turbulence sample (body frame)
    T ta[3];       // angular turbulence sample
};

// _rot_b_v Euler branch (pyfly.py:1757-1777) applied to a vector
template <typename T>
__device__ __forceinline__ void rot_euler_apply(T phi, T th, T psi, const T v[3], T out[3]) {
    T sph, cph, sth, cth, sps, cps;
    sincos_ni<T>(phi, &sph, &cph);
    sincos_ni<T>(th, &sth, &cth);
    sincos_ni<T>(psi, &sps, &cps);
    out[0] = cth * cps * v[0] + cth * sps * v[1] + (-sth) * v[2];
    out[1] = (sph * sth * cps - cph * sps) * v[0] + (sph * sth * sps + cph * cps) * v[1] + sph * cth * v[2];
    out[2] = (cph * sth * cps + sph * sps) * v[0] + (cph * sth * sps - sph * cps) * v[1] + cph * cth * v[2];
}

// PyFly._dynamics with _forces inlined.  `first` = the t == 0 call of solve_ivp (no state write-back, pyfly.py:1461):
// constraint checks are skipped and elevator/aileron come from `elev0/ail0` (they are 0 right after a reset because
// disabled ControlVariables reset to 0, pyfly.py:359-363).  Returns 0 or a FwTermCode.
// PE: the aircraft parameters come from the env's column of S.par (`pe` = S.par + env, stride `pn`) instead of the
// warp-uniform constant bank; with PE = false PRM(x) is c.x exactly as before.
template <typename T, bool TURB, bool PE = false, bool CT = true>
__device__ __forceinline__ int rhs(const DCfg<T>& c, const DynCtx<T>& x, const T (&y)[FW_NY], bool first, T elev0,
                                   T ail0, T (&dy)[FW_NY], const T* pe = nullptr, int pn = 0) {
#define PRM(name) (PE ? __ldg(pe + (size_t)(FW_NPARAM + PD_##name) * pn) : c.name)
    const T e0 = y[0], e1 = y[1], e2 = y[2], e3 = y[3];
    const T P = y[4], Q = y[5], R = y[6];
    const T u = y[10], v = y[11], w = y[12];
    // Constraint violations are carried as a code to the end instead of returning early: the RHS
    // stays ONE basic block, so the scheduler can interleave its independent chains; a failing env's derivatives are
    // never used (the caller stops at rc != 0), and the first violation in evaluation order wins as before.
    int rc_con = 0;
    if (!first) {
        // _set_states_from_ode_solution(save=False): ConstraintException on p, q, r (pyfly.py:1872-1874)
        if (R < c.omega_con_min[2] || R > c.omega_con_max[2]) rc_con = FW_TERM_OMEGA_R;
        if (Q < c.omega_con_min[1] || Q > c.omega_con_max[1]) rc_con = FW_TERM_OMEGA_Q;
        if (P < c.omega_con_min[0] || P > c.omega_con_max[0]) rc_con = FW_TERM_OMEGA_P;
    }
    // Actuation.set_states: the RHS sees CLIPPED actuator values / rates (pyfly.py:471-492, 312-328)
    const T er = clip(y[13], c.elevon_min, c.elevon_max), el = clip(y[14], c.elevon_min, c.elevon_max);
    const T thr = clip(y[15], c.throttle_min, c.throttle_max);
    const T erd = clip(y[16], -c.elevon_dot_max, c.elevon_dot_max), eld = clip(y[17], -c.elevon_dot_max, c.elevon_dot_max);
    const T elevator = first ? elev0 : (er + el) / (T)2;
    const T aileron = first ? ail0 : (-er + el) / (T)2;

    T p = P, q = Q, r = R;
    if (TURB) { p -= x.ta[0]; q -= x.ta[1]; r -= x.ta[2]; }
    // _calculate_airspeed_factors with the quaternion rotation (un-normalised quaternion, pyfly.py:1464,1782-1800)
    const T r00 = (T)-1 + (T)2 * (e0 * e0 + e1 * e1), r01 = (T)2 * (e1 * e2 + e3 * e0), r02 = (T)2 * (e1 * e3 - e2 * e0);
    const T r10 = (T)2 * (e1 * e2 - e3 * e0), r11 = (T)-1 + (T)2 * (e0 * e0 + e2 * e2), r12 = (T)2 * (e2 * e3 + e1 * e0);
    const T r20 = (T)2 * (e1 * e3 + e2 * e0), r21 = (T)2 * (e2 * e3 - e1 * e0), r22 = (T)-1 + (T)2 * (e0 * e0 + e3 * e3);
    T a0 = u - (r00 * x.wind[0] + r01 * x.wind[1] + r02 * x.wind[2] + (TURB ? x.tl[0] : (T)0));
    T a1 = v - (r10 * x.wind[0] + r11 * x.wind[1] + r12 * x.wind[2] + (TURB ? x.tl[1] : (T)0));
    T a2 = w - (r20 * x.wind[0] + r21 * x.wind[1] + r22 * x.wind[2] + (TURB ? x.tl[2] : (T)0));
    // airspeed triangle.  One reciprocal square root per length: Va = s * rsqrt(s) and 1/Va = rsqrt(s) are within
    // 1-2 ulp of the reference's sqrt / divide, far inside the 1e-9 parity bar, and cost a third of sqrt + 3 divides.
    const T s_xz = a0 * a0 + a2 * a2, s_all = a0 * a0 + a1 * a1 + a2 * a2;
    const T inv_Va_raw = M<T>::rsqrt_hot(s_all), inv_rxz = M<T>::rsqrt_hot(s_xz);
    T Va = s_all * inv_Va_raw;
    const T rxz = s_xz * inv_rxz;
    const T alpha = M<T>::template atan2_hot<CT>(a2, a0);
    const T sb = a1 * inv_Va_raw;               // == sin(beta): beta = asin(a1 / Va) (pyfly.py:1848)
    const T beta = M<T>::template asin_hot<CT>(sb);
    if (rc_con == 0 && c.va_con_max > (T)0 && Va > c.va_con_max) rc_con = FW_TERM_VA;
    // sin/cos of alpha = atan2(a2, a0) and cos of beta = asin(a1/Va) follow from the triangle without any
    // trigonometric evaluation (identical up to rounding): sin a = a2/r, cos a = a0/r, cos b = r/Va, r = |(a0, a2)|
    const T sa = (s_xz > (T)0) ? a2 * inv_rxz : (T)0, ca = (s_xz > (T)0) ? a0 * inv_rxz : (T)1;
    const T cb = (s_xz > (T)0) ? rxz * inv_Va_raw : (T)0;
    T inv2Va = (T)0.5 * inv_Va_raw;
    if (!(s_all > (T)0)) Va = (T)0;             // rsqrt(0) = inf: keep Va = 0 like sqrt(0)
    if (Va < c.va_value_min) { Va = c.va_value_min; inv2Va = (T)1 / ((T)2 * Va); }

    const T pre = c.half_rho * (Va * Va) * PRM(S_wing);
    const T fgx = PRM(mg) * ((T)2 * (e1 * e3 - e2 * e0)), fgy = PRM(mg) * ((T)2 * (e2 * e3 + e1 * e0));
    const T fgz = PRM(mg) * (e3 * e3 + e0 * e0 - e1 * e1 - e2 * e2);
    const T CLlin = PRM(C_L_0) + PRM(C_L_alpha) * alpha;
    T sigma;
    if (sizeof(T) == 8) {
        // sigma = (1 + e1 + e2) / ((1 + e1)(1 + e2)), e1 = exp(-M(a - a0)), e2 = exp(M(a + a0))  (pyfly.py:1541-1543).
        // With E = exp(M a), C = exp(M a0): e1 = C / E, e2 = C E, and multiplying through by E gives the same value
        // from ONE exponential and ONE division, all terms positive (no cancellation): |a| <= pi keeps E^2 < 1e137.
        const T E = M<T>::template exp_hot<CT>(PRM(M_) * alpha);
        sigma = (E + PRM(exp_M_a0) + PRM(exp_M_a0) * (E * E)) * M<T>::rcp_hot((E + PRM(exp_M_a0)) * ((T)1 + PRM(exp_M_a0) * E));
    } else {
        // overflow-safe in fp32: sigma = 1 - s(-M(a-a0)) s(M(a+a0)), s = logistic
        const T g1 = M<T>::exp(PRM(M_) * (alpha - PRM(a_0))), g2 = M<T>::exp(-PRM(M_) * (alpha + PRM(a_0)));
        sigma = (T)1 - (T)1 / (((T)1 + g1) * ((T)1 + g2));
    }
    const T sg = sgn(alpha);
    const T C_L = ((T)1 - sigma) * CLlin + sigma * ((T)2 * sg * (sa * sa) * ca);
    const T lift = pre * (C_L + PRM(C_L_q) * PRM(c) * inv2Va * q + PRM(C_L_delta_e) * elevator);
    const T C_Da = PRM(C_D_p) + ((T)1 - sigma) * (CLlin * CLlin) * PRM(inv_pi_e_ar) + sigma * ((T)2 * sg * (sa * sa * sa));
    const T C_Db = PRM(C_D_beta1) * beta + PRM(C_D_beta2) * (beta * beta);
    const T drag = pre * (C_Da + C_Db + PRM(C_D_q) * PRM(c) * inv2Va * q + PRM(C_D_delta_e) * (elevator * elevator));
    const T C_m = ((T)1 - sigma) * (PRM(C_m_0) + PRM(C_m_alpha) * alpha) + sigma * (PRM(C_m_fp) * sg * (sa * sa));
    const T bq = PRM(b) * inv2Va;
    const T m_ = pre * PRM(c) * (C_m + PRM(C_m_q) * bq * q + PRM(C_m_delta_e) * elevator);   // sic: b (pyfly.py:1579)
    const T fy = pre * (PRM(C_Y_0) + PRM(C_Y_beta) * beta + PRM(C_Y_p) * bq * p + PRM(C_Y_r) * bq * r + PRM(C_Y_delta_a) * aileron);
    const T l_ = pre * PRM(b) * (PRM(C_l_0) + PRM(C_l_beta) * beta + PRM(C_l_p) * bq * p + PRM(C_l_r) * bq * r + PRM(C_l_delta_a) * aileron);
    const T n_ = pre * PRM(b) * (PRM(C_n_0) + PRM(C_n_beta) * beta + PRM(C_n_p) * bq * p + PRM(C_n_r) * bq * r + PRM(C_n_delta_a) * aileron);
    // f_aero = R_euler(0, alpha, beta) . [-drag, fy, -lift]  (pyfly.py:1617-1620; phi = 0 -> sin 0, cos 1)
    const T s0 = -drag, s1 = fy, s2 = -lift;
    const T fax = ca * cb * s0 + ca * sb * s1 + (-sa) * s2;
    const T fay = (-sb) * s0 + cb * s1;
    const T faz = sa * cb * s0 + sa * sb * s1 + ca * s2;
    const T Vd = Va + thr * (PRM(k_motor) - Va);
    const T fprop = PRM(prop_k) * Vd * (Vd - Va);
    const T kt = PRM(k_Omega) * thr;
    const T fx = fprop + fgx + fax, fyb = fgy + fay, fz = fgz + faz;
    const T tx = l_ + (-PRM(k_T_P) * (kt * kt)), ty = m_, tz = n_;

    // _f_attitude_dot uses the STATE omega (not turbulence corrected) (pyfly.py:1466,1476)
    dy[0] = (T)0.5 * (-P * e1 - Q * e2 - R * e3);
    dy[1] = (T)0.5 * (P * e0 + R * e2 - Q * e3);
    dy[2] = (T)0.5 * (Q * e0 - R * e1 + P * e3);
    dy[3] = (T)0.5 * (R * e0 + Q * e1 - P * e2);
    dy[4] = c.gam[1] * P * Q - c.gam[2] * Q * R + c.gam[3] * tx + c.gam[4] * tz;
    dy[5] = c.gam[5] * P * R - c.gam[6] * (P * P - R * R) + ty * c.inv_Jy;
    dy[6] = c.gam[7] * P * Q - c.gam[1] * Q * R + c.gam[4] * tx + c.gam[8] * tz;
    // _f_p_dot: R(q)^T-like matrix of pyfly.py:1718-1736
    dy[7] = (e1 * e1 + e0 * e0 - e2 * e2 - e3 * e3) * u + (T)2 * (e1 * e2 - e3 * e0) * v + (T)2 * (e1 * e3 + e2 * e0) * w;
    dy[8] = (T)2 * (e1 * e2 + e3 * e0) * u + (e2 * e2 + e0 * e0 - e1 * e1 - e3 * e3) * v + (T)2 * (e2 * e3 - e1 * e0) * w;
    dy[9] = (T)2 * (e1 * e3 - e2 * e0) * u + (T)2 * (e2 * e3 + e1 * e0) * v + (e3 * e3 + e0 * e0 - e1 * e1 - e2 * e2) * w;
    dy[10] = R * v - Q * w + fx * PRM(inv_mass);
    dy[11] = P * w - R * u + fyb * PRM(inv_mass);
    dy[12] = Q * u - P * v + fz * PRM(inv_mass);
    // Actuation.rhs (pyfly.py:519-543): elevons 2nd order on clipped value/rate, throttle 1st order
    dy[13] = erd;
    dy[14] = eld;
    dy[15] = thr * (-c.inv_tau) + x.cmd[2] * c.inv_tau;
    dy[16] = er * (-c.w0sq) + x.cmd[0] * c.w0sq + erd * (-c.two_zeta_w0);
    dy[17] = el * (-c.w0sq) + x.cmd[1] * c.w0sq + eld * (-c.two_zeta_w0);
    dy[18] = (T)0;
    return rc_con;
#undef PRM
}

// Dormand-Prince 5(4) tableau (scipy rk.py class RK45) as "evaluation rows": row r gives the coefficients of the
// stage derivatives K[0..] that form the input of RHS evaluation r.
//   rows 1..5  : y + h * sum_{j<r} A[r][j] K[j]     -> K[r]                      (rk_step)
//   row 6      : y + h * sum_j B[j] K[j] = y_new    -> f_new (row 6 of A is B: FSAL)
__device__ __constant__ const double RK_ROW[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
__device__ __constant__ const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200,
                                                -22.0 / 525, 1.0 / 40};

template <typename T> __device__ __noinline__ T pow_ni(T x, T y) { return M<T>::pow(x, y); }

#define FW_NK 18   // ODE components that evolve: d/dt of y[18] (throttle rate state) is identically 0
#define FW_NS 15   // of those, components whose stage derivatives are stored (all but position y[7..9])
#define FW_NP 8    // ... held in shared memory as pairs (the last pair is half empty)
template <typename T> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

// scipy.integrate.solve_ivp(fun, (0, dt), y0) with RK45 defaults (pyfly.py:1393-1395) is split over two kernels:
//
//   rk45_init   rows 0 and 7 for every env in lock step (RungeKutta.__init__ + select_initial_step, common.py:68-134):
//               f0 and the first step size go to a scratch SoA.
//   the attempt loop (rk45_attempt_kernel in fw_step.cu): every lane always executes the same row 1..6 of an
//               attempt, so ONE inlined RHS instance serves all lanes at full occupancy of the warp; a lane whose env
//               finished its [0, dt] interval pulls the next env from a queue instead of idling (the number of
//               attempts per env-step varies 2..7, which cost 32 % of the lanes when envs were pinned to threads).
// A load the compiler may not satisfy from a register copy: the init kernel re-reads y0 and f0 after the second RHS
// evaluation instead of keeping 38 values live across it (which cost a fourth warp per scheduler, or spills).
__device__ __forceinline__ double ld_again(const double* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_again(const float* p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// ys enters holding y0 (which is also at ysrc[i * n]); f0 is written to f0dst[i * n] (i < FW_NK).
template <typename T, bool TURB, bool PE = false>
__device__ __forceinline__ int rk45_init(const DCfg<T>& c, const DynCtx<T>& x, T (&ys)[FW_NY], const T* ysrc, T* f0dst,
                                         int n, T elev0, T ail0, T& h_abs, const T* pe = nullptr) {
    const T rtol = c.rtol, atol = c.atol, t_bound = c.dt;
    T dyv[FW_NY];
    T d1 = 0, h0 = 0;
    int rc = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        rc = rhs<T, TURB, PE, false>(c, x, ys, pass == 0, elev0, ail0, dyv, pe, n);
        if (rc) return rc;
        if (pass == 0) {
            T s0 = 0, s1 = 0;
#pragma unroll
            for (int i = 0; i < FW_NY; ++i) {
                const T inv = M<T>::rcp_hot(atol + M<T>::fabs(ys[i]) * rtol);      // >= atol: normal range
                const T a = ys[i] * inv, b = dyv[i] * inv;
                s0 += a * a;
                s1 += b * b;
                if (i < FW_NK) f0dst[i * n] = dyv[i];
            }
            const T d0 = M<T>::sqrt(s0) / M<T>::sqrt((T)FW_NY);
            d1 = M<T>::sqrt(s1) / M<T>::sqrt((T)FW_NY);
            h0 = (d0 < (T)1e-5 || d1 < (T)1e-5) ? (T)1e-6 : (T)0.01 * d0 / d1;
            h0 = M<T>::fmin(h0, t_bound);
#pragma unroll
            for (int i = 0; i < FW_NY; ++i) ys[i] = ys[i] + h0 * dyv[i];
        } else {
            T s2 = 0;
#pragma unroll
            for (int i = 0; i < FW_NY; ++i) {
                const T y0 = ld_again(ysrc + i * n), f0 = (i < FW_NK) ? ld_again(f0dst + i * n) : (T)0;
                const T inv = M<T>::rcp_hot(atol + M<T>::fabs(y0) * rtol);
                const T d = (dyv[i] - f0) * inv;
                s2 += d * d;
            }
            const T d2 = (M<T>::sqrt(s2) / M<T>::sqrt((T)FW_NY)) / h0;
            T h1;
            if (d1 <= (T)1e-15 && d2 <= (T)1e-15) h1 = M<T>::fmax((T)1e-6, h0 * (T)1e-3);
            else h1 = M<T>::template pow_hot<false>((T)0.01 / M<T>::fmax(d1, d2), (T)0.2);
            h_abs = M<T>::fmin(M<T>::fmin((T)100 * h0, h1), t_bound);
        }
    }
    return 0;
}

template <typename T> __device__ __forceinline__ T ulp10(T t) {   // min_step = 10 * |nextafter(t, inf) - t| (rk.py:118)
    return (T)10 * (sizeof(T) == 8 ? (T)(::nextafter((double)t, CUDART_INF) - (double)t)
                                   : (T)(::nextafterf((float)t, CUDART_INF_F) - (float)t));
}

// classical RK4 x substeps, register resident (throughput mode; same RHS with its clip/constraint side effects),
// also one loop around one RHS instance.
template <typename T, bool TURB, bool PE = false>
__device__ __forceinline__ int solve_rk4(const DCfg<T>& c, const DynCtx<T>& x, T (&y)[FW_NY], T elev0, T ail0,
                                         int& nfev, int& natt, const T* pe = nullptr, int pn = 0) {
    const int n = c.rk4_substeps > 0 ? c.rk4_substeps : 1;
    const T h = c.dt / (T)n;
    T k[FW_NY], acc[FW_NK], ys[FW_NY];
    int rc = 0;
    nfev = 0;
    natt = n;
#pragma unroll
    for (int i = 0; i < FW_NK; ++i) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < FW_NY; ++i) k[i] = 0;
#pragma unroll 1
    for (int e = 0; e < 4 * n; ++e) {
        const int st = e & 3;
        const T cs = (st == 0) ? (T)0 : ((st == 3) ? h : (T)0.5 * h);
#pragma unroll
        for (int i = 0; i < FW_NK; ++i) ys[i] = y[i] + cs * k[i];
        ys[18] = y[18];
        rc = rhs<T, TURB, PE>(c, x, ys, e == 0, elev0, ail0, k, pe, pn);
        nfev++;
        if (rc) break;
        const T w = (st == 1 || st == 2) ? (T)2 : (T)1;
#pragma unroll
        for (int i = 0; i < FW_NK; ++i) acc[i] = (st == 0) ? k[i] : acc[i] + w * k[i];
        if (st == 3) {
#pragma unroll
            for (int i = 0; i < FW_NK; ++i) y[i] = y[i] + h / (T)6 * acc[i];
        }
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// gym-level helpers

template <typename T> __device__ __forceinline__ T py_mod(T a, T b) {
    T m = M<T>::fmod(a, b);
    if (m != (T)0 && ((m < (T)0) != (b < (T)0))) m += b;
    return m;
}
// _get_error (fixed_wing.py:1318-1344); roll wraps (pyfly_config.json "wrap": true):
//   dist = (value - target + pi) % (2 pi) - pi.   Python's float % is fmod plus a sign fix-up; for |x| < 4 pi the exact
// fmod is a single Sterbenz-exact subtraction, so the generic (large, loop-based) fmod is kept off the hot path.
template <typename T> __device__ __noinline__ T py_mod_generic(T a, T b) { return py_mod(a, b); }
template <typename T> __device__ __forceinline__ T err_roll(T target, T value) {
    const T PI = (T)3.141592653589793238462643383279502884, TWO_PI = (T)2 * PI;
    const T xx = value - target + PI;
    T m;
    if (xx >= (T)0 && xx < TWO_PI) m = xx;
    else if (xx >= TWO_PI && xx < (T)2 * TWO_PI) m = xx - TWO_PI;      // exact (Sterbenz)
    else if (xx < (T)0 && xx > -TWO_PI) m = xx + TWO_PI;              // fmod keeps x, python adds the divisor
    else m = py_mod_generic<T>(xx, TWO_PI);
    T dist = m - PI;
    if (dist < -PI) dist += TWO_PI;
    return dist;
}

// numpy pairwise-sum order for n <= 128 (see oracle np_sum_*): matters only for the float32 action path
template <typename F> __device__ __forceinline__ F np_sum(const F* a, int n) {
    if (n < 8) { F s = 0; for (int i = 0; i < n; ++i) s += a[i]; return s; }
    F r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    F res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// the same order with a compile-time length: everything stays in registers
template <typename F, int N> __device__ __forceinline__ F np_sum_n(const F (&a)[N]) {
    if (N < 8) {
        F s = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) s += a[i];
        return s;
    }
    F r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j < N ? j : 0];
#pragma unroll
    for (int i = 8; i < N - (N % 8); i += 8)
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a[i + j < N ? i + j : 0];
    F res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
    for (int i = (N < 8 ? 0 : N - (N % 8)); i < N; ++i) res += a[i];
    return res;
}

// sum over the trailing NP action deltas |a_t - a_{t-1}|, 3 columns each, chronological order (oldest first), of the
// action ring (age 1 = most recent previous action) plus the current action; F = float for float32 actions
template <int NP, typename F, typename T>
__device__ __forceinline__ F ring_delta_sum(const T (&cur)[3], const T (&ring)[12]) {
    F d[NP * 3];
#pragma unroll
    for (int e = 0; e < NP * 3; ++e) {
        const int age = NP - e / 3, j = e % 3;
        const T newer = (age == 1) ? cur[j] : ring[(age >= 2 ? age - 2 : 0) * 3 + j];
        const F x = (F)newer - (F)ring[(age - 1) * 3 + j];
        d[e] = M<F>::fabs(x);
    }
    return np_sum_n<F, NP * 3>(d);
}
template <typename F, typename T>
__device__ __forceinline__ F ring_delta_sum_np(int np_, const T (&cur)[3], const T (&ring)[12]) {
    switch (np_) {
        case 1: return ring_delta_sum<1, F, T>(cur, ring);
        case 2: return ring_delta_sum<2, F, T>(cur, ring);
        case 3: return ring_delta_sum<3, F, T>(cur, ring);
        case 4: return ring_delta_sum<4, F, T>(cur, ring);
        default: return (F)0;
    }
}

// register-resident view of one env during a step
template <typename T> struct EnvRegs {
    T y[FW_NY];
    T roll, pitch, Va, alpha, beta;
    T wind[3], tgt[3];
    int steps, steps_tgt, sim_step;
    unsigned long long episode;
};

// The six Dryden shaping filters have fixed orders (dryden.py:126-143): H_u 1, H_v 2, H_w 2, H_p 1, H_q 3, H_r 3,
// so their states pack into fx[12] at offsets 0,1,3,5,6,9 and every loop below has a compile-time trip count.
// The noise row of each filter is fixed too (dryden.py:238-252: u<-0, v<-1, w<-2, p<-3, q<-1, r<-2): a compile-time
// index keeps fu / un in registers (a runtime index put them in local memory: 8 % of the head kernel's stall samples).
template <typename T, int ORD, int OFF, int ROW>
__device__ __forceinline__ T filt_out(const DFilter<T>& F, const T (&fx)[12], const T (&fu)[4]) {
    T yv = 0;
#pragma unroll
    for (int a = 0; a < ORD; ++a) yv += fx[OFF + a] * F.C[a];
    return yv + fu[ROW] * F.D;
}
template <typename T, int ORD, int OFF, int ROW>
__device__ __forceinline__ void filt_adv(const DFilter<T>& F, T (&fx)[12], const T (&fu)[4], const T (&un)[4]) {
    T xn[ORD];
    const T up = fu[ROW], uc = un[ROW];
#pragma unroll
    for (int a = 0; a < ORD; ++a) {
        T sacc = 0;
#pragma unroll
        for (int b = 0; b < ORD; ++b) sacc += fx[OFF + b] * F.Ad[b * ORD + a];
        xn[a] = sacc + up * F.Bd0[a] + uc * F.Bd1[a];
    }
#pragma unroll
    for (int a = 0; a < ORD; ++a) fx[OFF + a] = xn[a];
}
// y = C x + D u  (lsim output equation, dryden.py:22-39)
template <typename T>
__device__ __forceinline__ void turb_eval(const DCfg<T>& c, const T (&fx)[12], const T (&fu)[4], T (&tl)[3], T (&ta)[3]) {
    tl[0] = filt_out<T, 1, 0, 0>(c.filt[0], fx, fu);
    tl[1] = filt_out<T, 2, 1, 1>(c.filt[1], fx, fu);
    tl[2] = filt_out<T, 2, 3, 2>(c.filt[2], fx, fu);
    ta[0] = filt_out<T, 1, 5, 3>(c.filt[3], fx, fu);
    ta[1] = filt_out<T, 3, 6, 1>(c.filt[4], fx, fu);
    ta[2] = filt_out<T, 3, 9, 2>(c.filt[5], fx, fu);
}
// x_{k+1} = Ad x_k + Bd0 u_k + Bd1 u_{k+1}  (lsim recurrence, row-vector convention)
template <typename T>
__device__ __forceinline__ void turb_advance(const DCfg<T>& c, T (&fx)[12], T (&fu)[4], const T (&un)[4]) {
    filt_adv<T, 1, 0, 0>(c.filt[0], fx, fu, un);
    filt_adv<T, 2, 1, 1>(c.filt[1], fx, fu, un);
    filt_adv<T, 2, 3, 2>(c.filt[2], fx, fu, un);
    filt_adv<T, 1, 5, 3>(c.filt[3], fx, fu, un);
    filt_adv<T, 3, 6, 1>(c.filt[4], fx, fu, un);
    filt_adv<T, 3, 9, 2>(c.filt[5], fx, fu, un);
#pragma unroll
    for (int r = 0; r < 4; ++r) fu[r] = un[r];
}

// x <- x Ablk (row-vector convention) for one filter
template <typename T, int ORD, int OFF>
__device__ __forceinline__ void filt_block(const DFilter<T>& F, T (&fx)[12]) {
    T xn[ORD];
#pragma unroll
    for (int a = 0; a < ORD; ++a) {
        T sacc = 0;
#pragma unroll
        for (int b = 0; b < ORD; ++b) sacc += fx[OFF + b] * F.Ablk[b * ORD + a];
        xn[a] = sacc;
    }
#pragma unroll
    for (int a = 0; a < ORD; ++a) fx[OFF + a] = xn[a];
}

// One step of the streamed turbulence: sample index `k_new` becomes current.  pyfly simulates blocks of
// turbulence_sim_length samples; every new block calls lsim with T[0] > 0, which first steps the carried state over
// [0, T[0]] with zero input: x <- x Ablk^m for block m, and the block's first output is C x + D u without a recurrence
// step (pyfly.py:870-871, dryden.py:30-36, scipy lsim).  Inside a block it is the plain recurrence.
template <typename T>
__device__ __forceinline__ void turb_step(const DCfg<T>& c, T (&fx)[12], T (&fu)[4], const T (&un)[4], int k_new) {
    if (c.turb_block_len > 0 && (k_new % c.turb_block_len) == 0) {
        for (int m = 0; m < k_new / c.turb_block_len; ++m) {
            filt_block<T, 1, 0>(c.filt[0], fx); filt_block<T, 2, 1>(c.filt[1], fx); filt_block<T, 2, 3>(c.filt[2], fx);
            filt_block<T, 1, 5>(c.filt[3], fx); filt_block<T, 3, 6>(c.filt[4], fx); filt_block<T, 3, 9>(c.filt[5], fx);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) fu[r] = un[r];
    } else {
        turb_advance(c, fx, fu, un);
    }
}

// scaled noise sample k for env (injected buffer or Philox)
template <typename T>
__device__ __forceinline__ void noise_sample(const DCfg<T>& c, const Soa<T>& S, int env, unsigned long long episode,
                                             int k, T (&un)[4]) {
    double z[4];
    if (S.noise) {
        const int kk = k < S.noise_len ? k : S.noise_len - 1;
#pragma unroll
        for (int r = 0; r < 4; ++r) z[r] = S.noise[((size_t)env * 4 + r) * S.noise_len + kk];
    } else {
        rng_noise4(env_seed(S, env), c.env_id_offset + env, episode, (uint32_t)k, z);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) un[r] = (T)z[r] * c.turb_noise_scale;
}

// the twelve uniform draws of one target sampling, four per target state in the order the reference consumes them
// (initial value; slope, sign | amplitude, period, phase): Philox blocks block0 .. block0+5 of `purpose`, or the override
template <typename T>
__device__ __forceinline__ void target_draws(const DCfg<T>& c, unsigned long long seed, long long gid,
                                             unsigned long long episode, uint32_t purpose, uint32_t block0, T (&u12)[12]) {
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const uint4 rr = rng_block(seed, gid, episode, purpose, block0 + (uint32_t)b);
        u12[2 * b] = (T)u53(rr.x, rr.y);
        u12[2 * b + 1] = (T)u53(rr.z, rr.w);
    }
    if (!M<T>::isnan(c.rng_u_override)) {
#pragma unroll
        for (int i = 0; i < 12; ++i) u12[i] = c.rng_u_override;
    }
}

// sample_target (fixed_wing.py:654-746): classes constant / compensate / linear / sinusoidal
template <typename T>
__device__ __forceinline__ void sample_target(const DCfg<T>& c, const ResetCfg<T>& rc, T roll, T pitch, T Va, int steps,
                                              const T (&u12)[12], T (&tgt)[3], int (&tcls)[3], T (&tp)[15]) {
    const T val[3] = {roll, pitch, Va};
    const T TWO_PI = (T)6.283185307179586476925286766559, D2R = (T)(3.141592653589793238462643383279502884 / 180.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        T low = rc.tgt_low[k], high = rc.tgt_high[k];
        if (!M<T>::isnan(rc.tgt_delta[k])) {
            low = M<T>::fmax(low, val[k] - rc.tgt_delta[k]);
            high = M<T>::fmax(M<T>::fmin(high, val[k] + rc.tgt_delta[k]), low);
        }
        const T initial = low + (high - low) * u12[4 * k];
        tcls[k] = c.tgt_class[k];
        if (c.tgt_class[k] == FW_TGT_LINEAR) {
            T slope = rc.tgt_slope_low[k] + (rc.tgt_slope_high[k] - rc.tgt_slope_low[k]) * u12[4 * k + 1];
            if (u12[4 * k + 2] < (T)0.5) slope *= (T)-1;
            if (c.tgt_radians[k]) slope = slope * D2R;
            tp[k] = slope;
        } else if (c.tgt_class[k] == FW_TGT_SINUSOIDAL) {
            T amp = rc.tgt_amp_low[k] + (rc.tgt_amp_high[k] - rc.tgt_amp_low[k]) * u12[4 * k + 1];
            if (c.tgt_radians[k]) amp = amp * D2R;
            const T period = rc.tgt_period_low[k] + (rc.tgt_period_high[k] - rc.tgt_period_low[k]) * u12[4 * k + 2];
            const T phase = ((T)0 + (TWO_PI - (T)0) * u12[4 * k + 3]) / (TWO_PI / period);
            tp[3 + k] = amp; tp[6 + k] = period; tp[9 + k] = phase;
            tp[12 + k] = initial - amp * M<T>::sin(TWO_PI / period * ((T)steps + phase));
        }
        tgt[k] = initial;
    }
}

// observation.noise (fixed_wing.py:1246-1247): every entry += N(mean, var); Philox purpose OBS, block = 32 * steps + b,
// four Box-Muller normals per block (same stream as oracle add_obs_noise).  Out of line: off by default.
template <typename T>
__device__ __noinline__ void add_obs_noise(const DCfg<T>& c, unsigned long long seed, long long gid,
                                           unsigned long long episode, int steps, T* o, int dim) {
#pragma unroll 1
    for (int b = 0; b * 4 < dim; ++b) {
        const uint4 r = rng_block(seed, gid, episode, RNG_OBS, (uint32_t)(steps * 32 + b));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double u1 = ((double)w[2 * i] + 0.5) * (1.0 / 4294967296.0);
            const double u2 = ((double)w[2 * i + 1] + 0.5) * (1.0 / 4294967296.0);
            const double rad = ::sqrt(-2.0 * ::log(u1));
            double sn, cs;
            ::sincos(6.283185307179586476925 * u2, &sn, &cs);
            const int j = b * 4 + i * 2;
            if (j < dim) o[j] += c.obs_noise_mean + c.obs_noise_std * (T)(rad * cs);
            if (j + 1 < dim) o[j + 1] += c.obs_noise_mean + c.obs_noise_std * (T)(rad * sn);
        }
    }
}

template <typename T> __device__ __forceinline__ int obs_dim(const DCfg<T>& c) {
    if (c.env_kind == FW_ENV_WAYPOINT) return FW_NOBS_WAYPOINT;
    return c.obs_generic ? c.obs_len * c.obs_n : FW_NOBS;
}

// observation (fixed_wing.py:1113-1262): row-major [obs_len][obs_n] floats per env (the default layout is one row of 14)
template <typename T>
__device__ __forceinline__ void write_obs(const T* o, int dim, int env, float* obs, double* obs64) {
    if (obs) for (int j = 0; j < dim; ++j) obs[(size_t)env * dim + j] = (float)o[j];
    if (obs64) for (int j = 0; j < dim; ++j) obs64[(size_t)env * dim + j] = (double)o[j];
}

// ---------------- target class attitude_angular (fixed_wing.py:671-675, 741-746, 1455-1460, 1558-1642) ----------------
// _attitude_to_angular_rates for omega state a (0 p, 1 q, 2 r) from the CURRENT roll / pitch errors, attitude and rate
// targets.  The reference's `damping = 0.05` branches are overwritten unconditionally (kept so); divisions by cos / sin of
// the roll angle are unguarded there and here.
template <typename T>
__device__ __noinline__ T attitude_to_angular_rate(const DCfg<T>& c, int a, T roll, T pitch, T roll_err, T pitch_err,
                                                   const T (&atgt)[3]) {
    const T max_vel = c.ang_max_vel[a], HALF_PI = (T)(0.5 * 3.141592653589793238462643383279502884);
    T r_w, q_w;
    M<T>::sincos(roll, &r_w, &q_w);
    const T max_pitch_change = max_vel * c.dt * (q_w + r_w);
    T res, damping;
    if (a == 0) {
        damping = M<T>::fabs(roll_err / HALF_PI);
        const T tp = (T)::tan((double)pitch);
        const T q_roll = r_w * tp * atgt[1] * c.dt, r_roll = q_w * tp * atgt[2] * c.dt;
        res = clip(-(roll_err - q_roll - r_roll) / c.dt, -max_vel, max_vel);
    } else if (a == 1) {
        damping = M<T>::fabs(pitch_err / HALF_PI);
        if (max_pitch_change > M<T>::fabs(pitch_err)) res = -pitch_err / ((T)2 * q_w);
        else res = sgn(q_w) * max_vel * sgn(pitch_err);
    } else {
        damping = M<T>::fabs(pitch_err / HALF_PI);
        if (max_pitch_change > M<T>::fabs(pitch_err)) res = pitch_err / r_w;
        else res = -sgn(r_w) * max_vel * sgn(pitch_err);
    }
    if (M<T>::isnan(damping)) damping = (T)0.05; else damping = M<T>::fmin((T)1, damping);
    return clip(atgt[a] + (res * damping - atgt[a]) * (T)1 / (T)20, -max_vel, max_vel);
}

// the three rate targets from the same inputs (every value from the OLD targets, like the dict the reference builds
// before it assigns, fixed_wing.py:609-612); from_zero: sample_target's re-derivation from zeroed rate targets (:671-746)
template <typename T>
__device__ __noinline__ void angular_targets(const DCfg<T>& c, T roll, T pitch, T roll_err, T pitch_err, bool from_zero,
                                             T (&atgt)[3]) {
    T old[3] = {atgt[0], atgt[1], atgt[2]};
    if (from_zero) old[0] = old[1] = old[2] = 0;
#pragma unroll 1
    for (int a = 0; a < 3; ++a) atgt[a] = attitude_to_angular_rate<T>(c, a, roll, pitch, roll_err, pitch_err, old);
}

// goal ring / streak bookkeeping of the three rate targets for history["goal"] entry `idx` (the code of the base states
// in head_kernel, on the IF_AGOAL_* fields)
template <typename T>
__device__ __noinline__ void angular_goal_update(const DCfg<T>& c, const Soa<T>& S, int env, int idx, const int (&ga)[3]) {
    const int n = S.n;
    int32_t* ii = S.i + env;
    const int w = (idx & 127) >> 5, b = idx & 31, idx_old = idx - c.streak_req;
#pragma unroll 1
    for (int a = 0; a < 3; ++a) {
        int32_t* ring = ii + (IF_AGOAL_RING + 4 * a) * n;
        int cnt = ii[(IF_AGOAL_CNT + a) * n], tot = ii[(IF_AGOAL_TOTAL + a) * n], settle = ii[(IF_ASETTLE + a) * n];
        if (idx_old >= 0) cnt -= (ring[((idx_old & 127) >> 5) * n] >> (idx_old & 31)) & 1;
        uint32_t word = (uint32_t)ring[w * n];
        word = (word & ~(1u << b)) | ((uint32_t)ga[a] << b);
        ring[w * n] = (int32_t)word;
        cnt += ga[a]; tot += ga[a];
        if (settle < 0 && idx + 1 >= c.streak_req && (double)cnt / (double)c.streak_req >= (double)c.streak_fraction) settle = idx;
        ii[(IF_AGOAL_CNT + a) * n] = cnt; ii[(IF_AGOAL_TOTAL + a) * n] = tot; ii[(IF_ASETTLE + a) * n] = settle;
    }
}

// streamed error statistics of the rate targets for the new history["error"] entry (index n_err), as for the base states
template <typename T>
__device__ __noinline__ void angular_stats_update(const DCfg<T>& c, const Soa<T>& S, int env, int n_err, const T (&ea)[3]) {
    const int n = S.n;
    T* r = S.r + env;
    int32_t* ii = S.i + env;
#pragma unroll 1
    for (int a = 0; a < 3; ++a) {
        const T e0 = r[(RF_AE0 + a) * n], v = ea[a], av = M<T>::fabs(v), prev = r[(RF_AEPREV + a) * n];
        const T low_lim = M<T>::fabs(c.rise_low * e0), high_lim = M<T>::fabs(c.rise_high * e0);
        if (ii[(IF_ARISE_LO + a) * n] < 0 && prev >= low_lim && av < low_lim) ii[(IF_ARISE_LO + a) * n] = n_err - 1;
        if (ii[(IF_ARISE_HI + a) * n] < 0 && prev >= high_lim && av < high_lim) ii[(IF_ARISE_HI + a) * n] = n_err - 1;
        r[(RF_AESUM + a) * n] += v; r[(RF_AEABS + a) * n] += av;
        r[(RF_AEMIN + a) * n] = M<T>::fmin(r[(RF_AEMIN + a) * n], v); r[(RF_AEMAX + a) * n] = M<T>::fmax(r[(RF_AEMAX + a) * n], v);
        r[(RF_AEPREV + a) * n] = av;
        S.err_ring_a[(size_t)((n_err % FW_END_ERR_WINDOW) * 3 + a) * n + env] = v;
    }
}

// the 24 metrics of the rate targets at the end of an episode (get_metric, fixed_wing.py:1644-1736; metric-major)
template <typename T>
__device__ __noinline__ void angular_metrics(const DCfg<T>& c, const Soa<T>& S, int env, int n_err, int n_goal, int off) {
    const int n = S.n;
    const T* r = S.r + env;
    const int32_t* ii = S.i + env;
    double* m = S.metrics_a + (size_t)env * FW_NMETRIC_ANG;
    const int end_cnt = n_err < FW_END_ERR_WINDOW ? n_err : FW_END_ERR_WINDOW;
#pragma unroll 1
    for (int a = 0; a < 3; ++a) {
        const T e0 = r[(RF_AE0 + a) * n];
        T s50 = 0;
#pragma unroll 1
        for (int t = n_err - end_cnt; t < n_err; ++t) s50 += S.err_ring_a[(size_t)((t % FW_END_ERR_WINDOW) * 3 + a) * n + env];
        m[0 + a] = (M<T>::fabs(e0) >= (T)0.01) ? (double)M<T>::fabs((r[(RF_AESUM + a) * n] / (T)n_err) / e0) : CUDART_NAN;
        m[3 + a] = (double)r[(RF_AEABS + a) * n];
        m[6 + a] = (double)M<T>::fabs(s50 / (T)end_cnt);
        const int lo = ii[(IF_ARISE_LO + a) * n], hi = ii[(IF_ARISE_HI + a) * n];
        m[9 + a] = (lo >= 0 ? (double)(lo + off) : CUDART_NAN) - (hi >= 0 ? (double)(hi + off) : CUDART_NAN);
        const T opp = (e0 > (T)0) ? r[(RF_AEMIN + a) * n] : r[(RF_AEMAX + a) * n];
        m[12 + a] = (sgn(opp) == sgn(e0)) ? CUDART_NAN : (double)M<T>::fabs(opp / e0);
        const int settle = ii[(IF_ASETTLE + a) * n];
        m[15 + a] = settle >= 0 ? 1.0 : 0.0;
        m[18 + a] = settle >= 0 ? (double)settle : CUDART_NAN;
        m[21 + a] = (double)ii[(IF_AGOAL_TOTAL + a) * n] / (double)n_goal;
    }
}

// sum of history["error"][name][start:stop] (absolute entry indices, clamped like a python slice to [0, len]) from the
// 50-deep error ring of the end_error metric, oldest first.  Callers keep stop - start + lag below the ring depth.
template <typename T>
__device__ __forceinline__ T err_ring_sum(const Soa<T>& S, int env, int k, int start, int stop, int len) {
    if (start < 0) start = 0;
    if (stop > len) stop = len;
    const T* ring = k < 3 ? S.err_ring : S.err_ring_a;         // target states 3..5: the rate targets' own ring
    const int kk = k < 3 ? k : k - 3;
    T s = 0;
#pragma unroll 1
    for (int t = start; t < stop; ++t) s += ring[(size_t)((t % FW_END_ERR_WINDOW) * 3 + kk) * S.n + env];
    return s;
}

// The value an "integrator" observation entry takes in the reset observation that follows an episode whose error history
// has `len` entries and first entry e0: every row is clamped to lag 1 with steps_count = 0 (fixed_wing.py:1141-1180), so
// sum(history[-W-1:-1]) + (W + 1) * history[0].
__device__ __forceinline__ int e0_field(int k) { return k < 3 ? RF_E0 + k : RF_AE0 + k - 3; }
__device__ __forceinline__ int esum_field(int k) { return k < 3 ? RF_ESUM + k : RF_AESUM + k - 3; }

template <typename T>
__device__ __forceinline__ T integrator_reset_value(const DCfg<T>& c, const Soa<T>& S, int env, int k, int len, T e0) {
    const int W = c.integration_window;
    return err_ring_sum<T>(S, env, k, len - W - 1, len - 1, len) + (T)(W + 1) * e0;
}

// get_reward (fixed_wing.py:941-1111), the general engine: any list of factors with linear / exponential / quadratic
// function classes, shaping and plain parts per term, absolute or potential form.  Out of line: the default factor
// family has its own straight-line code in the head kernel.
template <typename T>
__device__ __noinline__ T generic_reward(const DCfg<T>& c, const Soa<T>& S, int env, const T (&eg)[6], const T (&st8)[8],
                                         const T (&a_raw)[3], bool act_f32, const T* aring, int n_prev, int steps,
                                         const int (&gbits)[7] /* roll pitch Va all omega_p omega_q omega_r */, bool success) {
    const int n = S.n;
    T val_t[3] = {0, 0, 0}, shp_t[3] = {0, 0, 0};
#pragma unroll 1
    for (int i = 0; i < c.rew_n; ++i) {
        T val = 0;
        const int cls = c.rew_class[i];
        if (cls == FW_RF_STATE_ERROR) val = eg[c.rew_idx[i]];
        else if (cls == FW_RF_STATE_VALUE) val = st8[c.rew_idx[i]];
        else if (cls == FW_RF_STATE_INT_ERROR) {
            // history["error"] holds `steps` entries here (this step's is appended after the reward); [-0:] is the whole
            // list, whose sum the metrics already carry (fixed_wing.py:1003-1012)
            const int W = c.integration_window, k = c.rew_idx[i];
            if (W == 0) val = S.r[esum_field(k) * n + env];
            else {
                val = err_ring_sum<T>(S, env, k, steps - W, steps, steps);
                if (steps < W) val += (T)(W - steps) * S.r[e0_field(k) * n + env];
            }
        } else if (cls == FW_RF_ACTION_VALUE) {
            if (act_f32) { float sa = 0.f; for (int j = 0; j < 3; ++j) sa += fabsf((float)a_raw[j]); val = (T)sa; }
            else for (int j = 0; j < 3; ++j) val += M<T>::fabs(a_raw[j]);
        } else if (cls == FW_RF_ACTION_DELTA) {
            if (steps > 1) {
                const int np_ = (c.rew_window[i] - 1) < n_prev ? (c.rew_window[i] - 1) : n_prev;
                if (act_f32) {
                    float d[12];
                    int m = 0;
                    for (int age = np_; age >= 1; --age)
                        for (int j = 0; j < 3; ++j) {
                            const T newer = (age == 1) ? a_raw[j] : aring[(age - 2) * 3 + j];
                            d[m++] = fabsf((float)newer - (float)aring[(age - 1) * 3 + j]);
                        }
                    val = (T)np_sum<float>(d, m);
                } else {
                    T d[12];
                    int m = 0;
                    for (int age = np_; age >= 1; --age)
                        for (int j = 0; j < 3; ++j) {
                            const T newer = (age == 1) ? a_raw[j] : aring[(age - 2) * 3 + j];
                            d[m++] = M<T>::fabs(newer - aring[(age - 1) * 3 + j]);
                        }
                    val = np_sum<T>(d, m);
                }
            }
        } else if (cls == FW_RF_ACTION_BOUND) {
            T hi = 0, lo = 0;
            for (int j = 0; j < 3; ++j) {
                if (a_raw[j] > c.action_bounds_max[j]) hi += M<T>::fabs(a_raw[j] - c.action_bounds_max[j]);
                if (a_raw[j] < c.action_bounds_min[j]) lo += M<T>::fabs(a_raw[j] - c.action_bounds_min[j]);
            }
            val = hi + lo;
        } else if (cls == FW_RF_SUCCESS) {
            val = success ? (c.rew_value_timesteps[i] ? (T)(c.steps_max - steps) : c.rew_value[i]) : (T)0;
        } else if (cls == FW_RF_STEP) val = c.rew_value[i];
        else if (cls == FW_RF_GOAL_PER_STATE) {          // value / len(self.target) per achieved state
            const T nt = c.ang_on ? (T)6 : (T)3;
            for (int k = 0; k < 3; ++k) val += gbits[k] ? c.rew_value[i] / nt : (T)0;
            if (c.ang_on) for (int k = 4; k < 7; ++k) val += gbits[k] ? c.rew_value[i] / nt : (T)0;
        } else if (cls == FW_RF_GOAL_ALL) val = gbits[3] ? c.rew_value[i] : (T)0;
        // values derived from a float32 action array stay float32 through the function class
        const bool f32v = act_f32 && (cls == FW_RF_ACTION_DELTA || cls == FW_RF_ACTION_VALUE);
        const int fc = c.rew_fclass[i];
        if (fc == FW_FN_LINEAR) {
            if (f32v) val = (T)fminf(fmaxf(fabsf((float)val) / (float)c.rew_scaling[i], 0.f), (float)c.rew_maxv[i]);
            else val = clip(M<T>::fabs(val) / c.rew_scaling[i], (T)0, c.rew_maxv[i]);
        } else if (f32v) val = (T)(((float)val * (float)val) / (float)c.rew_scaling[i]);
        else val = val * val / c.rew_scaling[i];
        if (c.rew_shaping[i]) shp_t[fc] += val * c.rew_sign[i];
        else val_t[fc] += val * c.rew_sign[i];
    }
    T reward = 0;
#pragma unroll 1
    for (int t = 0; t < c.rew_nterms; ++t) {
        const int fc = c.term_fclass[t];
        const T prev = S.r[(RF_PREV_SHAPING + fc) * n + env];
        const bool has_prev = !M<T>::isnan(prev);
        T v;
        if (fc == FW_FN_EXPONENTIAL) {
            if (c.rew_potential) v = has_prev ? (T)-1 + M<T>::exp(val_t[fc] + (shp_t[fc] - prev)) : (T)-1 + M<T>::exp(val_t[fc]);
            else v = (T)-1 + M<T>::exp(val_t[fc] + shp_t[fc]);
        } else {
            v = val_t[fc];
            if (c.rew_potential) { if (has_prev) v += shp_t[fc] - prev; }
            else v += shp_t[fc];
        }
        S.r[(RF_PREV_SHAPING + fc) * n + env] = shp_t[fc];
        reward += c.term_weight[t] * v;
    }
    return reward;
}

// General observation layout (fixed_wing.py:1113-1262): obs_len rows, newest first; row i reads `.history[-i]` of the
// states, targets and errors, clamped to the start of the episode (then the row gets the `init_noise` offset
// U(-1,1) * dt, one draw per row, fixed_wing.py:1142-1145), action entries sum |diff| over a window of raw actions (or
// constrained commands when actions are not scaled) that ends i-1 steps back, or the backward-scaled actuator value
// while the episode is younger than the row.  `steps` = steps_count (0 at reset).  When `push` the rings are advanced
// with this step's values afterwards.
template <typename T>
__device__ __noinline__ void generic_observation(const DCfg<T>& c, const Soa<T>& S, int env, int steps, bool push,
                                                 const T (&cur)[14] /* 8 states, 3 targets, 3 errors */,
                                                 const T (&a_raw)[3], bool act_f32, const T (&cmd_in)[3],
                                                 const T (&actval)[3], unsigned long long episode, T* o,
                                                 const T* int_reset /* reset observation: integrator values (6), else null */,
                                                 const T* acur = nullptr /* attitude_angular: 3 rate targets, 3 rate errors */) {
    const int n = S.n, L = c.obs_len, ne = c.obs_n;
    T* r = S.r + env;
    T hist[4][14], gact[24], gcmd[24];
#pragma unroll 1
    for (int k = 0; k < 4; ++k)
        for (int q = 0; q < 14; ++q) hist[k][q] = r[(RF_HIST + k * 14 + q) * n];
#pragma unroll 1
    for (int k = 0; k < 24; ++k) { gact[k] = r[(RF_GACT + k) * n]; gcmd[k] = r[(RF_GCMD + k) * n]; }
    const int N = steps;                                     // number of actions so far, the current one included
    const int hist_len = steps + (push ? 1 : 0);             // len(history["error"]) the observation sees
#pragma unroll 1
    for (int row = 0; row < L; ++row) {
        const int i = 1 + row * c.obs_step;                  // range(1, length * step, step) (fixed_wing.py:1129-1138)
        int ie = i;
        T init_noise = 0;
        if (i > steps) {
            ie = steps + 1;
            if (L > 1) {
                T u;
                if (!M<T>::isnan(c.obs_init_noise)) u = c.obs_init_noise;
                else {
                    const uint4 rr = rng_block(env_seed(S, env), c.env_id_offset + env, episode, RNG_OBS_INIT, (uint32_t)(steps * 8 + (i - 1)));
                    u = (T)(2.0 * u53(rr.x, rr.y) - 1.0);
                }
                init_noise = u * c.dt;
            }
        }
#pragma unroll 1
        for (int k = 0; k < ne; ++k) {
            const int idx = c.obs_idx[k], kind = c.obs_kind[k];
            T val;
            if (kind == FW_OBS_TARGET_INT) {
                if (int_reset) val = int_reset[idx];
                else {
                    const int W = c.integration_window;
                    val = err_ring_sum<T>(S, env, idx, hist_len - W - ie, hist_len - ie, hist_len);
                    if (steps - ie < W) val += (T)(W - (steps - ie)) * r[e0_field(idx) * n];
                }
            } else if (kind != FW_OBS_ACTION && idx >= 3 && (kind == FW_OBS_TARGET_ABS || kind == FW_OBS_TARGET_REL)) {
                // a rate target (attitude_angular) or its error: current values, or the row of RF_AHIST
                const int q = (kind == FW_OBS_TARGET_ABS ? 0 : 3) + idx - 3;
                val = (ie == 1) ? acur[q] : r[(RF_AHIST + (ie - 2) * 6 + q) * n];
            } else if (kind != FW_OBS_ACTION) {
                const int q = (kind == FW_OBS_STATE) ? idx : (kind == FW_OBS_TARGET_ABS ? 8 + idx : 11 + idx);
                val = (ie == 1) ? cur[q] : hist[ie - 2][q];
            } else if (steps - ie < 0) {
                val = actval[idx];
                if (c.scale_actions)
                    val = (c.scale_high - c.scale_low) * (val - c.act_lo[idx]) / (c.act_hi[idx] - c.act_lo[idx]) + c.scale_low;
            } else {
                const bool f32 = c.scale_actions ? act_f32 : false;
                const T* ring = c.scale_actions ? gact : gcmd;
                const T curv = c.scale_actions ? a_raw[idx] : cmd_in[idx];
                const int hi = N - (ie - 1);
                int lo = N - c.obs_window[k] - ie + 1;
                if (lo < 0) lo = 0;
                float sacc = 0.f;
                for (int t = lo + 1; t < hi; ++t) {
                    const T newer = (t == N - 1) ? curv : ring[(N - 2 - t) * 3 + idx];
                    const T older = ring[(N - 2 - (t - 1)) * 3 + idx];
                    sacc += f32 ? fabsf((float)newer - (float)older) : (float)M<T>::fabs(newer - older);
                }
                val = (T)sacc;
            }
            val += init_noise;
            if (c.obs_normalize && c.obs_norm_flag[k]) { val -= c.obs_mean[k]; val /= c.obs_var[k]; }
            o[row * ne + k] = val;
        }
    }
    if (c.obs_noise_std > (T)0 || c.obs_noise_mean != (T)0) add_obs_noise<T>(c, env_seed(S, env), c.env_id_offset + env, episode, steps, o, L * ne);
    if (push && acur) {
#pragma unroll 1
        for (int k = 3; k >= 1; --k)
            for (int q = 0; q < 6; ++q) r[(RF_AHIST + k * 6 + q) * n] = r[(RF_AHIST + (k - 1) * 6 + q) * n];
        for (int q = 0; q < 6; ++q) r[(RF_AHIST + q) * n] = acur[q];
    }
    if (push) {
#pragma unroll 1
        for (int k = 3; k >= 1; --k)
            for (int q = 0; q < 14; ++q) r[(RF_HIST + k * 14 + q) * n] = hist[k - 1][q];
        for (int q = 0; q < 14; ++q) r[(RF_HIST + q) * n] = cur[q];
#pragma unroll 1
        for (int k = 23; k >= 3; --k) { r[(RF_GACT + k) * n] = gact[k - 3]; r[(RF_GCMD + k) * n] = gcmd[k - 3]; }
        for (int j = 0; j < 3; ++j) { r[(RF_GACT + j) * n] = a_raw[j]; r[(RF_GCMD + j) * n] = cmd_in[j]; }
    }
}

// Adds the integrator values of the ended episode to the precomputed reset observation of `env` (whose integrator
// entries were computed as 0): (0 - mean) / var + v / var for a normalised entry, every row alike.
template <typename T, typename SpareT>
__device__ __noinline__ void patch_integrator_reset_obs(const DCfg<T>& c, const SpareT& P, int env, int odim,
                                                        const T (&v)[6], float* obs, double* obs64) {
#pragma unroll 1
    for (int row = 0; row < c.obs_len; ++row)
#pragma unroll 1
        for (int k = 0; k < c.obs_n; ++k) {
            if (c.obs_kind[k] != FW_OBS_TARGET_INT) continue;
            double add = (double)v[c.obs_idx[k]];
            if (c.obs_normalize && c.obs_norm_flag[k]) add /= (double)c.obs_var[k];
            const size_t q = (size_t)env * odim + row * c.obs_n + k;
            const double val = P.obs64[q] + add;
            if (obs) obs[q] = (float)val;
            if (obs64) obs64[q] = val;
        }
}

// What the RHS reads, from the FW_NPARAM base parameters of one env (FwConfig order: mass 0, S_wing 5, b 6, c 7, S_prop 8,
// C_prop 9, k_motor 10, k_T_P 11, k_Omega 12, e 13, M 14, a_0 15, C_L_0 16 ...): the same expressions convert_cfg uses for
// the warp-uniform constants.  Jx..Jxz and the aspect ratio keep their construction-time values (pyfly.py:1086-1119).
template <typename T>
__device__ __forceinline__ void write_env_params(const DCfg<T>& c, const Soa<T>& S, int env, const double (&bp)[FW_NPARAM]) {
    const int n = S.n;
    T* p = S.par + env;
#pragma unroll 1
    for (int i = 0; i < FW_NPARAM; ++i) p[(size_t)i * n] = (T)bp[i];
    T* d = p + (size_t)FW_NPARAM * n;
    d[(size_t)PD_S_wing * n] = (T)bp[5]; d[(size_t)PD_b * n] = (T)bp[6]; d[(size_t)PD_c * n] = (T)bp[7];
    d[(size_t)PD_k_motor * n] = (T)bp[10]; d[(size_t)PD_k_T_P * n] = (T)bp[11]; d[(size_t)PD_k_Omega * n] = (T)bp[12];
    d[(size_t)PD_M_ * n] = (T)bp[14]; d[(size_t)PD_a_0 * n] = (T)bp[15];
#pragma unroll 1
    for (int k = 0; k < 19; ++k) d[(size_t)(PD_C_L_0 + k) * n] = (T)bp[16 + k];            // C_L_0 .. C_Y_delta_a (16 .. 34)
#pragma unroll 1
    for (int k = 0; k < 5; ++k) d[(size_t)(PD_C_l_0 + k) * n] = (T)bp[36 + k];             // C_l_0 .. C_l_delta_a
#pragma unroll 1
    for (int k = 0; k < 5; ++k) d[(size_t)(PD_C_n_0 + k) * n] = (T)bp[42 + k];             // C_n_0 .. C_n_delta_a
    d[(size_t)PD_mg * n] = (T)(bp[0] * (double)c.g_);
    d[(size_t)PD_inv_mass * n] = (T)(1.0 / bp[0]);
    d[(size_t)PD_prop_k * n] = (T)((double)c.half_rho * bp[8] * bp[9]);
    d[(size_t)PD_inv_pi_e_ar * n] = (T)(1.0 / (3.14159265358979323846 * bp[13] * (double)c.ar));
    d[(size_t)PD_exp_M_a0 * n] = (T)::exp(bp[14] * bp[15]);
}
static_assert(PD_C_Y_delta_a - PD_C_L_0 == 18 && PD_C_l_delta_a - PD_C_l_0 == 4 && PD_C_n_delta_a - PD_C_n_0 == 4, "parameter blocks");

// sample_simulator_parameters, "model" block (fixed_wing.py:758-800), same Philox stream as oracle sample_model_params
template <typename T>
__device__ __noinline__ void sample_env_params(const DCfg<T>& c, const Soa<T>& S, int env, unsigned long long seed,
                                               long long gid, unsigned long long episode) {
    const ResetCfg<T>& rc = *S.rc;
    double bp[FW_NPARAM];
#pragma unroll 1
    for (int i = 0; i < FW_NPARAM; ++i) {
        const double orig = rc.par_orig[i];
        double v = orig;
        if (rc.par_enabled[i] && orig != 0.0) {
            const uint4 r = rng_block(seed, gid, episode, RNG_MODEL, (uint32_t)i);
            const double u1 = u53(r.x, r.y), u2 = u53(r.z, r.w);
            if (rc.model_uniform) {
                const double lo = orig - rc.par_var[i], hi = orig + rc.par_var[i];
                v = lo + (hi - lo) * u1;
            } else {
                v = orig + rc.par_var[i] * (::sqrt(-2.0 * ::log(1.0 - u1)) * ::cos(6.283185307179586476925 * u2));
                if (!::isnan(rc.par_clip[i])) v = ::fmin(::fmax(v, orig - rc.par_clip[i]), orig + rc.par_clip[i]);
            }
        } else if (!rc.par_enabled[i]) {
            v = orig;
        }
        bp[i] = v;
    }
    write_env_params<T>(c, S, env, bp);
}

// FixedWingAircraft.reset -> PyFly.reset for one env; writes the full SoA row and the reset observation.
template <typename T>
__device__ void reset_env(const DCfg<T>& c, const Soa<T>& S, int env, const double* state_in, const double* target_in,
                          float* obs, double* obs64, bool live = false) {
    const int n = S.n;
    const unsigned long long episode = (unsigned long long)(uint32_t)S.i[IF_EPISODE * n + env] + 1ull;
    // `live`: S is the running env (fw_reset), not a precomputed next-episode row.  The reset observation's "integrator"
    // entries read the error history of the episode that is being replaced (fixed_wing.py:453-460, 1165-1180): taken
    // from the live ring before the row is rewritten; a precomputed row carries 0 there and head_kernel adds the value
    // when the row is consumed.  An env that was never reset has no history: error * window, filled in below.
    T int_reset[6] = {0, 0, 0, 0, 0, 0};
    const bool int_none = live && episode == 1ull;
    if (live && c.obs_generic && c.obs_has_int && !int_none) {
        const int steps_prev = S.i[IF_STEPS * n + env];
        const int failed = steps_prev > 0 && S.ep_term[env] >= FW_TERM_OMEGA_P;        // a failed step appends no error entry
        const int len = steps_prev + (failed ? 0 : 1);
#pragma unroll 1
        for (int k = 0; k < (c.ang_on ? 6 : 3); ++k) int_reset[k] = integrator_reset_value<T>(c, S, env, k, len, S.r[e0_field(k) * n + env]);
    }
    const long long gid = c.env_id_offset + env;
    const ResetCfg<T>& rc = *S.rc;
    const unsigned long long seed = rc.seed;
    // the episode keeps this key until its next reset (noise_sample / add_obs_noise / resampling read it back)
    S.i[IF_SEED_LO * n + env] = (int32_t)(uint32_t)seed;
    S.i[IF_SEED_HI * n + env] = (int32_t)(uint32_t)(seed >> 32);
    T s12[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const double inj = state_in ? state_in[(size_t)env * FW_NSTATE_INJECT + i] : CUDART_NAN;
        s12[i] = ::isnan(inj) ? rc.init_lo[i] + (rc.init_hi[i] - rc.init_lo[i]) * rng_uniform<T>(seed, gid, episode, RNG_RESET, i)
                              : (T)inj;
    }
    T a6[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double inj = state_in ? state_in[(size_t)env * FW_NSTATE_INJECT + 12 + i] : CUDART_NAN;
        a6[i] = ::isnan(inj) ? (T)0 : (T)inj;
    }
    a6[0] = clip(a6[0], c.elevon_min, c.elevon_max);
    a6[1] = clip(a6[1], c.elevon_min, c.elevon_max);
    a6[2] = clip(a6[2], c.throttle_min, c.throttle_max);
    a6[3] = clip(a6[3], -c.elevon_dot_max, c.elevon_dot_max);
    a6[4] = clip(a6[4], -c.elevon_dot_max, c.elevon_dot_max);
    T wind[3];
    {
        double w0 = state_in ? state_in[(size_t)env * FW_NSTATE_INJECT + 18] : CUDART_NAN;
        double w1 = state_in ? state_in[(size_t)env * FW_NSTATE_INJECT + 19] : CUDART_NAN;
        double w2 = state_in ? state_in[(size_t)env * FW_NSTATE_INJECT + 20] : CUDART_NAN;
        if (!::isnan(w0) && !::isnan(w1) && !::isnan(w2)) { wind[0] = (T)w0; wind[1] = (T)w1; wind[2] = (T)w2; }
        else {   // Wind.reset (pyfly.py:815-823)
            const T mag = rc.wind_mag_min + (rc.wind_mag_max - rc.wind_mag_min) * rng_uniform<T>(seed, gid, episode, RNG_RESET, 12);
            const T w_n = -mag + (mag - -mag) * rng_uniform<T>(seed, gid, episode, RNG_RESET, 13);
            const T w_e_max = M<T>::sqrt(mag * mag - w_n * w_n);
            const T w_e = -w_e_max + (w_e_max - -w_e_max) * rng_uniform<T>(seed, gid, episode, RNG_RESET, 14);
            wind[0] = w_n; wind[1] = w_e; wind[2] = M<T>::sqrt(mag * mag - w_n * w_n - w_e * w_e);
        }
    }
    const T roll = s12[0], pitch = s12[1], yaw = s12[2];
    // turbulence: x_0 = 0, u_0 = first noise sample; column 0 of the reference's tables is C.0 + D u_0
    T fx[12], fu[4], tl[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 12; ++i) fx[i] = 0;
    if (c.turbulence) { noise_sample(c, S, env, episode, 0, fu); turb_eval(c, fx, fu, tl, ta); }
    else { fu[0] = fu[1] = fu[2] = fu[3] = 0; }
    // Va, alpha, beta from Euler + vel (pyfly.py:1297-1306)
    T wb[3];
    rot_euler_apply(roll, pitch, yaw, wind, wb);
    const T a0 = s12[9] - (wb[0] + tl[0]), a1 = s12[10] - (wb[1] + tl[1]), a2 = s12[11] - (wb[2] + tl[2]);
    T Va = M<T>::sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    const T alpha = M<T>::atan2(a2, a0), beta = M<T>::asin(a1 / Va);
    if (Va < c.va_value_min) Va = c.va_value_min;
    // AttitudeQuaternion._from_euler_angles (pyfly.py:714-737)
    T sps, cps, sth, cth, sph, cph;
    M<T>::sincos(yaw / (T)2, &sps, &cps);
    M<T>::sincos(pitch / (T)2, &sth, &cth);
    M<T>::sincos(roll / (T)2, &sph, &cph);
    T y[FW_NY];
    y[0] = cps * cth * cph + sps * sth * sph;
    y[1] = cps * cth * sph - sps * sth * cph;
    y[2] = cps * sth * cph + sps * cth * sph;
    y[3] = sps * cth * cph - cps * sth * sph;
#pragma unroll
    for (int i = 0; i < 9; ++i) y[4 + i] = s12[3 + i];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[13 + i] = a6[i];
    if (S.par) sample_env_params<T>(c, S, env, seed, gid, episode);     // fixed_wing.py:442, before sample_target
    // targets (fixed_wing.py:443-450)
    T u12[12], tgt[3], tp[15];
    int tcls[3];
#pragma unroll
    for (int k = 0; k < 15; ++k) tp[k] = 0;
    target_draws<T>(c, seed, gid, episode, RNG_RESET, 8u, u12);       // blocks 8..13 of the reset stream (0..7: state, wind)
    sample_target<T>(c, rc, roll, pitch, Va, 0, u12, tgt, tcls, tp);
    // attitude_angular: the rate targets derive from the SAMPLED attitude targets, before injected ones override them
    T atgt[3] = {0, 0, 0};
    if (c.ang_on) angular_targets<T>(c, roll, pitch, err_roll(tgt[0], roll), tgt[1] - pitch, true, atgt);
    if (target_in) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double tv = target_in[(size_t)env * 3 + k];
            if (!::isnan(tv)) {
                if (tcls[k] != FW_TGT_CONSTANT && tcls[k] != FW_TGT_COMPENSATE) tcls[k] = FW_TGT_CONSTANT;
                tgt[k] = (T)tv;
            }
        }
    }
    // ---- store ----
    T* r = S.r + env;
    int32_t* ii = S.i + env;
#pragma unroll
    for (int i = 0; i < FW_NY; ++i) r[(RF_Y + i) * n] = y[i];
    r[RF_ROLL * n] = roll; r[RF_PITCH * n] = pitch; r[RF_VA * n] = Va; r[RF_ALPHA * n] = alpha; r[RF_BETA * n] = beta;
#pragma unroll
    for (int k = 0; k < 3; ++k) { r[(RF_WIND + k) * n] = wind[k]; r[(RF_TGT + k) * n] = tgt[k]; }
#pragma unroll
    for (int i = 0; i < 12; ++i) r[(RF_FX + i) * n] = fx[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[(RF_FU + i) * n] = fu[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) { r[(RF_ACT_RING + i) * n] = 0; r[(RF_CMD_RING + i) * n] = 0; }
    r[RF_CV_SUM * n] = 0;
    const T e[3] = {err_roll(tgt[0], roll), tgt[1] - pitch, tgt[2] - Va};
    int gbits[4];
    gbits[3] = 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r[(RF_E0 + k) * n] = e[k]; r[(RF_ESUM + k) * n] = e[k]; r[(RF_EABS + k) * n] = M<T>::fabs(e[k]);
        r[(RF_EMIN + k) * n] = e[k]; r[(RF_EMAX + k) * n] = e[k]; r[(RF_EPREV + k) * n] = M<T>::fabs(e[k]);
        S.err_ring[(size_t)(0 * 3 + k) * n + env] = e[k];
        gbits[k] = M<T>::fabs(e[k]) <= c.tgt_bound[k];
        gbits[3] &= gbits[k];
    }
    T ea[3] = {0, 0, 0};
    if (c.ang_on) {
#pragma unroll 1
        for (int a = 0; a < 3; ++a) {
            ea[a] = atgt[a] - y[4 + a];
            const int ga = M<T>::fabs(ea[a]) <= c.ang_bound[a];
            gbits[3] &= ga;
            r[(RF_ATGT + a) * n] = atgt[a];
            r[(RF_AE0 + a) * n] = ea[a]; r[(RF_AESUM + a) * n] = ea[a]; r[(RF_AEABS + a) * n] = M<T>::fabs(ea[a]);
            r[(RF_AEMIN + a) * n] = ea[a]; r[(RF_AEMAX + a) * n] = ea[a]; r[(RF_AEPREV + a) * n] = M<T>::fabs(ea[a]);
            S.err_ring_a[(size_t)a * n + env] = ea[a];
            ii[(IF_AGOAL_RING + 4 * a) * n] = ga;
            ii[(IF_AGOAL_RING + 4 * a + 1) * n] = 0; ii[(IF_AGOAL_RING + 4 * a + 2) * n] = 0; ii[(IF_AGOAL_RING + 4 * a + 3) * n] = 0;
            ii[(IF_AGOAL_CNT + a) * n] = ga; ii[(IF_AGOAL_TOTAL + a) * n] = ga;
            ii[(IF_ASETTLE + a) * n] = (c.streak_req == 1 && (double)ga >= (double)c.streak_fraction) ? 0 : -1;
            ii[(IF_ARISE_LO + a) * n] = -1; ii[(IF_ARISE_HI + a) * n] = -1;
        }
        for (int q = 0; q < 3; ++q) { r[(RF_AHIST + q) * n] = atgt[q]; r[(RF_AHIST + 3 + q) * n] = ea[q]; }
        for (int q = 6; q < 24; ++q) r[(RF_AHIST + q) * n] = 0;
    }
    r[RF_EP_RET * n] = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) r[(RF_PREV_SHAPING + k) * n] = M<T>::nan();
    if (c.tgt_moving) {
#pragma unroll
        for (int k = 0; k < 15; ++k) r[(RF_TPROP + k) * n] = tp[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) ii[(IF_TCLS + k) * n] = tcls[k];
    }
    ii[IF_STEPS * n] = 0; ii[IF_STEPS_TGT * n] = 0; ii[IF_EPISODE * n] = (int32_t)episode; ii[IF_SIM_STEP * n] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ii[(IF_GOAL_RING + 4 * k) * n] = gbits[k];
        ii[(IF_GOAL_RING + 4 * k + 1) * n] = 0; ii[(IF_GOAL_RING + 4 * k + 2) * n] = 0; ii[(IF_GOAL_RING + 4 * k + 3) * n] = 0;
        ii[(IF_GOAL_CNT + k) * n] = gbits[k];
        ii[(IF_GOAL_TOTAL + k) * n] = gbits[k];
        ii[(IF_SETTLE + k) * n] = (c.streak_req == 1 && (double)gbits[k] >= (double)c.streak_fraction) ? 0 : -1;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { ii[(IF_RISE_LO + k) * n] = -1; ii[(IF_RISE_HI + k) * n] = -1; }
    ii[IF_NFEV * n] = 0; ii[IF_NATT * n] = 0; ii[IF_ACT_F32 * n] = 0;
    // reset observation: action entries are the actuator values mapped backward (fixed_wing.py:1188-1196);
    // elevator/aileron read 0 right after reset (disabled ControlVariable.reset, pyfly.py:359-363)
    T o[FW_NOBS] = {roll, pitch, Va, y[4], y[5], y[6], tgt[0], tgt[1], tgt[2], alpha, beta, 0, 0, 0};
    const T av[3] = {(T)0, (T)0, a6[2]};
#pragma unroll
    for (int j = 0; j < 3; ++j)
        o[11 + j] = c.scale_actions ? (c.scale_high - c.scale_low) * (av[j] - c.act_lo[j]) / (c.act_hi[j] - c.act_lo[j]) + c.scale_low
                                    : av[j];
    if (c.obs_generic) {
        // history rings start with the reset entry (Variable.reset: history = [value]); the reset observation has every
        // row clamped to the current values
        const T cur[14] = {roll, pitch, Va, y[4], y[5], y[6], alpha, beta, tgt[0], tgt[1], tgt[2], e[0], e[1], e[2]};
        for (int q = 0; q < 14; ++q) r[(RF_HIST + q) * n] = cur[q];
        for (int k = 14; k < 56; ++k) r[(RF_HIST + k) * n] = 0;
        for (int k = 0; k < 24; ++k) { r[(RF_GACT + k) * n] = 0; r[(RF_GCMD + k) * n] = 0; }
        T og[FW_NOBS_MAX];
        const T zero3[3] = {0, 0, 0};
        if (int_none)
            for (int k = 0; k < 3; ++k) { int_reset[k] = e[k] * (T)c.integration_window; int_reset[3 + k] = ea[k] * (T)c.integration_window; }
        const T acur[6] = {atgt[0], atgt[1], atgt[2], ea[0], ea[1], ea[2]};
        generic_observation<T>(c, S, env, 0, false, cur, zero3, false, zero3, av, episode, og, int_reset, c.ang_on ? acur : nullptr);
        write_obs(og, obs_dim(c), env, obs, obs64);
        return;
    }
    if (c.obs_noise_std > (T)0 || c.obs_noise_mean != (T)0) add_obs_noise<T>(c, seed, gid, episode, 0, o, FW_NOBS);
    write_obs(o, FW_NOBS, env, obs, obs64);
}

// ---------------- waypoint head: FixedWingAircraft_simple (magpie/magpy/simple_train.py:197-702) ----------------

// simulator.reset(state=waypoint) + goal of the leg (simple_train.py:357-363 -> pyfly.py:1262-1311): position, attitude,
// velocity and wind from the start waypoint; omega from the row or, if NaN, uniform in the pyfly init range; actuators at
// their (0, 0) init range; turbulence restarted on a fresh noise segment.  Writes the SoA row of the simulator.
template <typename T>
__device__ __noinline__ void wp_start_leg(const DCfg<T>& c, const Soa<T>& S, int env, int wp_pos) {
    const int n = S.n;
    T* r = S.r + env;
    int32_t* ii = S.i + env;
    const double* row = S.wp_tasks + ((size_t)S.wp_task_of_env[env] * S.wp_len + wp_pos) * FW_WP_ROW;
    const unsigned long long episode = (unsigned long long)(uint32_t)ii[IF_EPISODE * n] + 1ull;
    const long long gid = c.env_id_offset + env;
    const ResetCfg<T>& rc = *S.rc;
    const unsigned long long seed = rc.seed;
    ii[IF_SEED_LO * n] = (int32_t)(uint32_t)seed;
    ii[IF_SEED_HI * n] = (int32_t)(uint32_t)(seed >> 32);
    const T roll = (T)row[3], pitch = (T)row[4], yaw = (T)row[5];
    T om[3], wind[3], vel[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        om[i] = ::isnan(row[12 + i]) ? rc.init_lo[3 + i] + (rc.init_hi[3 + i] - rc.init_lo[3 + i]) * rng_uniform<T>(seed, gid, episode, RNG_RESET, 3 + i)
                                     : (T)row[12 + i];
        wind[i] = (T)row[9 + i];
        vel[i] = (T)row[6 + i];
    }
    T fx[12], fu[4], tl[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 12; ++i) fx[i] = 0;
    if (c.turbulence) { noise_sample(c, S, env, episode, 0, fu); turb_eval(c, fx, fu, tl, ta); }
    else { fu[0] = fu[1] = fu[2] = fu[3] = 0; }
    T wb[3];
    rot_euler_apply(roll, pitch, yaw, wind, wb);
    const T a0 = vel[0] - (wb[0] + tl[0]), a1 = vel[1] - (wb[1] + tl[1]), a2 = vel[2] - (wb[2] + tl[2]);
    T Va = M<T>::sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    const T alpha = M<T>::atan2(a2, a0), beta = M<T>::asin(a1 / Va);
    if (Va < c.va_value_min) Va = c.va_value_min;
    T sps, cps, sth, cth, sph, cph;
    M<T>::sincos(yaw / (T)2, &sps, &cps);
    M<T>::sincos(pitch / (T)2, &sth, &cth);
    M<T>::sincos(roll / (T)2, &sph, &cph);
    r[(RF_Y + 0) * n] = cps * cth * cph + sps * sth * sph;
    r[(RF_Y + 1) * n] = cps * cth * sph - sps * sth * cph;
    r[(RF_Y + 2) * n] = cps * sth * cph + sps * cth * sph;
    r[(RF_Y + 3) * n] = sps * cth * cph - cps * sth * sph;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r[(RF_Y + 4 + i) * n] = om[i];
        r[(RF_Y + 7 + i) * n] = (T)row[i];
        r[(RF_Y + 10 + i) * n] = vel[i];
        r[(RF_WIND + i) * n] = wind[i];
        r[(RF_TGT + i) * n] = (T)row[FW_WP_ROW + i];         // goal = position of the next waypoint
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) r[(RF_Y + 13 + i) * n] = 0;
    r[RF_ROLL * n] = roll; r[RF_PITCH * n] = pitch; r[RF_VA * n] = Va; r[RF_ALPHA * n] = alpha; r[RF_BETA * n] = beta;
#pragma unroll
    for (int i = 0; i < 12; ++i) r[(RF_FX + i) * n] = fx[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[(RF_FU + i) * n] = fu[i];
    ii[IF_EPISODE * n] = (int32_t)episode;
    ii[IF_SIM_STEP * n] = 0;
    ii[IF_WP_POS * n] = wp_pos;
    // IF_STEPS of the simulator-level "fresh reset" quirk (elevator = aileron = 0 at the first RHS call) is tracked by
    // IF_STEPS_TGT here: steps since the last simulator reset
    ii[IF_STEPS_TGT * n] = 0;
}

template <typename T>
__device__ __forceinline__ void wp_observation(const Soa<T>& S, int env, T (&o)[FW_NOBS_WAYPOINT]) {
    const int n = S.n;
    const T* r = S.r + env;
    o[0] = r[RF_ROLL * n]; o[1] = r[RF_PITCH * n]; o[2] = r[RF_VA * n];
    o[3] = r[(RF_Y + 4) * n]; o[4] = r[(RF_Y + 5) * n]; o[5] = r[(RF_Y + 6) * n];
    o[6] = r[(RF_Y + 14) * n]; o[7] = r[(RF_Y + 13) * n]; o[8] = r[(RF_Y + 15) * n];    // elevon_left, elevon_right, throttle
    o[9] = r[(RF_Y + 7) * n]; o[10] = r[(RF_Y + 8) * n]; o[11] = r[(RF_Y + 9) * n];
}

// reset (simple_train.py:385-408): steps 0, first leg of the env's task
template <typename T>
__device__ void wp_reset_env(const DCfg<T>& c, const Soa<T>& S, int env, float* obs, double* obs64) {
    const int n = S.n;
    S.i[IF_STEPS * n + env] = 0;
    S.r[RF_EP_RET * n + env] = 0;
    wp_start_leg<T>(c, S, env, 0);
    T o[FW_NOBS_WAYPOINT];
    wp_observation<T>(S, env, o);
    write_obs(o, FW_NOBS_WAYPOINT, env, obs, obs64);
}

}  // namespace fw
