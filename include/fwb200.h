/* fwb200.h — C ABI of the B200-native batched fixed-wing simulator (libfwb200.so).
 *
 * Drop-in boundary for the hot path named by BASELINE.json:north_star: everything below the SB3 `VecEnv`
 * contract of the reference (magpie/libs/stable-baselines3/stable_baselines3/common/vec_env/base_vec_env.py:48-224)
 * for the `FixedWingAircraft` env (magpie/libs/fixed-wing-gym/gym_fixed_wing/fixed_wing.py:13-1736) on top of the
 * `PyFly` simulator (magpie/libs/pyfly/pyfly/pyfly.py:1030-1881, dryden.py:6-261), plus the GAE pass of the
 * forked SB3 RolloutBuffer (stable_baselines3/common/buffers.py:304-333).
 *
 * Conventions
 *  - Plain C: pointers, sizes, PODs.  No torch / C++ types.
 *  - Every `*_dev` pointer is a DEVICE pointer owned by the caller (e.g. a torch CUDA tensor's data_ptr()).
 *    Env state (structure-of-arrays in HBM) is owned by the handle.
 *  - All compute entry points are asynchronous on the `stream` argument (a cudaStream_t passed as void*;
 *    NULL = legacy default stream).  Nothing here synchronises except fw_create / fw_destroy / the host getters
 *    (and a step while fw_set_profiling is on).  The handle owns one more, non-blocking stream on which it
 *    recomputes precomputed reset rows after an auto-reset step; the next fw_step / fw_reset orders itself
 *    behind that work through events (fw_join does so explicitly).
 *  - Return value: 0 on success, negative FW_E* on USAGE errors only.  Simulation failures (constraint violations,
 *    pyfly's ConstraintException pyfly.py:11-16) are DATA: they set done=1 and term_code, never an error.
 *  - One handle per GPU; a handle is not thread-safe.  There is no CPU fallback: without a CUDA device
 *    fw_create returns FW_ENODEVICE.
 */
#ifndef FWB200_H
#define FWB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FW_ABI_VERSION 10

#define FW_NY 19        /* ODE state: quat[4] omega[3] pos[3] vel[3] act_value[3] act_rate[3]  (pyfly.py:1372-1389) */
#define FW_NOBS 14      /* default observation vector (fixed_wing_config.json "observation.states")               */
#define FW_OBS_ENTRIES_MAX 16  /* entries per observation row                                                     */
#define FW_OBS_LEN_MAX 5       /* observation.length (rows of history)                                            */
#define FW_NOBS_MAX (FW_OBS_ENTRIES_MAX * FW_OBS_LEN_MAX)
#define FW_REW_FACTORS_MAX 12
#define FW_NPARAM 48    /* aircraft parameters mass .. C_n_delta_r of FwConfig, in that order                             */
#define FW_NACT 3       /* elevator, aileron, throttle (fixed_wing_config.json "action.states")                   */
#define FW_NSTATE_INJECT 21 /* roll pitch yaw p q r pn pe pd u v w | er el thr | er_dot el_dot thr_dot | wind n e d */
#define FW_NMETRIC 28   /* see FwMetricIndex                                                                      */
#define FW_ACT_WINDOW_MAX 8
#define FW_END_ERR_WINDOW 50   /* fixed_wing.py:1666 */
#define FW_NMETRIC_ANG 24      /* attitude_angular: avg_error, total_error, end_error, rise_time, overshoot, success,
                                  settling_time, success_time_frac x (omega_p, omega_q, omega_r), metric-major */

enum FwError {
    FW_OK = 0,
    FW_EINVAL = -1,      /* bad argument / shape / config */
    FW_ENODEVICE = -2,   /* no CUDA device or wrong architecture */
    FW_ECUDA = -3,       /* CUDA runtime error (see fw_last_error) */
    FW_ENOMEM = -4
};

/* term_code[n] written on done (fixed_wing.py:521,550,595): */
enum FwTermCode {
    FW_TERM_NONE = 0,
    FW_TERM_STEPS = 1,
    FW_TERM_SUCCESS = 2,
    FW_TERM_OMEGA_P = 10,   /* ConstraintException variable (pyfly.py:121-125) */
    FW_TERM_OMEGA_Q = 11,
    FW_TERM_OMEGA_R = 12,
    FW_TERM_VA = 13
};

enum FwIntegrator {
    FW_INT_RK45_SCIPY = 0,  /* bit-for-logic replica of scipy solve_ivp RK45 as called at pyfly.py:1393-1395 */
    FW_INT_RK4_FIXED = 1    /* fixed-step classical RK4 x substeps (throughput mode, graded at its own tolerance) */
};

enum FwPrecision { FW_F64 = 0, FW_F32 = 1 };

/* which env head runs on top of the simulator */
enum FwEnvKind {
    FW_ENV_ATTITUDE = 0,   /* FixedWingAircraft (fixed-wing-gym/gym_fixed_wing/fixed_wing.py) */
    FW_ENV_WAYPOINT = 1    /* FixedWingAircraft_simple (magpie/magpy/simple_train.py:197-702): fly through a chain of waypoints */
};
#define FW_WP_ROW 15          /* waypoint row: position n e d | roll pitch yaw | velocity u v w | wind n e d | omega p q r (NaN = sample) */
#define FW_NOBS_WAYPOINT 12   /* roll pitch Va p q r elevon_left elevon_right throttle position n e d (simple_train.py:248-261) */

/* observation entry kinds (fixed_wing.py:1149-1234) and the state indices an entry of kind STATE may name */
/* general reward engine (fixed_wing.py:941-1111): factor classes/types and function classes */
enum FwRewFactor { FW_RF_STATE_ERROR = 0, FW_RF_STATE_VALUE, FW_RF_ACTION_VALUE, FW_RF_ACTION_DELTA, FW_RF_ACTION_BOUND,
                   FW_RF_SUCCESS, FW_RF_STEP, FW_RF_GOAL_PER_STATE, FW_RF_GOAL_ALL,
                   FW_RF_STATE_INT_ERROR /* "int_error": windowed error integral, fixed_wing.py:1003-1012 */ };
enum FwRewFunction { FW_FN_LINEAR = 0, FW_FN_EXPONENTIAL = 1, FW_FN_QUADRATIC = 2 };

enum FwObsKind { FW_OBS_STATE = 0, FW_OBS_TARGET_ABS = 1, FW_OBS_TARGET_REL = 2, FW_OBS_ACTION = 3,
                 FW_OBS_TARGET_INT = 4 /* value "integrator": windowed error integral, fixed_wing.py:1165-1180 */ };
enum FwObsState { FW_S_ROLL = 0, FW_S_PITCH, FW_S_VA, FW_S_OMEGA_P, FW_S_OMEGA_Q, FW_S_OMEGA_R, FW_S_ALPHA, FW_S_BETA };

enum FwTargetClass { FW_TGT_CONSTANT = 0, FW_TGT_COMPENSATE = 1, FW_TGT_LINEAR = 2, FW_TGT_SINUSOIDAL = 3 };   /* fixed_wing.py:1375-1452 */
enum FwOnSuccess { FW_SUCCESS_NONE = 0, FW_SUCCESS_DONE = 1, FW_SUCCESS_NEW = 2 }; /* fixed_wing.py:548-553 */

/* Layout of the per-env metric row written when an episode ends (fixed_wing.py:1644-1736). NaN = numpy nan. */
enum FwMetricIndex {
    FW_M_RISE_TIME = 0,        /* [3] roll pitch Va            */
    FW_M_SETTLING_TIME = 3,    /* [4] roll pitch Va all        */
    FW_M_OVERSHOOT = 7,        /* [3]                          */
    FW_M_TOTAL_ERROR = 10,     /* [3]                          */
    FW_M_AVG_ERROR = 13,       /* [3]                          */
    FW_M_CONTROL_VARIATION = 16, /* [1]                        */
    FW_M_SUCCESS = 17,         /* [4] 0/1                      */
    FW_M_SUCCESS_TIME_FRAC = 21, /* [4]                        */
    FW_M_END_ERROR = 25        /* [3]                          */
};

/* One Dryden shaping filter in lsim form (dryden.py:22-39 -> scipy.signal.lsim):
 *   x_i = Ad x_{i-1} + Bd0 u_{i-1} + Bd1 u_i ;  y_i = C x_i + D u_i ;  x_0 = 0.
 * Matrices are computed on the host at init from the reference's (mis-ordered) parameters. */
typedef struct FwFilter {
    int32_t order;        /* 1..3 */
    int32_t noise_row;    /* which of the 4 white-noise rows drives it (dryden.py:238-252: u0 v1 w2 p3 q1 r2) */
    double Ad[9];         /* row-major order x order */
    double Bd0[3];
    double Bd1[3];
    double C[3];
    double D;
    double Ablk[9];       /* expm(A^T * turb_block_len * dt_dryden): lsim restarts every block from t = T[0] > 0 by
                             propagating the carried state over [0, T[0]] with zero input (scipy lsim; dryden.py:30-36) */
} FwFilter;

typedef struct FwConfig {
    int32_t abi_version;      /* FW_ABI_VERSION */
    int32_t precision;        /* FwPrecision: arithmetic type of the step kernel */
    int32_t integrator;       /* FwIntegrator */
    int32_t rk4_substeps;     /* for FW_INT_RK4_FIXED */
    double rtol, atol;        /* for FW_INT_RK45_SCIPY (scipy defaults 1e-3 / 1e-6) */

    /* ---- aircraft parameters (x8_param.mat, pyfly.py:1076-1119) ---- */
    double mass, Jx, Jy, Jz, Jxz, S_wing, b, c, S_prop, C_prop, k_motor, k_T_P, k_Omega, e_oswald, M, a_0;
    double C_L_0, C_L_alpha, C_L_q, C_L_delta_e;
    double C_D_p, C_D_q, C_D_beta1, C_D_beta2, C_D_delta_e;
    double C_m_0, C_m_alpha, C_m_q, C_m_delta_e, C_m_fp;
    double C_Y_0, C_Y_beta, C_Y_p, C_Y_r, C_Y_delta_a, C_Y_delta_r;
    double C_l_0, C_l_beta, C_l_p, C_l_r, C_l_delta_a, C_l_delta_r;
    double C_n_0, C_n_beta, C_n_p, C_n_r, C_n_delta_a, C_n_delta_r;

    /* ---- simulator (pyfly_config.json + the gym config's "simulator.states" overrides) ---- */
    double dt, rho, g;
    double elevon_min, elevon_max;        /* value clip, rad (pyfly_config.json elevon_*: -30..35 deg) */
    double elevon_dot_max;                /* rate clip 3.4907 (NOT degree-converted: SURVEY App. E-6) */
    double elevon_omega0, elevon_zeta;    /* 2nd-order actuator (pyfly.py:299-304) */
    double throttle_min, throttle_max, throttle_tau; /* 1st-order actuator (pyfly.py:296-298) */
    double omega_con_min[3], omega_con_max[3];   /* ConstraintException limits on p,q,r (rad/s) */
    double va_value_min;                  /* Va clip floor 1e-6 */
    double va_con_max;                    /* Va constraint (70); <=0 disables */
    double init_lo[12], init_hi[12];      /* uniform init ranges: roll pitch yaw p q r pn pe pd u v w (pyfly.py:94) */
    double wind_mag_min, wind_mag_max;    /* steady wind sampling (pyfly.py:816-823) */
    int32_t turbulence;                   /* 0/1 */
    int32_t _pad0;
    double turb_noise_scale;              /* sqrt(pi/dt_dryden) (dryden.py:172,187) */
    FwFilter filt[6];                     /* H_u H_v H_w H_p H_q H_r */

    /* ---- gym env (fixed_wing_config.json) ---- */
    int32_t steps_max;
    int32_t scale_actions;                /* action.scale_space */
    double scale_low, scale_high;         /* action.scale_low/high (-1, 1) */
    double act_lo[3], act_hi[3];          /* actuator ranges used by linear_action_scaling (fixed_wing.py:250-251) */
    int32_t has_action_bounds;            /* action.bounds_multiplier present */
    int32_t _pad1;
    double action_bounds_min[3], action_bounds_max[3];  /* fixed_wing.py:264-276 */
    double tgt_low[3], tgt_high[3];       /* roll pitch Va, radians where applicable */
    double tgt_delta[3];                  /* NaN = no delta */
    double tgt_bound[3];                  /* goal bounds */
    int32_t tgt_class[3];                 /* FwTargetClass */
    int32_t tgt_radians[3];               /* convert_to_radians of the target state (slope / amplitude are converted AFTER sampling) */
    double tgt_slope_low[3], tgt_slope_high[3];          /* class linear (fixed_wing.py:699-707), per second */
    double tgt_amp_low[3], tgt_amp_high[3];              /* class sinusoidal (fixed_wing.py:710-727) */
    double tgt_period_low[3], tgt_period_high[3];        /* in steps (250, 500 by default) */
    double rng_u_override;                /* finite: every uniform draw of target sampling returns this value instead
                                             of the Philox stream (parity tests against a patched reference RNG) */
    int32_t on_success;                   /* FwOnSuccess */
    int32_t streak_req;                   /* target.success_streak_req (<=128) */
    int32_t resample_every;
    double streak_fraction;
    double rew_err_scaling[3], rew_err_max[3];   /* linear error factors; max = +inf when absent */
    double rew_delta_scaling, rew_delta_max;     /* action-delta factor; scaling<=0 disables */
    double rew_bound_scaling, rew_bound_max;     /* action-bound factor; scaling<=0 disables */
    int32_t rew_delta_window;             /* 5 */
    int32_t obs_act_window;               /* 5 (observation "window_size") */
    int32_t step_fail_timesteps;          /* reward.step_fail == "timesteps" */
    int32_t _pad2;
    double step_fail_value;
    double rise_low, rise_high;           /* rise_time metric limits (0.1, 0.9) */
    double obs_noise_mean, obs_noise_std; /* observation.noise {mean, var}: every entry += N(mean, var) where numpy's
                                             `scale=var` makes "var" a standard deviation (fixed_wing.py:1246-1247);
                                             std <= 0 and mean == 0 disables (the reference's default) */

    /* ---- general reward engine; rew_generic == 0 selects the default factor family above ---- */
    int32_t rew_generic, rew_n, rew_potential, rew_nterms;
    int32_t rew_class[FW_REW_FACTORS_MAX];      /* FwRewFactor */
    int32_t rew_idx[FW_REW_FACTORS_MAX];        /* target index (error) or FwObsState index (value) */
    int32_t rew_fclass[FW_REW_FACTORS_MAX];     /* FwRewFunction */
    int32_t rew_shaping[FW_REW_FACTORS_MAX], rew_window[FW_REW_FACTORS_MAX], rew_value_timesteps[FW_REW_FACTORS_MAX];
    double rew_scaling[FW_REW_FACTORS_MAX], rew_maxv[FW_REW_FACTORS_MAX], rew_sign[FW_REW_FACTORS_MAX], rew_value[FW_REW_FACTORS_MAX];
    int32_t term_fclass[4];                     /* terms in config order (3 used) */
    double term_weight[4];

    /* ---- general observation layout (fixed_wing.py:1113-1262): obs_len rows (history, newest first) of obs_n entries.
     * obs_generic == 0 selects the default 14-vector fast path and ignores the arrays below. ---- */
    int32_t obs_generic, obs_len, obs_n, obs_normalize;
    int32_t obs_kind[FW_OBS_ENTRIES_MAX], obs_idx[FW_OBS_ENTRIES_MAX], obs_window[FW_OBS_ENTRIES_MAX];
    int32_t obs_norm_flag[FW_OBS_ENTRIES_MAX];
    double obs_mean[FW_OBS_ENTRIES_MAX], obs_var[FW_OBS_ENTRIES_MAX];
    double obs_init_noise;                /* rows older than the episode get += U(-1,1)*dt (fixed_wing.py:1142-1145);
                                             a finite value here replaces the Philox draw (parity tests) */

    /* ---- env head selection and the waypoint head's constants (simple_train.py:236-300) ---- */
    int32_t env_kind;                     /* FwEnvKind */
    int32_t turb_block_len;               /* turbulence_sim_length: pyfly re-simulates the filters in blocks of this many
                                             samples; the state carried into block m is multiplied by Ablk^m (pyfly.py:870-871,
                                             dryden.py:193-261) */
    double wp_goal_bound[3];              /* goal box on position n e d (0.5 m) */
    double wp_rew_range[3];               /* reward = exp(-sum |err_k| / range_k) (6 m) */

    /* ---- counter-based RNG (Philox4x32-10) for auto-reset and turbulence noise ---- */
    uint64_t seed;
    int64_t env_id_offset;                /* global id of env 0 of this handle (sharding: rank*n_envs) */

    /* ---- per-episode aircraft-parameter randomisation: the gym config's "simulator.model" block
     * (FixedWingAircraft.sample_simulator_parameters, fixed_wing.py:748-813).  Index i = position in the aircraft
     * parameter block above (mass = 0 ... C_n_delta_r = 47).  At every reset an enabled parameter whose original value
     * is not 0 is re-drawn around par_orig[i]: N(orig, par_var[i]) clipped to [orig - par_clip[i], orig + par_clip[i]]
     * (NaN = no clip; min > max collapses to the max like np.clip), or U(orig - var, orig + var).  par_var / par_clip
     * arrive already scaled for var_type "relative" (var * |orig|, clip * orig — sign kept, as the reference does).
     * Jx, Jy, Jz, Jxz and the aspect ratio stay at their construction-time values inside the dynamics (pyfly computes
     * the inertia terms once, pyfly.py:1086-1119).  model_on selects kernels that read per-env parameters. ---- */
    int32_t model_on, model_uniform;
    int32_t par_enabled[FW_NPARAM];
    double par_orig[FW_NPARAM], par_var[FW_NPARAM], par_clip[FW_NPARAM];

    /* ---- error integrals and strided observation rows (fixed_wing.py:83, 1003-1012, 1129-1138, 1165-1180) ----
     * integration_window W: the "int_error" reward value is sum(history["error"][-W:]) padded with (W - steps) copies of
     * the first error while the episode is younger than W — and, python slicing being what it is, the sum of the WHOLE
     * history for W == 0, the value every config of the reference tree carries.  The "integrator" observation of the row
     * with lag i is sum(history["error"][-W-i:-i]) + (W - (steps - i)) * history["error"][0] while steps - i < W; the
     * reset observation reads the history of the episode that just ENDED (the new one is installed after the
     * observation is taken, fixed_wing.py:453-460), or error * W on the very first reset.  Errors live in the 50-deep
     * ring of the end_error metric: W + lag_max <= FW_END_ERR_WINDOW - 1.
     * obs_step s: row k of the observation has lag 1 + k s (range(1, length * s, s)); lag_max <= 5. ---- */
    int32_t integration_window, obs_step;

    /* ---- target class "attitude_angular" (fixed_wing.py:671-675, 741-746, 1455-1460, 1558-1642): omega_p, omega_q,
     * omega_r become target states 3, 4, 5 whose targets _attitude_to_angular_rates derives every step from the roll /
     * pitch errors and the previous rate targets.  They take part in the goal status (ang_bound, raw units: the
     * reference keeps the raw props of these states, no degree conversion; inf = no bound), in reward factors and
     * observation entries that name them (index 3 + a), in info["target"] and in the per-state metrics
     * (FW_NMETRIC_ANG values behind the 28 of the base states).  All three or none. ---- */
    int32_t ang_on, _pad_ang;
    double ang_max_vel[3];                /* props "max_vel", default radians(180) */
    double ang_bound[3];
} FwConfig;

typedef struct FwHandle FwHandle;

/* Replaces: FixedWingAircraft.__init__ (fixed_wing.py:14-306) + PyFly.__init__ (pyfly.py:1054-1249) for n_envs envs.
 * `device` is the CUDA ordinal.  All envs start un-reset (call fw_reset with mask=NULL). */
int fw_create(const FwConfig* cfg, int32_t n_envs, int32_t device, FwHandle** out);
int fw_destroy(FwHandle* h);
int fw_obs_dim(const FwHandle* h);   /* floats per observation row of this handle */

/* Replaces: FixedWingAircraft.set_curriculum_level (fixed_wing.py:334-412) and FixedWingAircraft.seed (:324-332) on a
 * LIVE handle, as the reference's training scripts call them in the middle of a run
 * (examples/train_rl_controller.py:137, env_method("set_curriculum_level", ...)).  `cfg` must equal the handle's
 * configuration except for the reset-time fields: init_lo / init_hi, wind_mag_min / wind_mag_max, tgt_low / tgt_high /
 * tgt_delta, the tgt_slope / tgt_amp / tgt_period ranges and seed (anything else: FW_EINVAL).  The new values apply to
 * every reset and target resampling from now on; episodes that are running keep their state, their turbulence stream
 * and their Philox key (an episode is keyed by the seed in force when it was reset).  Device pointers of the handle
 * do not change, so CUDA graphs captured over fw_step stay valid.  Asynchronous on `stream` (not during capture):
 * the precomputed next-episode rows are recomputed there. */
int fw_set_config(FwHandle* h, const FwConfig* cfg, void* stream);

/* Checkpoint / resume (the reference checkpoints the policy and the VecNormalize statistics,
 * common/base_class.py:560,645, vec_env/vec_normalize.py:222-243, and restarts every episode; here the envs themselves
 * can be saved too).  The blob is ONE contiguous device buffer of fw_state_blob_size(h) bytes owned by the caller: a
 * header (shape check) followed by the whole structure-of-arrays state, the per-env episode bookkeeping, the
 * precomputed next-episode rows and the reset-time configuration (fw_set_config values).  fw_set_state_blob on a handle
 * created with the same configuration and n_envs continues every episode bit-identically (turbulence streams included:
 * they are functions of the stored filter states and Philox counters).  Not stored: an injected-noise buffer (caller
 * memory, fw_reset) and profiling sums.  Asynchronous on `stream` except for the header check of fw_set_state_blob. */
int64_t fw_state_blob_size(const FwHandle* h);
int fw_get_state_blob(FwHandle* h, void* blob_dev, void* stream);
int fw_set_state_blob(FwHandle* h, const void* blob_dev, void* stream);

/* Measurement aid (no reference counterpart; bench.py's per-kernel roofline): with profiling on, every fw_step /
 * fw_step_random step records CUDA events on the launch stream around each of its kernels and ends with an event
 * synchronise; fw_get_profile returns the summed durations in ms of {init, integrate (RK45 attempt loop or RK4),
 * head} and the number of steps they cover.  fw_set_profiling(h, 0 or 1) also clears the sums. */
int fw_set_profiling(FwHandle* h, int32_t on);

/* fw_step with auto_reset recomputes, on a side stream of the handle, the precomputed next-episode rows that the step
 * consumed; the next fw_step / fw_reset waits for that work by itself.  fw_join makes `stream` wait for it explicitly,
 * e.g. before handing the handle to another thread.  CUDA stream capture: steps launched on a capturing stream join
 * their side-stream work themselves, so a capture may end after any step; a capture must BEGIN after the work issued
 * before it has completed (torch.cuda.graph() synchronises the device) or after fw_join + a stream synchronise.
 * No reference counterpart. */
int fw_join(FwHandle* h, void* stream);

/* Episode-end rows for a host-facing caller (VecEnv infos: Monitor's {"episode": {r, l}}, the gym env's metrics and
 * terminal_observation, fixed_wing.py:589-628 / vec_env auto-reset).  After every auto-reset fw_step the buffer holds
 *   int32 count at byte 0 (number of episodes that ended at the step; may exceed cap),
 *   then, from byte 8, min(count, cap) rows of FW_INFO_HEAD + obs_dim doubles:
 *   env index, FwTermCode, episode length, episode return, FW_NMETRIC metrics, terminal observation (float32 values)
 * in no particular order.  Placed behind the step outputs in one allocation it reaches the host in the same copy.
 * rows_dev = NULL switches it off.  The buffer must hold 8 + cap * (FW_INFO_HEAD + obs_dim) * 8 bytes. */
#define FW_INFO_HEAD (4 + FW_NMETRIC)
int fw_set_info_rows(FwHandle* h, double* rows_dev, int32_t cap);
int fw_get_profile(const FwHandle* h, double* ms_sum3, int64_t* steps);

/* Waypoint head only: the task table.  tasks_dev [n_tasks, wp_len, FW_WP_ROW] f64 (device, copied), task_of_env_dev [n]
 * int32 (device, copied): the task every env flies.  Replaces FixedWingAircraft_simple.sample_tasks / reset_task
 * (simple_train.py:330-375).  Must be called before the first fw_reset. */
int fw_set_waypoint_tasks(FwHandle* h, const double* tasks_dev, int32_t n_tasks, int32_t wp_len,
                          const int32_t* task_of_env_dev, void* stream);
const char* fw_last_error(void);
int fw_abi_version(void);
/* sizeof() of the structs a binding has to mirror field by field (FwConfig, FwRolloutPost, FwReplay, FwReplayNorm), as the
 * library was compiled: a ctypes / cgo / JNI stub compares them with its own layout at load time and refuses a mismatch
 * (tum_adlr_deep_reinforcement_learning_b200/_lib.py does). */
int fw_config_size(void);
int fw_rollout_post_size(void);
int fw_replay_size(void);
int fw_replay_norm_size(void);

/* Replaces: FixedWingAircraft.reset(state, target, turbulence_noise) (fixed_wing.py:414-481 -> pyfly.py:1262-1311).
 *  mask_dev        [n] uint8, nullable (NULL = all envs): which envs to reset.
 *  state_dev       [n, FW_NSTATE_INJECT] f64, nullable: injected initial state (NaN entries = sample that entry).
 *  target_dev      [n, 3] f64, nullable: injected targets (roll, pitch, Va).
 *  noise_dev       [n, 4, noise_len] f64, nullable: injected unit white noise (pyfly.py:1294, dryden.py:184-188);
 *                  when NULL turbulence noise is generated from Philox. The buffer must stay alive while in use.
 *  obs_dev         [n, obs_dim] f32, nullable: reset observation for the masked envs (others untouched);
 *                  obs_dim = FW_NOBS for the default layout, obs_len * obs_n for a general one (fw_obs_dim).
 *  obs64_dev       [n, obs_dim] f64, nullable.
 */
int fw_reset(FwHandle* h, const uint8_t* mask_dev, const double* state_dev, const double* target_dev,
             const double* noise_dev, int32_t noise_len, float* obs_dev, double* obs64_dev, void* stream);

/* Replaces: VecEnv.step_async/step_wait over FixedWingAircraft.step (fixed_wing.py:483-628 -> pyfly.py:1358-1420),
 * including the VecEnv auto-reset contract (subproc_vec_env.py:26-31 / dummy_vec_env.py:46-50).
 *  actions_dev     [n, 3]; f32 when actions_f64==0, f64 otherwise.  Raw agent actions (pre-clip, pre-scale).
 *  obs_dev         [n, FW_NOBS] f32: next observation (the RESET observation for envs that finished).
 *  rew_dev         [n] f32;  done_dev [n] uint8.
 *  term_obs_dev    [n, FW_NOBS] f32, nullable: info["terminal_observation"] rows (written only where done).
 *  obs64_dev, rew64_dev   nullable f64 copies (parity tests / PID harness that consumes f64 obs).
 *  auto_reset      0: leave finished envs in their terminal state (caller resets, e.g. evaluate_controller.py:174).
 */
int fw_step(FwHandle* h, const void* actions_dev, int32_t actions_f64, float* obs_dev, float* rew_dev,
            uint8_t* done_dev, float* term_obs_dev, double* obs64_dev, double* rew64_dev, int32_t auto_reset,
            void* stream);

/* k_steps env steps with Philox U(-1,1)^3 actions generated in-kernel (random-action throughput workloads C2/C3 of
 * BASELINE.json; the policy-free analogue of collect_rollouts, on_policy_algorithm.py:123-191): no action buffer is
 * read, the launches of consecutive steps are queued back to back on `stream`, auto-reset is on.  The outputs hold
 * the result of the LAST step. */
int fw_step_random(FwHandle* h, int32_t k_steps, uint64_t action_seed, float* obs_dev, float* rew_dev,
                   uint8_t* done_dev, void* stream);

/* Per-env info on done: term_code [n] int32, metrics [n, FW_NMETRIC] f64, episode return/length (Monitor,
 * common/monitor.py:99-113).  Valid for envs whose done flag was set by the most recent fw_step. */
/* attitude_angular configs: the FW_NMETRIC_ANG per-state metrics of omega_p / omega_q / omega_r [n, FW_NMETRIC_ANG] f64
 * (avg_error, total_error, end_error, rise_time, overshoot, success, settling_time, success_time_frac x 3 states,
 * metric-major; get_metric, fixed_wing.py:1644-1736), valid like fw_get_episode_info.  FW_EINVAL for other configs. */
int fw_get_episode_info_angular(FwHandle* h, double* metrics_ang_dev, void* stream);

int fw_get_episode_info(FwHandle* h, int32_t* term_code_dev, double* metrics_dev, double* ep_return_dev,
                        int32_t* ep_length_dev, void* stream);

/* State access for parity tests and checkpointing (SoA -> [n, width] row-major f64 / int32).
 * field ids: see FwField. */
enum FwField {
    FW_FIELD_Y = 0,          /* [n,19] f64 ODE state                                          */
    FW_FIELD_EULER = 1,      /* [n,3]  roll pitch (state .value); yaw is not materialised: NaN */
    FW_FIELD_VAB = 2,        /* [n,3]  Va alpha beta                                          */
    FW_FIELD_WIND = 3,       /* [n,3]  steady wind NED                                        */
    FW_FIELD_TARGET = 4,     /* [n,3]  roll pitch Va targets                                  */
    FW_FIELD_CMD = 5,        /* [n,3]  constrained elevator/aileron/throttle commands         */
    FW_FIELD_TURB = 6,       /* [n,6]  turbulence sample used by the NEXT step (lin3, ang3)   */
    FW_FIELD_COUNTERS = 7,   /* [n,4]  int32: steps_count, steps_for_target, sim_step, episode*/
    FW_FIELD_NFEV = 8,       /* [n,2]  int32: RHS evaluations, RK attempts of the last step   */
    FW_FIELD_PARAMS = 9,     /* [n,FW_NPARAM] f64 aircraft parameters of the running episode (model_on handles only) */
    FW_FIELD_ATARGET = 10,   /* [n,3]  omega_p omega_q omega_r rate targets (attitude_angular configs; zeros otherwise) */
    FW_FIELD_COUNT = 11
};
int fw_get_field(FwHandle* h, int32_t field, void* out_dev, void* stream);
int fw_set_field(FwHandle* h, int32_t field, const void* in_dev, void* stream);

/* Replaces: RolloutBuffer.compute_returns_and_advantage (buffers.py:304-333), time-major [T, N] f32 arrays.
 * dones[t] is the "episode ended before obs[t]" flag stored by RolloutBuffer.add (on_policy_algorithm.py:178);
 * last_done is the done vector returned by the final env.step (uint8).  Bit-exact with the reference's mixed
 * f32-delta / f64-carry arithmetic. */
int fw_gae(const float* rew_dev, const float* val_dev, const float* done_dev, const float* last_val_dev,
           const uint8_t* last_done_dev, float* adv_dev, float* ret_dev, int32_t T, int32_t N, float gamma,
           float gae_lambda, void* stream);

/* clip_grad_norm_(max_norm, L2) followed by one Adam step (torch.optim.Adam: no weight decay, no amsgrad) over a flat
 * float32 parameter vector and its flat gradient (stable_baselines3/ppo/ppo.py:212-214).  step_dev: device float,
 * incremented by the call.  max_norm <= 0 disables clipping.  One launch, one block: meant for small policies. */
int fw_adam_clip_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                      float* step_dev, int32_t n, float lr, float beta1, float beta2, float eps, float max_norm,
                      void* stream);

/* One env step of rollout glue on the device: VecNormalize.step_wait + RunningMeanStd.update
 * (common/vec_env/vec_normalize.py:106-127, common/running_mean_std.py:19-39), RolloutBuffer.add
 * (common/buffers.py:292-302; the row stores the PREVIOUS observation / done flags, on_policy_algorithm.py:178-180) and
 * Monitor-style episode totals (common/monitor.py:99-113).  All pointers are device pointers; statistics are float64. */
#define FW_ROLLOUT_BLOCKS 64
#define FW_ROLLOUT_SCRATCH(obs_dim) (3 * (obs_dim) + 4 + FW_ROLLOUT_BLOCKS * (2 * (obs_dim) + 5))
typedef struct FwRolloutPost {
    /* this step */
    const float* obs_raw;      /* [n, obs_dim] raw observation returned by fw_step                                   */
    const float* rew_raw;      /* [n]                                                                                */
    const uint8_t* done;       /* [n]                                                                                */
    const float* actions;      /* [n, act_dim] the actions that were stepped                                         */
    const float* values;       /* [n]  value head output for last_obs                                                */
    const float* log_probs;    /* [n]                                                                                */
    /* persistent state */
    float* last_obs;           /* [n, obs_dim] normalised observation the policy saw; replaced by the new one        */
    float* last_dones;         /* [n] float 0/1; replaced by `done`                                                  */
    double* ret;               /* [n] discounted return accumulators (VecNormalize.ret)                              */
    double* obs_mean; double* obs_var; double* obs_count;      /* [obs_dim], [obs_dim], [1]                           */
    double* ret_mean; double* ret_var; double* ret_count;      /* [1] each                                           */
    double* run_ret; double* run_len;                          /* [n] running episode return / length                */
    double* ep_stats;          /* [3] += sum of finished returns, sum of finished lengths, number of finished         */
    /* rollout buffer row t */
    float* buf_obs; float* buf_actions; float* buf_rewards; float* buf_dones; float* buf_values; float* buf_log_probs;
    double* scratch;           /* [FW_ROLLOUT_SCRATCH(obs_dim)] doubles, zeroed once by the caller (holds a ticket counter) */
    int32_t n, obs_dim, act_dim;
    float gamma, clip_obs, clip_reward, epsilon;
    int32_t norm_obs, norm_reward, training;
} FwRolloutPost;
int fw_rollout_post_step(const FwRolloutPost* p, void* stream);

/* PPO minibatch loss and its gradient with respect to the network outputs, fused (stable_baselines3/ppo/ppo.py:163-218:
 * advantage normalisation, ratio, clipped surrogate, value MSE, Gaussian entropy; diagonal Gaussian with a
 * state-independent log_std, common/distributions.py:130-175).  Inputs are the policy mean [B,3], the value head
 * output [B], log_std [3] and the rollout-buffer minibatch (actions [B,3], old_log_prob, advantages, returns [B]).
 * Outputs: losses_dev[3] = {loss, policy_loss, value_loss}; grad_mean_dev [B,3], grad_values_dev [B], grad_log_std_dev
 * [3] = d loss / d input (what autograd would hand to the networks' backward).  scratch_dev: FW_PPO_SCRATCH_DOUBLES
 * doubles (sums are formed without floating-point atomics, block partials added in a fixed order: the same inputs give
 * the same bits on every run). */
#define FW_PPO_SCRATCH_DOUBLES (9 + 7 * 592)
int fw_ppo_loss(const float* mean_dev, const float* values_dev, const float* log_std_dev, const float* actions_dev,
                const float* old_log_prob_dev, const float* adv_dev, const float* returns_dev, int32_t batch,
                float clip_range, float ent_coef, float vf_coef, double* scratch_dev, float* grad_mean_dev,
                float* grad_values_dev, float* grad_log_std_dev, float* losses_dev, void* stream);

/* Replay ring of the off-policy (SAC) path in HBM.  Replaces ReplayBuffer.add / sample / _get_samples
 * (stable_baselines3/common/buffers.py:146-256), widened from the reference's single env to n envs per insert.
 * A transition is one packed row  obs[obs_dim] | next_obs[obs_dim] | action[act_dim] | reward | done(0/1)  padded to a
 * multiple of four floats; the ring stores ORIGINAL observations and rewards and fw_replay_sample normalises with the
 * statistics current at sample time, as the reference does (buffers.py:245-254, off_policy_algorithm.py:430-436).
 * head_dev / size_dev / sample_calls_dev are device scalars (int64, zero-initialised by the caller) so that both calls
 * can be captured in a CUDA graph.  Sample b of a call reads ring row floor(u * size), u = Philox4x32-10(seed; b, number of
 * the call); indices_out_dev (nullable) receives the rows that were read. */
#define FW_REPLAY_ROW_FLOATS(obs_dim, act_dim) ((2 * (obs_dim) + (act_dim) + 2 + 3) / 4 * 4)
typedef struct FwReplay {
    float* rows;               /* [capacity, row_floats], 16-byte aligned */
    int64_t* head_dev;         /* next row to write */
    int64_t* size_dev;         /* rows filled so far, <= capacity */
    int64_t* sample_calls_dev; /* number of fw_replay_sample calls so far (Philox counter) */
    int64_t capacity;
    int32_t obs_dim, act_dim, row_floats, _pad;
} FwReplay;
typedef struct FwReplayNorm {  /* VecNormalize state used at sample time (vec_normalize.py:150-172); NULL pointers with flags 0 */
    const double* obs_mean; const double* obs_var; const double* ret_var;
    float clip_obs, clip_reward, epsilon;
    int32_t norm_obs, norm_reward, _pad;
} FwReplayNorm;
int fw_replay_insert(const FwReplay* rb, const float* obs_dev, const float* next_obs_dev, const float* actions_dev,
                     const float* rewards_dev, const uint8_t* dones_dev, int32_t n, void* stream);
int fw_replay_sample(const FwReplay* rb, const FwReplayNorm* norm, int32_t batch, uint64_t seed, float* obs_out_dev,
                     float* actions_out_dev, float* next_obs_out_dev, float* dones_out_dev, float* rewards_out_dev,
                     int64_t* indices_out_dev, void* stream);

/* The PPO gradient all-reduce (one mean of the flattened 42 KB policy gradient per optimiser step, the only
 * collective of the path) fused with clip_grad_norm_ + Adam (ppo/ppo.py:212-214) in ONE single-block kernel per rank over
 * NVLink peer memory: each rank's buffer is mapped by its peers through CUDA IPC, ranks signal and wait through flag words
 * in each other's buffers, every rank pushes its gradient into the peers' buffers, adds the slots in rank order
 * (bit-identical replicas) and applies the update.  Replaces ncclAllReduce + a divide + fw_adam_clip_step.  One process per GPU on ONE node:
 *   fw_comm_create   allocates the rank's buffer (flags + 2 x world slots of n_floats: every peer pushes into its slot)
 *   fw_comm_export   -> FW_COMM_HANDLE_BYTES bytes to hand to every peer (e.g. torch.distributed.all_gather)
 *   fw_comm_connect  handles of ALL ranks, world x FW_COMM_HANDLE_BYTES bytes in rank order (the own entry is ignored)
 *   fw_comm_allreduce_adam   grad_dev is replaced by the mean gradient; then as fw_adam_clip_step.  Asynchronous on `stream`,
 *                    capturable; every rank must make the same sequence of calls.  A peer that does not arrive within ~2 s
 *                    sets the error word (fw_comm_error returns 1) instead of hanging the GPU. */
#define FW_COMM_MAX_WORLD 8
#define FW_COMM_HANDLE_BYTES 64
typedef struct FwComm FwComm;
int fw_comm_create(int32_t n_floats, int32_t world, int32_t rank, int32_t device, FwComm** out);
int fw_comm_export(FwComm* c, void* handle64);
int fw_comm_connect(FwComm* c, const void* handles);
int fw_comm_allreduce_adam(FwComm* c, float* param_dev, float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                           float* step_dev, int32_t n, float lr, float beta1, float beta2, float eps, float max_norm,
                           void* stream);
int fw_comm_error(FwComm* c);
int fw_comm_destroy(FwComm* c);
const char* fw_comm_last_error(void);

/* Vector-pipe peak micro-benchmarks (dependent-chain-free FMA loops) used as roofline denominators by bench.py:
 * returns achieved TFLOP/s (2 flop per FMA) measured with CUDA events on `device`. */
int fw_measure_fma_peak(int32_t device, int32_t precision, double* tflops_out);

/* Diagnostic: evaluates the straight-line FP64 math of the RHS hot loop elementwise (op 0: exp(x), 1: asin(x),
 * 2: atan2(y, x), 3: 1/sqrt(x), 4..6: ops 0..2 built with immediate instead of constant-bank coefficients, 7: log(x),
 * 8 / 9: x^y as exp(y log x), constant-bank / immediate build); used by the tests to bound their error against the host libm. */
int fw_debug_math(int32_t op, const double* x_dev, const double* y_dev, double* out_dev, int32_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FWB200_H */
