"""Host-side configuration: reference-format JSON / dict configs -> the POD `FwConfig` of include/fwb200.h.

Mirrors what the reference does at construction time:
  * PyFly.__init__ (magpie/libs/pyfly/pyfly/pyfly.py:1054-1249): aircraft parameters, pyfly_config.json,
    recursive `config_kw` overrides (:1067-1073, :1125-1126), Actuation.finalize limits (:584-623);
  * FixedWingAircraft.__init__ / set_curriculum_level (magpie/libs/fixed-wing-gym/gym_fixed_wing/fixed_wing.py:14-306,
    :334-412): the gym JSON, `config_kw` / `sim_config_kw` overrides (:34-62), simulator-state overrides and the
    curriculum scaling of init ranges and target ranges;
  * Wind.__init__ -> DrydenGustModel.__init__ (pyfly.py:780-783, dryden.py:52-143) including the reference's
    mis-ordered constructor arguments, and scipy.signal.lsim's discretisation.
The numbers of the default configs below are those of pyfly_config.json / fixed_wing_config.json.
"""
import copy
import ctypes
import json
import math
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

FW_ABI_VERSION = 10
FW_NY, FW_NOBS, FW_NACT, FW_NSTATE_INJECT, FW_NMETRIC = 19, 14, 3, 21, 28
FW_F64, FW_F32 = 0, 1
FW_INT_RK45_SCIPY, FW_INT_RK4_FIXED = 0, 1
TERM_NAMES = {0: None, 1: "steps", 2: "success", 10: "omega_p", 11: "omega_q", 12: "omega_r", 13: "Va"}
TARGET_STATES = ("roll", "pitch", "Va")
ANGULAR_TARGET_STATES = ("omega_p", "omega_q", "omega_r")      # target class attitude_angular: target states 3, 4, 5
ALL_TARGET_STATES = TARGET_STATES + ANGULAR_TARGET_STATES
GOAL_STATES = ("roll", "pitch", "Va", "all")
INIT_STATES = ("roll", "pitch", "yaw", "omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d",
               "velocity_u", "velocity_v", "velocity_w")
METRIC_LAYOUT = (("rise_time", 0, TARGET_STATES), ("settling_time", 3, GOAL_STATES), ("overshoot", 7, TARGET_STATES),
                 ("total_error", 10, TARGET_STATES), ("avg_error", 13, TARGET_STATES),
                 ("control_variation", 16, ("all",)), ("success", 17, GOAL_STATES),
                 ("success_time_frac", 21, GOAL_STATES), ("end_error", 25, TARGET_STATES))


class FwFilter(ctypes.Structure):
    _fields_ = [("order", ctypes.c_int32), ("noise_row", ctypes.c_int32), ("Ad", ctypes.c_double * 9),
                ("Bd0", ctypes.c_double * 3), ("Bd1", ctypes.c_double * 3), ("C", ctypes.c_double * 3),
                ("D", ctypes.c_double), ("Ablk", ctypes.c_double * 9)]


_AERO = ("mass Jx Jy Jz Jxz S_wing b c S_prop C_prop k_motor k_T_P k_Omega e_oswald M a_0 "
         "C_L_0 C_L_alpha C_L_q C_L_delta_e C_D_p C_D_q C_D_beta1 C_D_beta2 C_D_delta_e "
         "C_m_0 C_m_alpha C_m_q C_m_delta_e C_m_fp C_Y_0 C_Y_beta C_Y_p C_Y_r C_Y_delta_a C_Y_delta_r "
         "C_l_0 C_l_beta C_l_p C_l_r C_l_delta_a C_l_delta_r C_n_0 C_n_beta C_n_p C_n_r C_n_delta_a C_n_delta_r").split()

_d, _i = ctypes.c_double, ctypes.c_int32


class FwConfig(ctypes.Structure):
    """ctypes mirror of `struct FwConfig` (include/fwb200.h) — field order and types must match exactly;
    tests/test_capi.py compares sizeof with the value both shared libraries report."""
    _fields_ = (
        [("abi_version", _i), ("precision", _i), ("integrator", _i), ("rk4_substeps", _i), ("rtol", _d), ("atol", _d)]
        + [(n, _d) for n in _AERO]
        + [("dt", _d), ("rho", _d), ("g", _d),
           ("elevon_min", _d), ("elevon_max", _d), ("elevon_dot_max", _d), ("elevon_omega0", _d), ("elevon_zeta", _d),
           ("throttle_min", _d), ("throttle_max", _d), ("throttle_tau", _d),
           ("omega_con_min", _d * 3), ("omega_con_max", _d * 3), ("va_value_min", _d), ("va_con_max", _d),
           ("init_lo", _d * 12), ("init_hi", _d * 12), ("wind_mag_min", _d), ("wind_mag_max", _d),
           ("turbulence", _i), ("_pad0", _i), ("turb_noise_scale", _d), ("filt", FwFilter * 6),
           ("steps_max", _i), ("scale_actions", _i), ("scale_low", _d), ("scale_high", _d),
           ("act_lo", _d * 3), ("act_hi", _d * 3), ("has_action_bounds", _i), ("_pad1", _i),
           ("action_bounds_min", _d * 3), ("action_bounds_max", _d * 3),
           ("tgt_low", _d * 3), ("tgt_high", _d * 3), ("tgt_delta", _d * 3), ("tgt_bound", _d * 3),
           ("tgt_class", _i * 3), ("tgt_radians", _i * 3), ("tgt_slope_low", _d * 3), ("tgt_slope_high", _d * 3),
           ("tgt_amp_low", _d * 3), ("tgt_amp_high", _d * 3), ("tgt_period_low", _d * 3), ("tgt_period_high", _d * 3),
           ("rng_u_override", _d), ("on_success", _i), ("streak_req", _i), ("resample_every", _i),
           ("streak_fraction", _d),
           ("rew_err_scaling", _d * 3), ("rew_err_max", _d * 3), ("rew_delta_scaling", _d), ("rew_delta_max", _d),
           ("rew_bound_scaling", _d), ("rew_bound_max", _d), ("rew_delta_window", _i), ("obs_act_window", _i),
           ("step_fail_timesteps", _i), ("_pad2", _i), ("step_fail_value", _d), ("rise_low", _d), ("rise_high", _d),
           ("obs_noise_mean", _d), ("obs_noise_std", _d),
           ("rew_generic", _i), ("rew_n", _i), ("rew_potential", _i), ("rew_nterms", _i),
           ("rew_class", _i * 12), ("rew_idx", _i * 12), ("rew_fclass", _i * 12), ("rew_shaping", _i * 12),
           ("rew_window", _i * 12), ("rew_value_timesteps", _i * 12), ("rew_scaling", _d * 12), ("rew_maxv", _d * 12),
           ("rew_sign", _d * 12), ("rew_value", _d * 12), ("term_fclass", _i * 4), ("term_weight", _d * 4),
           ("obs_generic", _i), ("obs_len", _i), ("obs_n", _i), ("obs_normalize", _i),
           ("obs_kind", _i * 16), ("obs_idx", _i * 16), ("obs_window", _i * 16), ("obs_norm_flag", _i * 16),
           ("obs_mean", _d * 16), ("obs_var", _d * 16), ("obs_init_noise", _d),
           ("env_kind", _i), ("turb_block_len", _i), ("wp_goal_bound", _d * 3), ("wp_rew_range", _d * 3),
           ("seed", ctypes.c_uint64), ("env_id_offset", ctypes.c_int64),
           ("model_on", _i), ("model_uniform", _i), ("par_enabled", _i * 48), ("par_orig", _d * 48), ("par_var", _d * 48),
           ("par_clip", _d * 48), ("integration_window", _i), ("obs_step", _i),
           ("ang_on", _i), ("_pad_ang", _i), ("ang_max_vel", _d * 3), ("ang_bound", _d * 3)])


def _target_index(name, c):
    """Index of a target state: roll, pitch, Va = 0..2; omega_p/q/r = 3..5 when the config carries attitude_angular targets."""
    if name in TARGET_STATES:
        return TARGET_STATES.index(name)
    if name in ANGULAR_TARGET_STATES and c.ang_on:
        return 3 + ANGULAR_TARGET_STATES.index(name)
    raise KeyError("%r is not a target state of this config" % name)


def _var(name, **kw):
    kw["name"] = name
    return kw


def default_sim_config():
    """The reference's pyfly_config.json, restated (plots omitted: rendering is out of scope)."""
    elevon = dict(value_min=-30, value_max=35, init_min=0, init_max=0, convert_to_radians=True, order=2,
                  dot_max=3.4907, omega_0=100, zeta=1.71)
    ang = dict(convert_to_radians=True, init_min=-30, init_max=30)
    rate = dict(convert_to_radians=True, init_min=-40, init_max=40, constraint_min=-180, constraint_max=180)
    return {
        "dt": 0.01, "rho": 1.225, "g": 9.81, "wind_magnitude_min": -8, "wind_magnitude_max": 8,
        "turbulence": True, "turbulence_intensity": "light", "turbulence_sim_length": 300,
        "actuation": {"dynamics": ["elevon_right", "elevon_left", "throttle"],
                      "inputs": ["elevator", "aileron", "throttle"]},
        "variables": [
            _var("roll", wrap=True, **ang), _var("pitch", **ang), _var("yaw", wrap=True, **ang),
            _var("omega_p", **rate), _var("omega_q", **rate), _var("omega_r", **rate),
            _var("position_n", init_min=-100, init_max=100), _var("position_e", init_min=-100, init_max=100),
            _var("position_d", init_min=-20, init_max=-100),
            _var("velocity_u", init_min=13, init_max=22), _var("velocity_v", init_min=-5, init_max=5),
            _var("velocity_w", init_min=-5, init_max=5),
            _var("alpha"), _var("beta", convert_to_radians=True), _var("Va", value_min=1e-6),
            _var("elevon_left", **elevon), _var("elevon_right", **elevon),
            _var("throttle", value_min=0, value_max=1, init_min=0, init_max=0, order=1, tau=0.2),
        ],
    }


def default_env_config():
    """The reference's fixed_wing_config.json, restated (render block omitted)."""
    def st(name, lo=None, hi=None, rad=False, **kw):
        d = {"name": name, "type": "state"}
        if lo is not None:
            d.update(low=lo, high=hi)
        if rad:
            d["convert_to_radians"] = True
        d.update(kw)
        return d

    rate = dict(init_min=-60, init_max=60, constraint_min=-720, constraint_max=720, convert_to_radians=True)
    return {
        "steps_max": 2000, "integration_window": 0,
        "observation": {
            "length": 1, "step": 1, "shape": "vector", "normalize": False, "noise": {"mean": 0, "var": 0},
            "states": [st("roll", -180, 180, True), st("pitch", -85, 85, True), st("Va", 0, 70),
                       st("omega_p", -720, 720, True), st("omega_q", -720, 720, True), st("omega_r", -720, 720, True),
                       {"name": "roll", "type": "target", "value": "absolute"},
                       {"name": "pitch", "type": "target", "value": "absolute"},
                       {"name": "Va", "type": "target", "value": "absolute"},
                       st("alpha"), st("beta"),
                       {"name": "elevator", "type": "action", "window_size": 5},
                       {"name": "aileron", "type": "action", "window_size": 5},
                       {"name": "throttle", "type": "action", "window_size": 5}]},
        "action": {"scale_space": True, "scale_low": -1, "scale_high": 1, "bounds_multiplier": 1.5,
                   "states": [{"name": n, "low": "max", "high": "max"} for n in ("elevator", "aileron", "throttle")]},
        "target": {
            "resample_every": 0, "success_streak_req": 100, "success_streak_fraction": 0.95, "on_success": "none",
            "states": [
                {"name": "roll", "convert_to_radians": True, "low": -60, "high": 60, "delta": 180,
                 "class": "constant", "bound": 5},
                {"name": "pitch", "convert_to_radians": True, "low": -25, "high": 25, "delta": 45,
                 "class": "constant", "bound": 5},
                {"name": "Va", "low": 15, "high": 28, "delta": 6, "class": "compensate", "bound": 2}]},
        "reward": {
            "form": "absolute", "randomize_scaling": False, "step_fail": "timesteps",
            "terms": [{"function_class": "linear", "weight": 1}],
            "factors": [
                {"name": "roll", "class": "state", "type": "error", "function_class": "linear", "scaling": 3.2,
                 "shaping": True, "max": 0.3, "sign": -1},
                {"name": "pitch", "class": "state", "type": "error", "function_class": "linear", "scaling": 3.2,
                 "shaping": True, "max": 0.3, "sign": -1},
                {"name": "Va", "class": "state", "type": "error", "function_class": "linear", "scaling": 25,
                 "shaping": True, "max": 0.3, "sign": -1},
                {"name": "action", "class": "action", "type": "delta", "function_class": "linear", "window_size": 5,
                 "scaling": 60, "shaping": False, "sign": -1},
                {"name": "action_bound", "class": "action", "type": "bound", "function_class": "linear",
                 "scaling": 1, "shaping": False, "sign": -1}]},
        "simulator": {"states": [
            {"name": "roll", "init_min": -110, "init_max": 110, "convert_to_radians": True},
            {"name": "pitch", "init_min": -45, "init_max": 45, "convert_to_radians": True},
            {"name": "velocity_u", "init_min": 10, "init_max": 23},
            {"name": "velocity_v", "init_min": -5, "init_max": 5},
            {"name": "velocity_w", "init_min": -5, "init_max": 5},
            {"name": "Va", "constraint_max": 70},
            dict(name="omega_p", **rate), dict(name="omega_q", **rate), dict(name="omega_r", **rate)]},
        "metrics": [{"name": "rise_time", "high": 0.9, "low": 0.1}, {"name": "settling_time"}, {"name": "overshoot"},
                    {"name": "total_error"}, {"name": "avg_error"}, {"name": "control_variation"},
                    {"name": "success"}, {"name": "success_time_frac"}, {"name": "end_error"}],
    }


def load_aircraft_parameters(path=None):
    """x8_param.mat scalars (pyfly.py:1076-1078); shipped as JSON printed to 17 digits by tests/golden/make_golden.py.
    A `.mat` path is read with scipy.io like the reference does."""
    if path is None:
        path = os.path.join(_DATA, "x8_params.json")
    if path.endswith(".mat"):
        import scipy.io
        raw = scipy.io.loadmat(path, squeeze_me=True)
        return {k: float(v) for k, v in raw.items() if not k.startswith("__") and np.size(v) == 1}
    with open(path) as f:
        return json.load(f)


def apply_overrides(parent, kws):
    """Recursive override, same semantics as fixed_wing.py:34-40 (dict keys, or int keys into lists)."""
    for attr, val in kws.items():
        if isinstance(val, dict):
            apply_overrides(parent[attr], val)
        else:
            parent[attr] = val


def _tf2ss(num, den):
    """Controller-canonical realisation as produced by scipy.signal.tf2ss for a strictly proper SISO system."""
    num = np.atleast_1d(np.asarray(num, dtype=np.float64))
    den = np.atleast_1d(np.asarray(den, dtype=np.float64))
    num = num / den[0]
    den = den / den[0]
    n = len(den) - 1
    num = np.concatenate([np.zeros(n + 1 - len(num)), num])
    A = np.zeros((n, n))
    A[0, :] = -den[1:]
    if n > 1:
        A[1:, :-1] = np.eye(n - 1)
    B = np.zeros((n, 1))
    B[0, 0] = 1.0
    D = num[0]
    C = (num[1:] - num[0] * den[1:]).reshape(1, n)
    return A, B, C, D


def dryden_filters(sim_length, sim_dt, wingspan, intensity="light", spec=False):
    """Discretised Dryden shaping filters in scipy.signal.lsim's linear-interpolation form.

    Reference parameterisation (default): `Wind` calls DrydenGustModel(sim_length, dt, b, intensity=...)
    (pyfly.py:781-783) against the signature (dt, b, h=100, V_a=25, intensity) (dryden.py:52), hence inside the
    model dt = sim_length, b = sim dt, h = wingspan.  `spec=True` gives the intended MIL-F-8785C parameters
    (dt = sim dt, b = wingspan, h = 100 m) — a labelled deviation from the reference, never used for parity.
    Returns (filters [(order, noise_row, Ad, Bd0, Bd1, C, D, Ablk)], noise_scale).
    """
    from scipy.linalg import expm
    if spec:
        dt_d, b, h = sim_dt, wingspan, 100.0
    else:
        dt_d, b, h = float(sim_length), sim_dt, wingspan
    V_a = 25.0
    meters2feet = 3.281
    feet2meters = 1 / meters2feet
    knots2mpers = 0.5144
    W_20 = {"light": 15, "moderate": 30, "severe": 45, None: 15}[intensity] * knots2mpers
    h, b, V_a, W_20 = h * meters2feet, b * meters2feet, V_a * meters2feet, W_20 * meters2feet
    sigma_w = 0.1 * W_20
    sigma_u = sigma_w / (0.177 + 0.000823 * h) ** 0.4
    sigma_v = sigma_u
    L_u = h / (0.177 + 0.000823 * h) ** 1.2
    L_v, L_w = L_u, h
    K_u = sigma_u * math.sqrt((2 * L_u) / (math.pi * V_a))
    K_v = sigma_v * math.sqrt(L_v / (math.pi * V_a))
    K_w = sigma_w * math.sqrt(L_w / (math.pi * V_a))
    T_u = L_u / V_a
    T_v1, T_v2 = math.sqrt(3.0) * L_v / V_a, L_v / V_a
    T_w1, T_w2 = math.sqrt(3.0) * L_w / V_a, L_w / V_a
    K_p = sigma_w * math.sqrt(0.8 / V_a) * ((math.pi / (4 * b)) ** (1 / 6)) / (L_w ** (1 / 3))
    K_q = K_r = 1 / V_a
    T_p = 4 * b / (math.pi * V_a)
    T_q, T_r = T_p, 3 * b / (math.pi * V_a)
    tfs = [([feet2meters * K_u], [T_u, 1], 0),
           ([feet2meters * K_v * T_v1, feet2meters * K_v], [T_v2 ** 2, 2 * T_v2, 1], 1),
           ([feet2meters * K_w * T_w1, feet2meters * K_w], [T_w2 ** 2, 2 * T_w2, 1], 2),
           ([K_p], [T_p, 1], 3),
           ([-K_w * K_q * T_w1, -K_w * K_q, 0], [T_q * T_w2 ** 2, T_w2 ** 2 + 2 * T_q * T_w2, T_q + 2 * T_w2, 1], 1),
           ([K_v * K_r * T_v1, K_v * K_r, 0], [T_r * T_v2 ** 2, T_v2 ** 2 + 2 * T_r * T_v2, T_r + 2 * T_v2, 1], 2)]
    # lsim time grid: np.linspace(0, L*dt_d, L) (dryden.py:205) -> step L*dt_d/(L-1)
    step = sim_length * dt_d / (sim_length - 1)
    out = []
    for num, den, row in tfs:
        A, B, C, D = _tf2ss(num, den)
        n = A.shape[0]
        M = np.zeros((n + 2, n + 2))
        M[:n, :n] = A * step
        M[:n, n:n + 1] = B * step
        M[n, n + 1] = 1.0
        eM = expm(M.T)
        Ad = eM[:n, :n]
        Bd1 = eM[n + 1:, :n]
        Bd0 = eM[n:n + 1, :n] - Bd1
        # second and later blocks: lsim is called with T[0] = block_start_time > 0 and first steps the carried state
        # forward over [0, T[0]] with zero input, i.e. multiplies it by expm(A^T T[0]) (= Ablk^m for block m)
        Ablk = expm(A.T * (sim_length * dt_d))
        out.append((n, row, Ad, Bd0.ravel(), Bd1.ravel(), C.ravel(), float(D), Ablk))
    return out, math.sqrt(math.pi / dt_d)


def resolve_configs(env_cfg=None, sim_cfg=None, config_kw=None, sim_config_kw=None, env_kind="attitude"):
    """The reference-format (env, sim) config dicts after the constructor overrides — what `FixedWingAircraft.cfg` and
    `PyFly.cfg` hold in the reference (fixed_wing.py:34-62, pyfly.py:1067-1073).  `env_cfg` / `sim_cfg`: dict or JSON
    path (defaults above)."""
    def load(x, default):
        if x is None:
            return default()
        if isinstance(x, str):
            with open(x) as f:
                return json.load(f)
        return copy.deepcopy(x)

    env = load(env_cfg, default_env_config)
    sim = load(sim_cfg, default_sim_config)
    waypoint = env_kind == "waypoint"
    if waypoint:
        # FixedWingAircraft_simple builds PyFly from pyfly_config.json as is (simple_train.py:219-222): no gym overrides of
        # the simulator, 500 steps, commands passed straight through, 12 raw states observed
        env["steps_max"] = 500
        env["action"]["scale_space"] = False
        env["simulator"] = {"states": []}
    if config_kw:
        apply_overrides(env, copy.deepcopy(config_kw))
    sim_kw = copy.deepcopy(sim_config_kw) if sim_config_kw else {}
    if not waypoint:
        sim_kw["turbulence_sim_length"] = env["steps_max"]    # fixed_wing.py:62
    apply_overrides(sim, sim_kw)
    return env, sim


def build_config(env_cfg=None, sim_cfg=None, config_kw=None, sim_config_kw=None, params=None,
                 curriculum_level=1.0, precision="f64", integrator="rk45", rk4_substeps=4, rtol=1e-3, atol=1e-6,
                 seed=0, env_id_offset=0, dryden_spec=False, obs_init_noise=None, rng_u_override=None, env_kind="attitude"):
    """Flatten reference-format configs into an `FwConfig`.  `env_cfg` / `sim_cfg`: dict or JSON path (defaults above)."""
    env, sim = resolve_configs(env_cfg, sim_cfg, config_kw, sim_config_kw, env_kind)
    waypoint = env_kind == "waypoint"
    for key in env.get("simulator", {}):
        # sample_simulator_parameters (fixed_wing.py:748-813) re-draws aircraft parameters ("model", supported below) or
        # simulator attributes (any other key, e.g. a list of turbulence intensities) at every reset; the latter is not
        # implemented: refuse it instead of silently ignoring it
        if key not in ("states", "model"):
            raise NotImplementedError("simulator.%s: per-episode randomisation of simulator attributes "
                                      "(fixed_wing.py:801-813)" % key)
    P = load_aircraft_parameters() if params is None else dict(params)

    c = FwConfig()
    c.abi_version = FW_ABI_VERSION
    c.precision = {"f64": FW_F64, "f32": FW_F32}[precision]
    c.integrator = {"rk45": FW_INT_RK45_SCIPY, "rk4": FW_INT_RK4_FIXED}[integrator]
    c.rk4_substeps, c.rtol, c.atol = int(rk4_substeps), float(rtol), float(atol)
    for name in _AERO:
        setattr(c, name, float(P["e" if name == "e_oswald" else name]))
    c.dt, c.rho, c.g = float(sim["dt"]), float(sim["rho"]), float(sim["g"])

    # ---- simulator variables (Variable.__init__ radians conversion pyfly.py:61-64) ----
    var = {}
    for v in sim["variables"]:
        v = dict(v)
        if v.get("convert_to_radians"):
            for k in ("value_min", "value_max", "init_min", "init_max", "constraint_min", "constraint_max"):
                if v.get(k) is not None:
                    v[k] = float(np.radians(v[k]))
        var[v["name"]] = v
    # gym "simulator.states" overrides with curriculum scaling (fixed_wing.py:343-362)
    for s in env.get("simulator", {}).get("states", []):
        s = dict(s)
        name = s.pop("name")
        rad = s.pop("convert_to_radians", False)
        for prop, val in s.items():
            if val is not None:
                if "constraint" not in prop and any(m in prop for m in ("min", "max")):
                    mid = (s[prop[:-3] + "max"] + s[prop[:-3] + "min"]) / 2
                    val = mid - curriculum_level * (mid - val)
                if rad:
                    val = float(np.radians(val))
            var[name][prop] = val
    el = var["elevon_right"]
    c.elevon_min, c.elevon_max = el["value_min"], el["value_max"]
    c.elevon_dot_max = el["dot_max"]            # not degree-converted in the reference
    c.elevon_omega0, c.elevon_zeta = el["omega_0"], el["zeta"]
    th = var["throttle"]
    c.throttle_min, c.throttle_max, c.throttle_tau = th["value_min"], th["value_max"], th["tau"]
    for i, n in enumerate(("omega_p", "omega_q", "omega_r")):
        c.omega_con_min[i] = var[n].get("constraint_min") if var[n].get("constraint_min") is not None else -np.inf
        c.omega_con_max[i] = var[n].get("constraint_max") if var[n].get("constraint_max") is not None else np.inf
    c.va_value_min = var["Va"].get("value_min") or 0.0
    c.va_con_max = var["Va"].get("constraint_max") or 0.0
    for i, n in enumerate(INIT_STATES):
        c.init_lo[i], c.init_hi[i] = var[n]["init_min"], var[n]["init_max"]
    c.wind_mag_min, c.wind_mag_max = sim["wind_magnitude_min"], sim["wind_magnitude_max"]
    c.turbulence = int(bool(sim["turbulence"]))
    intensity = sim.get("turbulence_intensity")
    if intensity in ("None", "none"):
        intensity = None
    L_turb = int(sim["turbulence_sim_length"])
    filters, scale = dryden_filters(L_turb, c.dt, c.b, intensity, spec=dryden_spec)
    c.turb_noise_scale = scale
    for fi, (n, row, Ad, Bd0, Bd1, C, D, Ablk) in enumerate(filters):
        f = c.filt[fi]
        f.order, f.noise_row, f.D = n, row, D
        for a in range(n):
            f.Bd0[a], f.Bd1[a], f.C[a] = Bd0[a], Bd1[a], C[a]
            for b_ in range(n):
                f.Ad[a * n + b_] = Ad[a, b_]
                f.Ablk[a * n + b_] = Ablk[a, b_]

    # ---- gym env ----
    c.steps_max = int(env["steps_max"])
    act = env["action"]
    c.scale_actions = int(bool(act.get("scale_space", False)))
    c.scale_low, c.scale_high = float(act.get("scale_low", -1)), float(act.get("scale_high", 1))
    # Actuation.finalize (pyfly.py:599-623): elevator/aileron limits from the elevon limits
    lo = [(c.elevon_min + c.elevon_min) / 2, (-c.elevon_max + c.elevon_min) / 2, c.throttle_min]
    hi = [(c.elevon_max + c.elevon_max) / 2, (-c.elevon_min + c.elevon_max) / 2, c.throttle_max]
    for j in range(3):
        c.act_lo[j], c.act_hi[j] = lo[j], hi[j]
    mult = act.get("bounds_multiplier")
    c.has_action_bounds = int(mult is not None)
    for j in range(3):
        c.action_bounds_max[j] = act.get("scale_high", 1) * (mult or 0)
        c.action_bounds_min[j] = act.get("scale_low", -1) * (mult or 0)
    tgt = env["target"]
    tstates = {s["name"]: s for s in tgt["states"]}
    # target class attitude_angular (fixed_wing.py:671-675, 741-746): omega_p/q/r as derived target states.  The reference
    # needs all three (sample_target fills all three, _attitude_to_angular_rates reads the props of each)
    ang = [n_ for n_ in ANGULAR_TARGET_STATES if n_ in tstates]
    extra = [n_ for n_ in tstates if n_ not in ALL_TARGET_STATES]
    if extra:
        raise NotImplementedError("target states %r" % extra)
    if ang and (len(ang) != 3 or any(tstates[n_].get("class") != "attitude_angular" for n_ in ang)):
        raise NotImplementedError("omega_p / omega_q / omega_r targets: all three, of class attitude_angular")
    c.ang_on = int(bool(ang))
    for a, n_ in enumerate(ANGULAR_TARGET_STATES):
        p_ = tstates.get(n_, {})
        c.ang_max_vel[a] = float(p_.get("max_vel", np.radians(180)))
        c.ang_bound[a] = np.inf if p_.get("bound") is None else float(p_["bound"])      # raw props: no degree conversion
    for k, name in enumerate(TARGET_STATES):
        s = tstates[name]
        rad = s.get("convert_to_radians", False)

        def cur(key, v):       # set_curriculum_level target scaling (fixed_wing.py:371-387)
            if key == "low":
                mid = (s["high"] + v) / 2
            elif key == "high":
                mid = (v + s["low"]) / 2
            else:
                mid = 0
            return mid - curriculum_level * (mid - v)

        low, high = cur("low", s["low"]), cur("high", s["high"])
        delta = s.get("delta")
        delta = cur("delta", delta) if delta is not None else None
        bound = s.get("bound")
        if rad:
            low, high = float(np.radians(low)), float(np.radians(high))
            delta = float(np.radians(delta)) if delta is not None else None
            bound = float(np.radians(bound)) if bound is not None else None
        c.tgt_low[k], c.tgt_high[k] = low, high
        c.tgt_delta[k] = np.nan if delta is None else delta
        c.tgt_bound[k] = np.inf if bound is None else bound
        cls = s.get("class", "constant")
        if cls not in ("constant", "compensate", "linear", "sinusoidal"):
            raise NotImplementedError("target class %r for state %r" % (cls, name))
        if cls == "compensate" and name != "Va":
            raise NotImplementedError("target class compensate is only defined for Va (fixed_wing.py:1432-1435)")
        c.tgt_class[k] = {"constant": 0, "compensate": 1, "linear": 2, "sinusoidal": 3}[cls]
        c.tgt_radians[k] = int(bool(rad))
        if cls == "linear":
            c.tgt_slope_low[k], c.tgt_slope_high[k] = cur("slope_low", s["slope_low"]), cur("slope_high", s["slope_high"])
        if cls == "sinusoidal":
            c.tgt_amp_low[k] = cur("amplitude_low", s["amplitude_low"])
            c.tgt_amp_high[k] = cur("amplitude_high", s["amplitude_high"])
            c.tgt_period_low[k] = cur("period_low", s["period_low"]) if s.get("period_low") is not None else 250.0
            c.tgt_period_high[k] = cur("period_high", s["period_high"]) if s.get("period_high") is not None else 500.0
    if c.tgt_class[2] == 1 and c.tgt_class[1] not in (0, 2, 3):
        raise NotImplementedError("Va compensate needs a constant / linear / sinusoidal pitch target")
    c.rng_u_override = float("nan") if rng_u_override is None else float(rng_u_override)
    c.on_success = {"none": 0, "done": 1, "new": 2}[tgt.get("on_success", "none")]
    c.streak_req = int(tgt["success_streak_req"])
    if c.streak_req > 128:
        raise NotImplementedError("success_streak_req > 128")
    c.streak_fraction = float(tgt["success_streak_fraction"])
    c.resample_every = int(tgt.get("resample_every", 0) or 0)
    rew = env["reward"]
    fn_id = {"linear": 0, "exponential": 1, "quadratic": 2}
    state_names = {"roll": 0, "pitch": 1, "Va": 2, "omega_p": 3, "omega_q": 4, "omega_r": 5, "alpha": 6, "beta": 7}
    for k in range(3):
        c.rew_err_scaling[k], c.rew_err_max[k] = 0.0, np.inf
    c.rew_delta_scaling = c.rew_bound_scaling = 0.0
    c.rew_delta_max = c.rew_bound_max = np.inf
    c.rew_delta_window = 5
    default_family = (rew.get("form", "absolute") == "absolute" and len(rew["terms"]) == 1
                      and rew["terms"][0]["weight"] == 1 and rew["terms"][0]["function_class"] == "linear")
    seen = set()
    for f in rew["factors"]:
        key = (f["class"], f.get("type"), f.get("name"))
        simple = f["function_class"] == "linear" and f.get("sign", -1) < 0 and key not in seen and (
            (f["class"] == "state" and f.get("type") == "error" and f.get("name") in TARGET_STATES)
            or (f["class"] == "action" and f.get("type") in ("delta", "bound")))
        if f["class"] == "action":
            key = (f["class"], f.get("type"), None)
            simple = simple and key not in seen
        seen.add(key)
        default_family = default_family and simple
    default_family = default_family and not c.ang_on      # angular targets live in the general env head only
    c.rew_generic = int(not default_family)
    if default_family:
        for f in rew["factors"]:
            mx = f.get("max")
            mx = np.inf if mx is None else float(mx)
            if f["class"] == "state":
                k = TARGET_STATES.index(f["name"])
                c.rew_err_scaling[k], c.rew_err_max[k] = float(f["scaling"]), mx
            elif f["type"] == "delta":
                c.rew_delta_scaling, c.rew_delta_max, c.rew_delta_window = float(f["scaling"]), mx, int(f["window_size"])
            else:
                c.rew_bound_scaling, c.rew_bound_max = float(f["scaling"]), mx
    else:
        # the general engine of fixed_wing.py:941-1111
        if len(rew["factors"]) > 12 or len(rew["terms"]) > 3:
            raise NotImplementedError("more than 12 reward factors / 3 terms")
        c.rew_potential = int(rew.get("form", "absolute") == "potential")
        c.rew_nterms = len(rew["terms"])
        term_classes = []
        for t, term in enumerate(rew["terms"]):
            c.term_fclass[t], c.term_weight[t] = fn_id[term["function_class"]], float(term["weight"])
            term_classes.append(term["function_class"])
        c.rew_n = len(rew["factors"])
        for i, f in enumerate(rew["factors"]):
            if f["function_class"] not in term_classes:
                raise KeyError("reward factor %r uses function_class %r that has no term" % (f.get("name"), f["function_class"]))
            cls, typ = f["class"], f.get("type")
            c.rew_fclass[i] = fn_id[f["function_class"]]
            c.rew_scaling[i] = float(f["scaling"])
            c.rew_maxv[i] = np.inf if f.get("max") is None else float(f["max"])
            c.rew_shaping[i] = int(bool(f.get("shaping", False)))
            c.rew_sign[i] = float(np.sign(f.get("sign", -1)))
            c.rew_window[i] = int(f.get("window_size", 0) or 0)
            if cls == "state" and typ == "error":
                c.rew_class[i], c.rew_idx[i] = 0, _target_index(f["name"], c)
            elif cls == "state" and typ == "int_error":
                c.rew_class[i], c.rew_idx[i] = 9, _target_index(f["name"], c)
            elif cls == "state" and typ == "value":
                if f["name"] not in state_names:
                    raise NotImplementedError("reward on state %r" % f["name"])
                c.rew_class[i], c.rew_idx[i] = 1, state_names[f["name"]]
            elif cls == "action" and typ == "value":
                c.rew_class[i] = 2
            elif cls == "action" and typ == "delta":
                c.rew_class[i] = 3
                if c.rew_window[i] > 5:
                    raise NotImplementedError("reward action-delta window > 5")
                c.rew_delta_window = max(c.rew_delta_window, c.rew_window[i])
            elif cls == "action" and typ == "bound":
                c.rew_class[i] = 4
            elif cls == "success":
                c.rew_class[i] = 5
                c.rew_value_timesteps[i] = int(f["value"] == "timesteps")
                c.rew_value[i] = 0.0 if f["value"] == "timesteps" else float(f["value"])
            elif cls == "step":
                c.rew_class[i], c.rew_value[i] = 6, float(f["value"])
            elif cls == "goal" and typ == "per_state" and c.ang_on and not all(np.isfinite(c.ang_bound[a]) for a in range(3)):
                # the reference indexes history["goal"][state] for EVERY target state here (fixed_wing.py:1038-1044)
                raise KeyError("goal per_state reward needs a bound on every target state, omega_p / omega_q / omega_r included")
            elif cls == "goal" and typ in ("per_state", "all"):
                c.rew_class[i], c.rew_value[i] = (7 if typ == "per_state" else 8), float(f["value"])
            else:
                raise NotImplementedError("reward factor %r" % (f,))
    fail = rew.get("step_fail", 0)
    c.step_fail_timesteps = int(fail == "timesteps")
    c.step_fail_value = 0.0 if fail == "timesteps" else float(fail)
    obs = env["observation"]
    noise = obs.get("noise") or {}
    c.obs_noise_mean, c.obs_noise_std = float(noise.get("mean", 0) or 0), float(noise.get("var", 0) or 0)
    L = int(obs.get("length", 1))
    ostep = int(obs.get("step", 1) or 1)
    states = obs["states"]
    if L > 5 or len(states) > 16:
        raise NotImplementedError("observation.length > 5 or more than 16 entries per row")
    # row k has lag 1 + k * step (range(1, length * step, step), fixed_wing.py:1129-1138); history rows reach 5 steps back
    lag_max = 1 + (L - 1) * ostep
    if ostep < 1 or lag_max > 5:
        raise NotImplementedError("observation rows reach %d steps back (length %d, step %d): at most 5" % (lag_max, L, ostep))
    c.obs_step = ostep
    W = int(env.get("integration_window", 0) or 0)
    c.integration_window = W
    uses_int = any(s_["type"] == "target" and s_.get("value") == "integrator" for s_ in states) or any(
        f.get("class") == "state" and f.get("type") == "int_error" for f in rew["factors"])
    if uses_int and (W < 0 or W + lag_max > 49):
        raise NotImplementedError("integration_window %d + observation lag %d exceed the 50-deep error ring" % (W, lag_max))
    names = [(s_["name"], s_["type"]) for s_ in states]
    default_layout = [("roll", "state"), ("pitch", "state"), ("Va", "state"), ("omega_p", "state"),
                      ("omega_q", "state"), ("omega_r", "state"), ("roll", "target"), ("pitch", "target"),
                      ("Va", "target"), ("alpha", "state"), ("beta", "state"), ("elevator", "action"),
                      ("aileron", "action"), ("throttle", "action")]
    c.obs_normalize = int(bool(obs.get("normalize", False)))
    c.obs_len, c.obs_n = L, len(states)
    c.obs_generic = int(not (L == 1 and ostep == 1 and names == default_layout and not c.obs_normalize
                             and all(s_.get("value", "absolute") == "absolute" for s_ in states if s_["type"] == "target")))
    c.obs_init_noise = float("nan") if obs_init_noise is None else float(obs_init_noise)
    state_idx = {"roll": 0, "pitch": 1, "Va": 2, "omega_p": 3, "omega_q": 4, "omega_r": 5, "alpha": 6, "beta": 7}
    act_names = [a_["name"] for a_ in act["states"]]
    f32max = float(np.finfo(np.float32).max)
    windows = [1]
    for e, s_ in enumerate(states):
        kind = s_["type"]
        if kind == "state":
            if s_["name"] not in state_idx:
                raise NotImplementedError("observation of state %r" % s_["name"])
            c.obs_kind[e], c.obs_idx[e] = 0, state_idx[s_["name"]]
        elif kind == "target":
            value = s_.get("value", "absolute")
            if value not in ("absolute", "relative", "integrator"):
                raise ValueError("Unexpected observation variable target value type: %r" % value)
            c.obs_kind[e], c.obs_idx[e] = {"absolute": 1, "relative": 2, "integrator": 4}[value], _target_index(s_["name"], c)
        elif kind == "action":
            c.obs_kind[e], c.obs_idx[e] = 3, act_names.index(s_["name"])
            c.obs_window[e] = int(s_.get("window_size", 1))
            windows.append(c.obs_window[e])
        else:
            raise NotImplementedError("observation entry type %r" % kind)
        # normalisation constants (fixed_wing.py:95-161); limits come from the entry or the pyfly variable
        v = var.get(s_["name"], {})
        hi, lo = s_.get("high"), s_.get("low")
        if hi is None:
            hi = v.get("value_max") if v.get("value_max") is not None else v.get("constraint_max")
            hi = f32max if hi is None else hi
        elif s_.get("convert_to_radians"):
            hi = float(np.radians(hi))
        if lo is None:
            lo = v.get("value_min") if v.get("value_min") is not None else v.get("constraint_min")
            lo = -f32max if lo is None else lo
        elif s_.get("convert_to_radians"):
            lo = float(np.radians(lo))
        finite = hi != f32max and lo != -f32max
        c.obs_mean[e] = s_["mean"] if s_.get("mean") is not None else ((hi - lo) if finite else 0.0)
        c.obs_var[e] = s_["var"] if s_.get("var") is not None else ((hi - lo) / 16 if finite else 1.0)
        c.obs_norm_flag[e] = int(bool(s_.get("norm", True)))
    c.obs_act_window = max(windows)
    if max(c.obs_act_window, c.rew_delta_window) + (lag_max - 1) > 9:
        raise NotImplementedError("action windows + observation length exceed the 8-deep action ring")
    c.rise_low, c.rise_high = 0.1, 0.9
    for m in env.get("metrics", []):
        if m["name"] == "rise_time":
            c.rise_low, c.rise_high = m.get("low", 0.1), m.get("high", 0.9)
    c.env_kind = 1 if waypoint else 0
    c.turb_block_len = L_turb
    for k in range(3):
        c.wp_goal_bound[k], c.wp_rew_range[k] = 0.5, 6.0
    c.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    c.env_id_offset = int(env_id_offset)
    # ---- simulator.model: per-episode aircraft-parameter randomisation (fixed_wing.py:758-800) ----
    model = env.get("simulator", {}).get("model")
    for i, name in enumerate(_AERO):
        c.par_orig[i] = getattr(c, name)
        c.par_clip[i] = float("nan")
    if model is not None:
        if waypoint:
            raise NotImplementedError("simulator.model with the waypoint env")
        dist = model.get("distribution", "gaussian")
        if dist not in ("gaussian", "uniform"):
            raise ValueError("Unexpected distribution type {}".format(dist))
        c.model_on, c.model_uniform = 1, int(dist == "uniform")
        relative = model["var_type"] == "relative"
        for prm in model["parameters"]:
            key = "e_oswald" if prm["name"] == "e" else prm["name"]
            if key not in _AERO:
                if prm["name"] in P:        # C_D_0, C_D_alpha1/2, r_cg, ...: in x8_param.mat, read by nothing in the dynamics
                    continue
                raise KeyError(prm["name"])
            i = _AERO.index(key)
            orig = prm.get("original")
            orig = getattr(c, key) if orig is None else float(orig)
            var = float(prm.get("var", model["var"]))
            clip = prm.get("clip", model.get("clip"))
            if relative:
                var *= abs(orig)
                clip = None if clip is None else clip * orig      # sic: not abs (fixed_wing.py:785-786)
            c.par_enabled[i], c.par_orig[i], c.par_var[i] = 1, orig, var
            c.par_clip[i] = float("nan") if clip is None else float(clip)
    return c


def obs_dim(cfg):
    if cfg.env_kind == 1:
        return 12
    return cfg.obs_len * cfg.obs_n if cfg.obs_generic else FW_NOBS


def _simulator_limits(sim):
    """name -> {value_min, value_max, constraint_min, constraint_max} of the pyfly states as PyFly.__init__ leaves them
    (Variable.__init__ degree conversion pyfly.py:61-64; elevator / aileron limits from the elevons, Actuation.finalize
    pyfly.py:599-623) — before the gym config's simulator overrides, which set_curriculum_level applies after the spaces
    are built (fixed_wing.py:412)."""
    var = {}
    for v in sim["variables"]:
        v = dict(v)
        if v.get("convert_to_radians"):
            for k in ("value_min", "value_max", "constraint_min", "constraint_max"):
                if v.get(k) is not None:
                    v[k] = float(np.radians(v[k]))
        var[v["name"]] = v
    if "elevon_right" in var and "elevon_left" in var:
        er, el = var["elevon_right"], var["elevon_left"]
        if None not in (er.get("value_min"), er.get("value_max"), el.get("value_min"), el.get("value_max")):
            var["elevator"] = {"value_min": (er["value_min"] + el["value_min"]) / 2,
                               "value_max": (er["value_max"] + el["value_max"]) / 2}
            var["aileron"] = {"value_min": (-er["value_max"] + el["value_min"]) / 2,
                              "value_max": (-er["value_min"] + el["value_max"]) / 2}
    return var


def _state_limit(var, name, lim, sign):
    st = var.get(name, {})
    if st.get("value_" + lim) is not None:
        return st["value_" + lim]
    if st.get("constraint_" + lim) is not None:
        return st["constraint_" + lim]
    return sign * float(np.finfo(np.float32).max)


def action_space_bounds(env, sim):
    """(low, high) of `FixedWingAircraft.action_space`, float32 (fixed_wing.py:199-258): per action state "max" ->
    +-float32 max, absent -> the simulator state's limit, else the number given."""
    f32max = float(np.finfo(np.float32).max)
    var = _simulator_limits(sim)
    lo, hi = [], []
    for a in env["action"]["states"]:
        h, l = a.get("high"), a.get("low")
        hi.append(f32max if h == "max" else _state_limit(var, a["name"], "max", 1.0) if h is None else h)
        lo.append(-f32max if l == "max" else _state_limit(var, a["name"], "min", -1.0) if l is None else l)
    return np.array(lo, dtype=np.float64).astype(np.float32), np.array(hi, dtype=np.float64).astype(np.float32)


def observation_space_bounds(env, sim):
    """(low, high) of `FixedWingAircraft.observation_space` for ANY observation layout, float32 and shaped like the
    reference's Box: [n] per entry, tiled `length` times for shape "vector", [length, n] for shape "matrix"
    (fixed_wing.py:92-140, 178-190, 245-247).  `env` / `sim`: the reference-format dicts of resolve_configs.  An entry's
    bound is its own "low" / "high" (degrees converted if it says so), else the simulator state's value limit, else its
    constraint, else +-float32 max; a relative target spans high - low either way when both are finite.  The reference
    builds the space BEFORE set_curriculum_level applies the gym config's simulator overrides (fixed_wing.py:412), so
    the simulator limits are pyfly_config's own (_simulator_limits)."""
    f32max = float(np.finfo(np.float32).max)
    var = _simulator_limits(sim)

    def bound(o, key, lim, sign):
        b = o.get(key)
        if b is None:
            return _state_limit(var, o["name"], lim, sign)
        return float(np.radians(b)) if o.get("convert_to_radians", False) else b

    lo, hi = [], []
    for o in env["observation"]["states"]:
        high, low = bound(o, "high", "max", 1.0), bound(o, "low", "min", -1.0)
        if o["type"] == "target" and o.get("value") == "relative":
            if high != f32max and low != -f32max:
                high, low = high - low, low - high
            else:
                high, low = f32max, -f32max
        hi.append(high)
        lo.append(low)
    L = int(env["observation"].get("length", 1))
    if L > 1:
        shape = env["observation"].get("shape", "vector")
        if shape == "vector":
            lo, hi = lo * L, hi * L
        elif shape == "matrix":
            lo, hi = [lo] * L, [hi] * L
        else:
            raise ValueError("observation.shape %r (fixed_wing.py:178-192)" % (shape,))
    return np.array(lo, dtype=np.float64).astype(np.float32), np.array(hi, dtype=np.float64).astype(np.float32)


def observation_bounds(env_cfg=None, cfg=None):
    """observation_space low/high (fixed_wing.py:92-140) of the default 14-vector from the flattened config alone (the
    adapters use observation_space_bounds, which covers every layout); a general layout reports unbounded boxes."""
    if cfg is not None and cfg.obs_generic:
        f32max = float(np.finfo(np.float32).max)
        return (np.full(obs_dim(cfg), -f32max, np.float32), np.full(obs_dim(cfg), f32max, np.float32))
    f32max = float(np.finfo(np.float32).max)
    d = np.radians
    lo = [d(-180), d(-85), 0, d(-720), d(-720), d(-720)]
    hi = [d(180), d(85), 70, d(720), d(720), d(720)]
    # targets: pyfly state limits -> none for roll/pitch; Va has value_min 1e-6
    lo += [-f32max, -f32max, 1e-6, -f32max, -f32max]
    hi += [f32max, f32max, f32max, f32max, f32max]
    a_lo = [cfg.act_lo[j] for j in range(3)] if cfg is not None else [d(-30), d(-32.5), 0]
    a_hi = [cfg.act_hi[j] for j in range(3)] if cfg is not None else [d(35), d(32.5), 1]
    return (np.array(lo + a_lo, dtype=np.float32), np.array(hi + a_hi, dtype=np.float32))
