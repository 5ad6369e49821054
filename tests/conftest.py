import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


# (fixture file, config_kw, sim_config_kw, actions were float32)
TRAJ_CASES = [
    ("traj_calm", None, {"turbulence": False}, False),
    ("traj_wind", None, {"turbulence": False}, False),
    ("traj_f32act", None, {"turbulence": False}, True),
    ("traj_turb", None, {"turbulence": True}, False),
    ("traj_turb_severe", None, {"turbulence": True, "turbulence_intensity": "severe"}, False),
    ("traj_turb_moderate", None, {"turbulence": True, "turbulence_intensity": "moderate"}, False),
    ("traj_fail", None, {"turbulence": False}, False),
    ("traj_full400", {"steps_max": 400}, {"turbulence": False}, False),
    ("traj_full300_turb", {"steps_max": 300}, {"turbulence": True}, False),
]
METRIC_KEYS = ("rise_time", "settling_time", "overshoot", "total_error", "avg_error", "control_variation", "success",
               "success_time_frac", "end_error")


def golden_metric_rows(g):
    """[n_done_episodes, 28] metric rows in FwMetricIndex order + the episode index of each row."""
    rows = np.concatenate([g["m_" + k] for k in METRIC_KEYS], axis=1)
    eps = [ep for ep in range(len(g["n_valid"])) if bool(g["done"][ep, int(g["n_valid"][ep]) - 1])]
    return rows, eps


def close_or_both_nan(a, b, rtol, atol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.isclose(a, b, rtol=rtol, atol=atol) | (np.isnan(a) & np.isnan(b))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return 0


def cnn_env_config():
    """examples/models/cnn_controller/fixed_wing_config.json of the reference, as a delta on the default config:
    observation length 5 / shape matrix, relative targets, no alpha/beta entries."""
    from tum_adlr_deep_reinforcement_learning_b200.config import default_env_config
    cfg = default_env_config()

    def st(name, lo=None, hi=None, rad=False):
        d = {"name": name, "type": "state"}
        if lo is not None:
            d["low"] = lo
        if hi is not None:
            d["high"] = hi
        if rad:
            d["convert_to_radians"] = True
        return d

    cfg["observation"] = {
        "length": 5, "step": 1, "shape": "matrix",
        "states": [st("roll", -180, 180, True), st("pitch", -85, 85, True), st("Va", None, 60),
                   st("omega_p", -720, 720, True), st("omega_q", -720, 720, True), st("omega_r", -720, 720, True),
                   {"name": "roll", "type": "target", "value": "relative"},
                   {"name": "pitch", "type": "target", "value": "relative"},
                   {"name": "Va", "type": "target", "value": "relative"},
                   {"name": "elevator", "type": "action", "window_size": 5},
                   {"name": "aileron", "type": "action", "window_size": 5},
                   {"name": "throttle", "type": "action", "window_size": 5}]}
    return cfg


def rich_reward_env_config():
    """The config of tests/golden/make_golden.py:gen_reward: potential form, three terms, every factor class."""
    import importlib.util
    from tum_adlr_deep_reinforcement_learning_b200.config import default_env_config
    cfg = default_env_config()
    cfg["reward"] = RICH_REWARD
    cfg["steps_max"] = 80
    cfg["target"]["success_streak_req"] = 6
    cfg["target"]["success_streak_fraction"] = 0.5
    for s_, b in zip(cfg["target"]["states"], (60, 40, 12)):
        s_["bound"] = b
    return cfg


RICH_REWARD = {
    "form": "potential", "randomize_scaling": False, "step_fail": -50,
    "terms": [{"function_class": "linear", "weight": 1.0}, {"function_class": "exponential", "weight": 0.5},
              {"function_class": "quadratic", "weight": 0.1}],
    "factors": [
        {"name": "roll", "class": "state", "type": "error", "function_class": "linear", "scaling": 3.2,
         "shaping": True, "max": 0.3, "sign": -1},
        {"name": "pitch", "class": "state", "type": "error", "function_class": "exponential", "scaling": 2.0,
         "shaping": True, "sign": -1},
        {"name": "Va", "class": "state", "type": "error", "function_class": "quadratic", "scaling": 100,
         "shaping": False, "sign": -1},
        {"name": "omega_q", "class": "state", "type": "value", "function_class": "quadratic", "scaling": 50,
         "shaping": True, "sign": -1},
        {"name": "action", "class": "action", "type": "value", "function_class": "linear", "scaling": 30,
         "shaping": False, "sign": -1},
        {"name": "action", "class": "action", "type": "delta", "function_class": "linear", "window_size": 3,
         "scaling": 60, "shaping": False, "sign": -1},
        {"name": "action_bound", "class": "action", "type": "bound", "function_class": "linear", "scaling": 1,
         "shaping": False, "sign": -1},
        {"name": "success", "class": "success", "value": "timesteps", "function_class": "linear", "scaling": 100,
         "shaping": False, "sign": 1},
        {"name": "step", "class": "step", "value": 1, "function_class": "linear", "scaling": 10, "shaping": False,
         "sign": -1},
        {"name": "goal", "class": "goal", "type": "per_state", "value": 0.3, "function_class": "linear",
         "scaling": 1, "shaping": False, "sign": 1},
        {"name": "goal_all", "class": "goal", "type": "all", "value": 1.0, "function_class": "exponential",
         "scaling": 4, "shaping": False, "sign": 1}]}


MOVING_TARGETS = [
    {"name": "roll", "convert_to_radians": True, "low": -60, "high": 60, "delta": 180, "class": "linear",
     "slope_low": 2, "slope_high": 8, "bound": 5},
    {"name": "pitch", "convert_to_radians": True, "low": -25, "high": 25, "delta": 45, "class": "sinusoidal",
     "amplitude_low": 3, "amplitude_high": 9, "period_low": 60, "period_high": 140, "bound": 5},
    {"name": "Va", "low": 15, "high": 28, "delta": 6, "class": "compensate", "bound": 2}]


def resample_env_config():
    """Default target classes with periodic and on-success resampling (fixed_wing.py:536-580): wide goal bounds so
    that the 5-step success streak is reached."""
    import copy
    from tum_adlr_deep_reinforcement_learning_b200.config import default_env_config
    cfg = copy.deepcopy(default_env_config())
    cfg["steps_max"] = 150
    cfg["target"].update(resample_every=37, on_success="new", success_streak_req=5, success_streak_fraction=0.6)
    for st, b in zip(cfg["target"]["states"], (80, 50, 15)):
        st["bound"] = b
    return cfg


def moving_targets_env_config():
    from tum_adlr_deep_reinforcement_learning_b200.config import default_env_config
    cfg = default_env_config()
    cfg["target"]["states"] = MOVING_TARGETS
    cfg["steps_max"] = 150
    return cfg


MODEL_BLOCK = {"var_type": "relative", "var": 0.1, "clip": 0.15, "distribution": "gaussian",
               "parameters": [{"name": "C_L_alpha"}, {"name": "C_m_q"}, {"name": "mass", "var": 0.05}, {"name": "C_D_p"},
                              {"name": "k_motor"}, {"name": "C_l_p"}, {"name": "C_n_r", "clip": 0.05}, {"name": "C_Y_beta"},
                              {"name": "S_prop"}, {"name": "b"}, {"name": "c"}, {"name": "M", "var": 0.02}, {"name": "e"},
                              {"name": "Jx"}, {"name": "C_L_0"}, {"name": "C_D_q"}]}


def model_env_config(distribution="gaussian"):
    """The default gym config plus the simulator.model block the fixtures traj_model_*.npz were recorded with
    (tests/golden/make_golden.py gen_model)."""
    from tum_adlr_deep_reinforcement_learning_b200.config import default_env_config
    cfg = default_env_config()
    cfg["simulator"]["model"] = dict(MODEL_BLOCK, distribution=distribution)
    return cfg


def integrator_env_config(W, L, step):
    """The configs of tests/golden/make_golden.py:gen_integrator: `integration_window` W, observation rows at lags
    1, 1 + step, ..., "integrator" target entries and "int_error" reward factors."""
    from tum_adlr_deep_reinforcement_learning_b200.config import default_env_config
    cfg = default_env_config()
    cfg["integration_window"] = W
    ob = cfg["observation"]
    ob["length"], ob["step"], ob["shape"] = L, step, "vector"
    ob["states"] = [s_ for s_ in ob["states"] if s_["name"] not in ("alpha", "beta")]
    for name in ("roll", "pitch", "Va"):
        ob["states"].append({"name": name, "type": "target", "value": "integrator"})
    for name, sc in zip(("roll", "pitch", "Va"), (40.0, 25.0, 300.0)):
        cfg["reward"]["factors"].append({"name": name, "class": "state", "type": "int_error", "function_class": "linear",
                                         "scaling": sc, "shaping": False, "max": 2.0, "sign": -1})
    cfg["steps_max"] = 60
    return cfg


INTEGRATOR_CASES = [("traj_integrator_w4", 4, 3, 2), ("traj_integrator_w0", 0, 1, 1)]


def angular_env_config():
    """The config of tests/golden/make_golden.py:gen_angular: target class attitude_angular on omega_p/q/r with bounds,
    observed, rewarded and part of the success streak."""
    from tum_adlr_deep_reinforcement_learning_b200.config import default_env_config
    cfg = default_env_config()
    names = ("omega_p", "omega_q", "omega_r")
    for name, bound in zip(names, (0.6, 0.4, 0.5)):
        cfg["target"]["states"].append({"name": name, "class": "attitude_angular", "bound": bound})
    cfg["target"]["states"][-1]["max_vel"] = 2.5
    ob = cfg["observation"]
    ob["states"] = [s_ for s_ in ob["states"] if s_["name"] not in ("alpha", "beta")]
    for name in names:
        ob["states"].append({"name": name, "type": "target", "value": "absolute"})
    ob["states"].append({"name": "omega_q", "type": "target", "value": "relative"})
    for name, sc in zip(names, (6.0, 5.0, 4.0)):
        cfg["reward"]["factors"].append({"name": name, "class": "state", "type": "error", "function_class": "linear",
                                         "scaling": sc, "shaping": False, "max": 0.5, "sign": -1})
    cfg["reward"]["factors"].append({"name": "goal", "class": "goal", "type": "per_state", "value": 0.6,
                                     "function_class": "linear", "scaling": 1, "shaping": False, "sign": 1})
    cfg["steps_max"] = 50
    cfg["target"]["success_streak_req"] = 5
    cfg["target"]["success_streak_fraction"] = 0.6
    for s_, b in zip(cfg["target"]["states"][:3], (70, 45, 14)):
        s_["bound"] = b
    return cfg


def angular_metric_rows(metrics28, metrics24):
    """The fixture's 52-value row (5 error metrics x 6 states | 3 goal metrics x (6 states + all) | control_variation)
    from the C-ABI layout: 28 base metrics (FwMetricIndex) + 24 angular ones (metric-major)."""
    m, a = np.asarray(metrics28, dtype=np.float64), np.asarray(metrics24, dtype=np.float64)
    row = []
    base_off = {"avg_error": 13, "total_error": 10, "end_error": 25, "rise_time": 0, "overshoot": 7}     # FwMetricIndex
    for q, name in enumerate(("avg_error", "total_error", "end_error", "rise_time", "overshoot")):
        row += list(m[base_off[name]:base_off[name] + 3]) + list(a[q * 3:q * 3 + 3])
    for q, off in enumerate((17, 3, 21)):              # success, settling_time, success_time_frac: roll pitch Va all
        row += list(m[off:off + 3]) + list(a[15 + q * 3:15 + q * 3 + 3]) + [m[off + 3]]
    row.append(m[16])
    return np.array(row)
