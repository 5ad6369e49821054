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
