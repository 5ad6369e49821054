"""Generate the golden fixtures in this directory by running the UNMODIFIED reference where it lies.

TEST INFRASTRUCTURE — run here (the build container, where /root/reference is mounted), never on the
GPU box.  Usage:  python tests/golden/make_golden.py [angular|integrator|params|traj|turb|fail|full|pid|gae|vecnorm|ppo_update|curriculum|dryden|all]

Everything is recorded through the reference's public surface:
  FixedWingAircraft.reset(state=, target=, turbulence_noise=) / .step(action)
      (magpie/libs/fixed-wing-gym/gym_fixed_wing/fixed_wing.py:414, :483)
  PyFly state objects (magpie/libs/pyfly/pyfly/pyfly.py:1129-1221)
  RolloutBuffer.compute_returns_and_advantage (magpie/libs/stable-baselines3/stable_baselines3/common/buffers.py:304)
Per-step RHS-evaluation counts are taken by wrapping PyFly._dynamics (pyfly.py:1450) with a counter.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.install()
from gym_fixed_wing.fixed_wing import FixedWingAircraft  # noqa: E402
from pyfly.pid_controller import PIDController  # noqa: E402

STATE_KEYS = ["roll", "pitch", "yaw", "omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d",
              "velocity_u", "velocity_v", "velocity_w"]
TARGET_KEYS = ["roll", "pitch", "Va"]


def make_env(turbulence, config_kw=None, intensity="light", config_path=None):
    env = FixedWingAircraft(config_path or refshim.GYM_CONFIG, config_kw=config_kw,
                            sim_config_kw={"turbulence": turbulence, "turbulence_intensity": intensity})
    env.seed(0)
    sim = env.simulator
    sim._nfev = 0
    inner = sim._dynamics

    def counted(t, y, control_sp=None):
        sim._nfev += 1
        return inner(t, y, control_sp)

    sim._dynamics = counted
    return env


def term_code(term):
    """0 none, 1 steps, 2 success, 10+k constraint on (omega_p, omega_q, omega_r, Va)[k]."""
    if term in ("", None):
        return 0
    if term in ("steps", "success"):
        return 1 if term == "steps" else 2
    return 10 + ["omega_p", "omega_q", "omega_r", "Va"].index(term)


def snapshot(env):
    sim = env.simulator
    y = list(sim.state["attitude"].value)
    y += sim.get_states_vector(["omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d",
                                "velocity_u", "velocity_v", "velocity_w"])
    y += sim.actuation.get_values()
    euler = sim.get_states_vector(["roll", "pitch", "yaw"])
    vab = sim.get_states_vector(["Va", "alpha", "beta"])
    cmd = [sim.state[k].command for k in ("elevator", "aileron", "throttle")]
    cmd = [np.nan if c is None else c for c in cmd]
    tgt = [env.target[k] for k in TARGET_KEYS]
    return (np.array(y, dtype=np.float64), np.array(euler, dtype=np.float64), np.array(vab, dtype=np.float64),
            np.array(cmd, dtype=np.float64), np.array(tgt, dtype=np.float64))


def random_scenario(rs, wind_mag=0.0, hard=False):
    """Initial state drawn from the default-task init ranges, as an injectable dict."""
    d = np.radians
    om = 60 if not hard else 600
    st = {
        "roll": rs.uniform(d(-110), d(110)), "pitch": rs.uniform(d(-45), d(45)), "yaw": rs.uniform(d(-30), d(30)),
        "omega_p": rs.uniform(d(-om), d(om)), "omega_q": rs.uniform(d(-om), d(om)), "omega_r": rs.uniform(d(-om), d(om)),
        "position_n": rs.uniform(-100, 100), "position_e": rs.uniform(-100, 100), "position_d": rs.uniform(-100, -20),
        "velocity_u": rs.uniform(10, 23) if not hard else rs.uniform(40, 60),
        "velocity_v": rs.uniform(-5, 5), "velocity_w": rs.uniform(-5, 5),
        "elevon_right": (rs.uniform(d(-20), d(20)), rs.uniform(-1, 1)),
        "elevon_left": (rs.uniform(d(-20), d(20)), rs.uniform(-1, 1)),
        "throttle": (rs.uniform(0, 1), 0.0),
    }
    if wind_mag > 0:
        w = rs.uniform(-wind_mag, wind_mag, 3)
    else:
        w = np.zeros(3)
    st["wind"] = [float(w[0]), float(w[1]), float(w[2])]
    tgt = {"roll": rs.uniform(d(-60), d(60)), "pitch": rs.uniform(d(-25), d(25)), "Va": rs.uniform(15, 28)}
    return st, tgt


def scenario_arrays(st, tgt):
    s = [st[k] for k in STATE_KEYS]
    s += [st["elevon_right"][0], st["elevon_left"][0], st["throttle"][0],
          st["elevon_right"][1], st["elevon_left"][1], st["throttle"][1]]
    s += list(st["wind"])
    return np.array(s, dtype=np.float64), np.array([tgt[k] for k in TARGET_KEYS], dtype=np.float64)


def run_episodes(env, n_ep, n_steps, rs, turbulence, wind_mag, action_amp, hard=False, noise_len=None,
                 metrics=False, f32_actions=False, after_reset=None):
    """Roll `n_ep` injected episodes; every record has a leading [episode, step] shape."""
    rec = {k: [] for k in ("init_state", "init_target", "actions", "y", "euler", "vab", "cmd", "target", "obs",
                           "reward", "done", "nfev", "term", "obs0", "y0", "euler0", "vab0", "n_valid")}
    if turbulence:
        rec.update({"noise": [], "turb_lin": [], "turb_ang": []})
    mrec = []
    L = env.steps_max if noise_len is None else noise_len
    for ep in range(n_ep):
        st, tgt = random_scenario(rs, wind_mag, hard)
        kw = {}
        if turbulence:
            noise = rs.standard_normal((4, L))
            kw["turbulence_noise"] = noise
        obs0 = env.reset(state=dict(st), target=dict(tgt), **kw)
        if after_reset is not None:
            rec.setdefault("params", []).append(after_reset(env))
        s_arr, t_arr = scenario_arrays(st, tgt)
        rec["init_state"].append(s_arr)
        rec["init_target"].append(t_arr)
        rec["obs0"].append(np.array(obs0, dtype=np.float64))
        y0, e0, v0, _, _ = snapshot(env)
        rec["y0"].append(y0), rec["euler0"].append(e0), rec["vab0"].append(v0)
        if turbulence:
            rec["noise"].append(noise)
            rec["turb_lin"].append(np.array(env.simulator.wind.dryden.vel_lin))
            rec["turb_ang"].append(np.array(env.simulator.wind.dryden.vel_ang))
        ep_rec = {k: [] for k in ("actions", "y", "euler", "vab", "cmd", "target", "obs", "reward", "done", "nfev",
                                  "term")}
        n_valid = n_steps
        for t in range(n_steps):
            a = rs.uniform(-action_amp, action_amp, 3)
            if f32_actions:
                a = a.astype(np.float32)
            env.simulator._nfev = 0
            obs, rew, done, info = env.step(a)
            y, e, v, c, tg = snapshot(env)
            ep_rec["actions"].append(np.asarray(a, dtype=np.float64))
            ep_rec["y"].append(y), ep_rec["euler"].append(e), ep_rec["vab"].append(v), ep_rec["cmd"].append(c)
            ep_rec["target"].append(tg), ep_rec["obs"].append(np.array(obs, dtype=np.float64))
            ep_rec["reward"].append(float(rew)), ep_rec["done"].append(bool(done))
            ep_rec["nfev"].append(env.simulator._nfev)
            ep_rec["term"].append(term_code(info.get("termination", "")))
            if done:
                n_valid = t + 1
                if metrics:
                    mrec.append({k: info[k] for k in ("rise_time", "settling_time", "overshoot", "total_error",
                                                      "avg_error", "control_variation", "success",
                                                      "success_time_frac", "end_error")})
                break
        rec["n_valid"].append(n_valid)
        for k, v in ep_rec.items():
            arr = np.array(v)
            pad = [(0, n_steps - arr.shape[0])] + [(0, 0)] * (arr.ndim - 1)
            rec[k].append(np.pad(arr, pad))
    out = {k: np.array(v) for k, v in rec.items()}
    if metrics:
        out.update(pack_metrics(mrec))
    return out


def pack_metrics(mrec):
    def f(x):
        return np.nan if x is None else float(x)

    out = {}
    for name, keys in (("rise_time", TARGET_KEYS), ("settling_time", TARGET_KEYS + ["all"]),
                       ("overshoot", TARGET_KEYS), ("total_error", TARGET_KEYS), ("avg_error", TARGET_KEYS),
                       ("control_variation", ["all"]), ("success", TARGET_KEYS + ["all"]),
                       ("success_time_frac", TARGET_KEYS + ["all"]), ("end_error", TARGET_KEYS)):
        out["m_" + name] = np.array([[f(m[name][k]) for k in keys] for m in mrec], dtype=np.float64)
    return out


def gen_params():
    import scipy.io
    p = scipy.io.loadmat(refshim.X8_PARAMS, squeeze_me=True)
    params = {k: float(v) for k, v in p.items() if not k.startswith("__") and np.size(v) == 1}
    dst = os.path.join(ROOT, "tum_adlr_deep_reinforcement_learning_b200", "data", "x8_params.json")
    with open(dst, "w") as f:
        json.dump(params, f, indent=1, sort_keys=True)
    print("wrote", dst, len(params), "scalars")


def gen_traj():
    rs = np.random.RandomState(1234)
    env = make_env(False)
    out = run_episodes(env, 6, 250, rs, False, wind_mag=0.0, action_amp=1.3)
    np.savez_compressed(os.path.join(HERE, "traj_calm.npz"), **out)
    rs = np.random.RandomState(99)
    out = run_episodes(env, 4, 250, rs, False, wind_mag=6.0, action_amp=2.0)
    np.savez_compressed(os.path.join(HERE, "traj_wind.npz"), **out)
    rs = np.random.RandomState(5)
    out = run_episodes(env, 3, 120, rs, False, wind_mag=3.0, action_amp=1.2, f32_actions=True)
    np.savez_compressed(os.path.join(HERE, "traj_f32act.npz"), **out)


def gen_turb():
    rs = np.random.RandomState(4321)
    env = make_env(True)
    out = run_episodes(env, 5, 250, rs, True, wind_mag=6.0, action_amp=1.3)
    np.savez_compressed(os.path.join(HERE, "traj_turb.npz"), **out)
    env = make_env(True, intensity="severe")
    rs = np.random.RandomState(777)
    out = run_episodes(env, 2, 200, rs, True, wind_mag=4.0, action_amp=1.0)
    np.savez_compressed(os.path.join(HERE, "traj_turb_severe.npz"), **out)


def gen_turb_moderate():
    """Moderate Dryden intensity (the third setting of pyfly_config.json's turbulence_intensity) with injected noise."""
    env = make_env(True, intensity="moderate")
    rs = np.random.RandomState(2468)
    out = run_episodes(env, 2, 160, rs, True, wind_mag=5.0, action_amp=1.1)
    np.savez_compressed(os.path.join(HERE, "traj_turb_moderate.npz"), **out)


MODEL_BLOCK = {"var_type": "relative", "var": 0.1, "clip": 0.15, "distribution": "gaussian",
               "parameters": [{"name": "C_L_alpha"}, {"name": "C_m_q"}, {"name": "mass", "var": 0.05}, {"name": "C_D_p"},
                              {"name": "k_motor"}, {"name": "C_l_p"}, {"name": "C_n_r", "clip": 0.05}, {"name": "C_Y_beta"},
                              {"name": "S_prop"}, {"name": "b"}, {"name": "c"}, {"name": "M", "var": 0.02}, {"name": "e"},
                              {"name": "Jx"}, {"name": "C_L_0"}, {"name": "C_D_q"}]}
AERO_NAMES = ("mass Jx Jy Jz Jxz S_wing b c S_prop C_prop k_motor k_T_P k_Omega e M a_0 "
              "C_L_0 C_L_alpha C_L_q C_L_delta_e C_D_p C_D_q C_D_beta1 C_D_beta2 C_D_delta_e "
              "C_m_0 C_m_alpha C_m_q C_m_delta_e C_m_fp C_Y_0 C_Y_beta C_Y_p C_Y_r C_Y_delta_a C_Y_delta_r "
              "C_l_0 C_l_beta C_l_p C_l_r C_l_delta_a C_l_delta_r C_n_0 C_n_beta C_n_p C_n_r C_n_delta_a C_n_delta_r").split()


class CyclingDraws:
    """Stand-in for the env-level RandomState while sample_simulator_parameters runs: normal(loc, scale) walks a fixed
    list of standard-normal values (one of them far enough out to hit the clip), uniform() a fixed list of fractions."""
    Z = (0.7, -1.3, 2.5, -0.2, 1.1, -2.2, 0.05, 1.9)
    U = (0.625, 0.1, 0.9, 0.4)

    def __init__(self):
        self.kz = self.ku = 0

    def normal(self, loc=0.0, scale=1.0):
        z = self.Z[self.kz % len(self.Z)]
        self.kz += 1
        return loc + scale * z

    def uniform(self, low=0.0, high=1.0):
        u = self.U[self.ku % len(self.U)]
        self.ku += 1
        return low + u * (high - low)


def gen_model():
    """Per-episode aircraft-parameter randomisation (the gym config's simulator.model block,
    FixedWingAircraft.sample_simulator_parameters fixed_wing.py:748-813): the parameters the reference draws at every
    reset (env RNG replaced by fixed draws) and the trajectories flown with them; gaussian with clip, then uniform."""
    import tempfile
    cfg = json.load(open(refshim.GYM_CONFIG))
    for dist in ("gaussian", "uniform"):
        cfg["simulator"]["model"] = dict(MODEL_BLOCK, distribution=dist)
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(cfg, f)
        env = make_env(True, config_path=f.name)
        env.np_random = CyclingDraws()
        rs = np.random.RandomState(515 + len(dist))
        out = run_episodes(env, 3, 150, rs, True, wind_mag=4.0, action_amp=1.1,
                           after_reset=lambda e: np.array([e.simulator.params[k] for k in AERO_NAMES], dtype=np.float64))
        os.unlink(f.name)
        np.savez_compressed(os.path.join(HERE, "traj_model_%s.npz" % dist), **out)
        print(dist, "C_L_alpha", out["params"][:, AERO_NAMES.index("C_L_alpha")], "C_n_r", out["params"][:, AERO_NAMES.index("C_n_r")])


def gen_fail():
    """Episodes started near/over the constraint envelope so that ConstraintException paths fire."""
    rs = np.random.RandomState(31337)
    env = make_env(False)
    out = run_episodes(env, 24, 150, rs, False, wind_mag=5.0, action_amp=3.0, hard=True, metrics=True)
    print("fail terms:", out["term"].max(axis=1), "n_valid", out["n_valid"])
    np.savez_compressed(os.path.join(HERE, "traj_fail.npz"), **out)


def gen_full():
    """Whole episodes to the step limit (done by 'steps') with the 9 end-of-episode metrics."""
    rs = np.random.RandomState(2024)
    env = make_env(False, config_kw={"steps_max": 400})
    out = run_episodes(env, 4, 400, rs, False, wind_mag=3.0, action_amp=1.0, metrics=True)
    np.savez_compressed(os.path.join(HERE, "traj_full400.npz"), **out)
    env = make_env(True, config_kw={"steps_max": 300})
    rs = np.random.RandomState(2025)
    out = run_episodes(env, 3, 300, rs, True, wind_mag=3.0, action_amp=0.6, metrics=True, noise_len=300)
    np.savez_compressed(os.path.join(HERE, "traj_full300_turb.npz"), **out)


class FixedDraws:
    """Stand-in for the env-level RandomState: uniform(lo, hi) always returns lo + 0.625 (hi - lo) (i.e. 0.25 for the
    U(-1, 1) draw behind `init_noise`, fixed_wing.py:1145), normal(loc, scale) returns loc."""

    def uniform(self, low=0.0, high=1.0):
        return low + 0.625 * (high - low)

    def normal(self, loc=0.0, scale=1.0):
        return loc


def gen_cnn():
    """The reference's CNN-controller config (examples/models/cnn_controller/fixed_wing_config.json): observation
    length 5, shape matrix, relative targets, no alpha/beta — the general observation path."""
    path = os.path.join(os.path.dirname(refshim.GYM_CONFIG), "examples", "models", "cnn_controller",
                        "fixed_wing_config.json")
    env = make_env(False, config_path=path)
    env.np_random = FixedDraws()
    rs = np.random.RandomState(77)
    flat_env = env

    class Flat:                      # run_episodes stores obs as arrays: flatten the [5, 12] matrix
        def __getattr__(self, k):
            return getattr(flat_env, k)

        def reset(self, **kw):
            return np.asarray(flat_env.reset(**kw)).ravel()

        def step(self, a):
            o, r, d, i = flat_env.step(a)
            return np.asarray(o).ravel(), r, d, i

    out = run_episodes(Flat(), 4, 60, rs, False, wind_mag=4.0, action_amp=1.3)
    np.savez_compressed(os.path.join(HERE, "traj_cnn_obs.npz"), **out)


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


class FixedDrawsU(FixedDraws):
    def __init__(self, u):
        self.u = u

    def uniform(self, low=0.0, high=1.0):
        return low + self.u * (high - low)


MOVING_TARGETS = [
    {"name": "roll", "convert_to_radians": True, "low": -60, "high": 60, "delta": 180, "class": "linear",
     "slope_low": 2, "slope_high": 8, "bound": 5},
    {"name": "pitch", "convert_to_radians": True, "low": -25, "high": 25, "delta": 45, "class": "sinusoidal",
     "amplitude_low": 3, "amplitude_high": 9, "period_low": 60, "period_high": 140, "bound": 5},
    {"name": "Va", "low": 15, "high": 28, "delta": 6, "class": "compensate", "bound": 2}]


def integrator_env_config(cfg, W, L, step):
    """Default config + error integrals: `integration_window` W, observation rows with lags 1, 1 + step, ... carrying the
    default entries (without alpha / beta) and "integrator" targets, and "int_error" reward factors (general engine)."""
    cfg["integration_window"] = W
    ob = cfg["observation"]
    ob["length"], ob["step"], ob["shape"] = L, step, "vector"
    ob["states"] = [s_ for s_ in ob["states"] if s_["name"] not in ("alpha", "beta")]      # 12 + 3 entries (max 16)
    for name in TARGET_KEYS:
        ob["states"].append({"name": name, "type": "target", "value": "integrator"})
    for name, sc in zip(TARGET_KEYS, (40.0, 25.0, 300.0)):
        cfg["reward"]["factors"].append({"name": name, "class": "state", "type": "int_error", "function_class": "linear",
                                         "scaling": sc, "shaping": False, "max": 2.0, "sign": -1})
    cfg["steps_max"] = 60
    return cfg


def gen_integrator():
    """`integrator` observations and `int_error` rewards (fixed_wing.py:1003-1012, 1165-1180) with a window of 4 and rows
    at lags 1 / 3 / 5 (observation.step 2), and with the window of 0 every config of the reference tree carries (the sum
    then runs over the WHOLE history).  ONE env object per fixture across all episodes: the reset observation of an
    integrator entry reads the error history of the episode that just ended."""
    import tempfile
    for tag, W, L, step in (("w4", 4, 3, 2), ("w0", 0, 1, 1)):
        cfg = integrator_env_config(json.load(open(refshim.GYM_CONFIG)), W, L, step)
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(cfg, f)
        env = make_env(False, config_path=f.name)
        env.np_random = FixedDraws()
        rs = np.random.RandomState(515 + W)
        out = run_episodes(env, 5, 45, rs, False, wind_mag=3.0, action_amp=1.4)
        np.savez_compressed(os.path.join(HERE, "traj_integrator_%s.npz" % tag), **out)
        print(tag, "obs dim", out["obs"].shape[-1], "n_valid", out["n_valid"], "reward range", out["reward"].min(), out["reward"].max())


ANG_KEYS = ["omega_p", "omega_q", "omega_r"]


def angular_env_config(cfg):
    """Default config + target class attitude_angular (fixed_wing.py:671-675, 1455-1460, 1558-1642): omega_p/q/r as
    derived target states with bounds, observed (absolute and relative), rewarded (error factors, goal per_state) and
    part of the success streak; short episodes with a wide streak so that goals fire."""
    for name, bound in zip(ANG_KEYS, (0.6, 0.4, 0.5)):
        cfg["target"]["states"].append({"name": name, "class": "attitude_angular", "bound": bound})
    cfg["target"]["states"][-1]["max_vel"] = 2.5
    ob = cfg["observation"]
    ob["states"] = [s_ for s_ in ob["states"] if s_["name"] not in ("alpha", "beta")]
    for name in ANG_KEYS:
        ob["states"].append({"name": name, "type": "target", "value": "absolute"})
    ob["states"].append({"name": "omega_q", "type": "target", "value": "relative"})
    for name, sc in zip(ANG_KEYS, (6.0, 5.0, 4.0)):
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


def gen_angular():
    """Target class attitude_angular against the live reference: six target states.  The env RNG is replaced by fixed draws
    (the rate targets sampled at reset derive from the SAMPLED roll / pitch targets, before the injected ones override
    them, fixed_wing.py:441-450): u = 0.625, fed to our side as rng_u_override."""
    import tempfile
    cfg = angular_env_config(json.load(open(refshim.GYM_CONFIG)))
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(cfg, f)
    env = make_env(False, config_path=f.name)
    env.np_random = FixedDraws()
    rs = np.random.RandomState(909)
    n_ep, n_steps = 5, 50
    keys6 = TARGET_KEYS + ANG_KEYS
    rec = {k: [] for k in ("init_state", "init_target", "obs0", "target0", "actions", "obs", "reward", "done", "term",
                           "target", "n_valid")}
    mrec = []
    for ep in range(n_ep):
        st, tgt = random_scenario(rs, 3.0, False)
        obs0 = env.reset(state=dict(st), target=dict(tgt))
        s_arr, t_arr = scenario_arrays(st, tgt)
        rec["init_state"].append(s_arr), rec["init_target"].append(t_arr)
        rec["obs0"].append(np.array(obs0, dtype=np.float64))
        rec["target0"].append(np.array([env.target[k] for k in keys6], dtype=np.float64))
        er = {k: [] for k in ("actions", "obs", "reward", "done", "term", "target")}
        n_valid = n_steps
        for t in range(n_steps):
            a = rs.uniform(-1.3, 1.3, 3)
            obs, rew, done, info = env.step(a)
            er["actions"].append(a), er["obs"].append(np.array(obs, dtype=np.float64)), er["reward"].append(float(rew))
            er["done"].append(bool(done)), er["term"].append(term_code(info.get("termination", "")))
            er["target"].append(np.array([env.target[k] for k in keys6], dtype=np.float64))
            if done:
                n_valid = t + 1
                row = []
                for name in ("avg_error", "total_error", "end_error", "rise_time", "overshoot"):
                    row += [float(info[name][k]) for k in keys6]
                for name in ("success", "settling_time", "success_time_frac"):
                    row += [float(info[name][k]) for k in keys6 + ["all"]]
                row.append(float(info["control_variation"]["all"]))
                mrec.append(row)
                break
        rec["n_valid"].append(n_valid)
        for k, v in er.items():
            arr = np.array(v)
            rec[k].append(np.pad(arr, [(0, n_steps - arr.shape[0])] + [(0, 0)] * (arr.ndim - 1)))
    out = {k: np.array(v) for k, v in rec.items()}
    out["metrics52"] = np.array(mrec)        # 5 x 6 error metrics | 3 x 7 goal metrics | control_variation
    np.savez_compressed(os.path.join(HERE, "traj_angular.npz"), **out)
    print("n_valid", out["n_valid"], "terms", [int(out["term"][e, out["n_valid"][e] - 1]) for e in range(n_ep)],
          "reward range", out["reward"].min(), out["reward"].max(), "metric rows", out["metrics52"].shape)


def gen_targets():
    """Target classes linear / sinusoidal (+ Va compensate on a sinusoidal pitch target), fixed_wing.py:698-727,
    1438-1452.  Targets are NOT injected (an injected target forces class constant, :446-450): the env-level RNG is
    replaced by fixed draws u, and the same u is fed to our side through FwConfig.rng_u_override."""
    import tempfile
    cfg = json.load(open(refshim.GYM_CONFIG))
    cfg["target"]["states"] = MOVING_TARGETS
    cfg["steps_max"] = 150
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(cfg, f)
    out = {}
    for tag, u in (("a", 0.625), ("b", 0.3)):
        env = make_env(False, config_path=f.name)
        env.np_random = FixedDrawsU(u)
        rs = np.random.RandomState(int(u * 1000))
        rec = {k: [] for k in ("init_state", "actions", "target", "obs", "reward", "obs0", "target0")}
        for ep in range(3):
            st, _ = random_scenario(rs, 3.0)
            obs0 = env.reset(state=dict(st))
            rec["init_state"].append(scenario_arrays(st, {"roll": 0, "pitch": 0, "Va": 0})[0])
            rec["obs0"].append(np.array(obs0))
            rec["target0"].append([env.target[k] for k in TARGET_KEYS])
            A, Tg, Ob, Rw = [], [], [], []
            for t in range(150):
                a = rs.uniform(-1.0, 1.0, 3)
                obs, rew, done, info = env.step(a)
                A.append(a), Tg.append([env.target[k] for k in TARGET_KEYS]), Ob.append(np.array(obs)), Rw.append(rew)
            for k, v in (("actions", A), ("target", Tg), ("obs", Ob), ("reward", Rw)):
                rec[k].append(np.array(v))
        out.update({tag + "_" + k: np.array(v) for k, v in rec.items()})
        out[tag + "_u"] = u
    np.savez_compressed(os.path.join(HERE, "traj_moving_targets.npz"), **out)


def gen_resample():
    """resample_every and on_success = "new" (fixed_wing.py:536-580, 654-746) with the default target classes: the
    env-level RNG is replaced by fixed draws u (as in gen_targets), targets are sampled, not injected."""
    import tempfile
    from conftest import resample_env_config
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(resample_env_config(), f)
    out = {}
    for tag, u in (("a", 0.41), ("b", 0.83)):
        env = make_env(False, config_path=f.name)
        env.np_random = FixedDrawsU(u)
        rs = np.random.RandomState(int(u * 1000))
        rec = {k: [] for k in ("init_state", "actions", "target", "obs", "reward", "obs0", "target0", "done")}
        for ep in range(3):
            st, _ = random_scenario(rs, 3.0)
            obs0 = env.reset(state=dict(st))
            rec["init_state"].append(scenario_arrays(st, {"roll": 0, "pitch": 0, "Va": 0})[0])
            rec["obs0"].append(np.array(obs0))
            rec["target0"].append([env.target[k] for k in TARGET_KEYS])
            A, Tg, Ob, Rw, Dn = [], [], [], [], []
            for t in range(150):
                a = rs.uniform(-0.6, 0.6, 3)
                obs, rew, done, info = env.step(a)
                A.append(a), Tg.append([env.target[k] for k in TARGET_KEYS]), Ob.append(np.array(obs)), Rw.append(rew)
                Dn.append(done)
                if done:
                    break
            n = len(A)
            pad = lambda v, w: np.concatenate([np.array(v, dtype=np.float64).reshape(n, -1), np.full((150 - n, w), np.nan)])
            rec["actions"].append(pad(A, 3)); rec["target"].append(pad(Tg, 3)); rec["obs"].append(pad(Ob, 14))
            rec["reward"].append(pad(Rw, 1)[:, 0]); rec["done"].append(np.array(Dn + [False] * (150 - n)))
        out.update({tag + "_" + k: np.array(v) for k, v in rec.items()})
        out[tag + "_u"] = u
    np.savez_compressed(os.path.join(HERE, "traj_resample.npz"), **out)


def gen_waypoint():
    """FixedWingAircraft_simple (magpie/magpy/simple_train.py:197-702), the waypoint env head.  The class is taken
    from the reference file where it lies; PyFly is constructed with turbulence off and the waypoints carry explicit
    omega so that nothing is drawn from an unseeded RandomState (the env re-creates its simulator on every reset)."""
    import ast
    import types
    import gym
    from pyfly import pyfly as pyfly_mod
    src_path = os.path.join(refshim.REFERENCE_ROOT, "magpie", "magpy", "simple_train.py")
    tree = ast.parse(open(src_path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "FixedWingAircraft_simple"][0]
    mod = types.ModuleType("simple_train_env")
    cfgdir = os.path.dirname(refshim.X8_PARAMS) + "/"

    def quiet_pyfly(config_path, parameter_path):
        return pyfly_mod.PyFly(config_path, parameter_path, config_kw={"turbulence": False})

    class NoRecorder:
        def __init__(self, *a, **k):
            pass

        def savestate(self, *a, **k):
            pass

    mod.__dict__.update(gym=gym, np=np, os=os, random=__import__("random"), PyFly=quiet_pyfly, configDir=cfgdir,
                        startingDir=os.path.dirname(src_path), simrecorder=NoRecorder, print=lambda *a, **k: None)
    exec(compile(ast.Module(body=[cls], type_ignores=[]), src_path, "exec"), mod.__dict__)
    rs = np.random.RandomState(31)
    n_tasks, wp_len = 3, 5
    keys = ["position_n", "position_e", "position_d", "roll", "pitch", "yaw", "velocity_u", "velocity_v", "velocity_w",
            "wind_n", "wind_e", "wind_d", "omega_p", "omega_q", "omega_r"]
    tasks = np.zeros((n_tasks, wp_len, 15))
    for t in range(n_tasks):
        p0 = np.array([rs.uniform(-20, 20), rs.uniform(-20, 20), rs.uniform(-90, -60)])
        for w in range(wp_len):
            tasks[t, w, :3] = p0 + w * np.array([1.6, 0.05 * (t - 1), -0.02 * t])    # ~1.6 m legs: reached within a few steps
            tasks[t, w, 3:6] = [rs.uniform(-0.1, 0.1), rs.uniform(-0.05, 0.05), rs.uniform(-0.02, 0.02)]
            tasks[t, w, 6:9] = [rs.uniform(16, 19), rs.uniform(-0.2, 0.2), rs.uniform(-0.2, 0.2)]
            tasks[t, w, 9:12] = [rs.uniform(-1, 1), rs.uniform(-1, 1), rs.uniform(-0.3, 0.3)]
            tasks[t, w, 12:15] = rs.uniform(-0.2, 0.2, 3)
    task_dicts = [np.array([{k: float(v) for k, v in zip(keys, row)} for row in tasks[t]], dtype=object)
                  for t in range(n_tasks)]

    class Env(mod.FixedWingAircraft_simple):
        def sample_tasks(self, num_tasks):
            return task_dicts

    out = {"tasks": tasks}
    T = 120
    for t in range(n_tasks):
        env = Env(None, n_tasks=n_tasks)
        env.set_skip(True)
        env.idx = t
        obs0 = env.reset()
        A, Ob, Rw, Dn, Pos = [], [], [], [], []
        for k in range(T):
            a = np.array([rs.uniform(-0.1, 0.1), rs.uniform(-0.1, 0.1), rs.uniform(0.3, 0.9)])
            obs, rew, done, info = env.step(a)
            A.append(a), Ob.append(np.array(obs)), Rw.append(float(rew)), Dn.append(bool(done)), Pos.append(env.cur_pos)
        out.update({"obs0_%d" % t: np.array(obs0), "actions_%d" % t: np.array(A), "obs_%d" % t: np.array(Ob),
                    "reward_%d" % t: np.array(Rw), "done_%d" % t: np.array(Dn), "pos_%d" % t: np.array(Pos)})
        print("task", t, "waypoint index over time:", sorted(set(Pos)), "final reward", Rw[-1])
    np.savez_compressed(os.path.join(HERE, "traj_waypoint.npz"), **out)


def gen_reward():
    """A config that exercises the whole reward engine (fixed_wing.py:941-1111): potential form, three terms, every
    factor class; wide goal bounds and a short streak so that goals and the one-off success bonus fire."""
    import tempfile
    cfg = json.load(open(refshim.GYM_CONFIG))
    cfg["reward"] = RICH_REWARD
    cfg["steps_max"] = 80
    cfg["target"]["success_streak_req"] = 6
    cfg["target"]["success_streak_fraction"] = 0.5
    for s_, b in zip(cfg["target"]["states"], (60, 40, 12)):
        s_["bound"] = b
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(cfg, f)
    env = make_env(False, config_path=f.name)
    rs = np.random.RandomState(4242)
    out = run_episodes(env, 5, 80, rs, False, wind_mag=3.0, action_amp=1.8)
    np.savez_compressed(os.path.join(HERE, "traj_reward_rich.npz"), **out)
    rs = np.random.RandomState(4243)
    out = run_episodes(env, 3, 80, rs, False, wind_mag=3.0, action_amp=1.8, f32_actions=True)
    np.savez_compressed(os.path.join(HERE, "traj_reward_rich_f32.npz"), **out)


def gen_pid(max_scen=100, num_envs=6):
    """Lock-step emulation of examples/evaluate_controller.py:57-232 (use_pid=True) with `num_envs` env slots.

    The reference's golden file was produced with 6 parallel envs (the metric completion order in the file is only reproduced by 6 slots; scenarios 0-5 start with `info is None`,
    every later scenario's FIRST action is computed with the PID reference still set to the final target of the
    episode that previously ran in the same slot, evaluate_controller.py:203-208).  We record that stale
    reference per scenario (`first_ref`) so each scenario can be replayed on its own.
    """
    scenarios = list(np.load(refshim.PID_TEST_SET, allow_pickle=True))[:max_scen]
    golden = np.load(refshim.PID_GOLDEN, allow_pickle=True).item()
    config_kw = {"steps_max": 1500,
                 "target": {"on_success": "done", "success_streak_fraction": 1, "success_streak_req": 100,
                            "states": {0: {"bound": 5}, 1: {"bound": 5}, 2: {"bound": 2}}},
                 "action": {"scale_space": False}}
    envs = [make_env(False, config_kw=config_kw) for _ in range(num_envs)]
    pids = [PIDController(envs[0].simulator.dt) for _ in range(num_envs)]
    obs_states = [v["name"] for v in envs[0].cfg["observation"]["states"]]
    i_phi, i_th, i_va = obs_states.index("roll"), obs_states.index("pitch"), obs_states.index("Va")
    i_om = [obs_states.index(k) for k in ("omega_p", "omega_q", "omega_r")]
    S = len(scenarios)
    skeys = STATE_KEYS + ["Va", "alpha", "beta", "elevator", "aileron", "throttle", "wind_n", "wind_e", "wind_d"]
    init_state = np.array([[float(sc["state"][k]) for k in skeys] for sc in scenarios])
    init_target = np.array([[float(sc["target"][k]) for k in TARGET_KEYS] for sc in scenarios])
    first_ref = np.full((S, 3), np.nan)
    live_rewards = np.zeros((S, 1500))
    live_len = np.zeros(S, dtype=np.int64)
    live_final_y = np.zeros((S, 19))
    live_metrics = [None] * S
    gold_len = np.array([len(golden["rewards"][i]) for i in range(S)], dtype=np.int64)
    gold_rewards = np.zeros((S, int(gold_len.max())))
    for i in range(S):
        gold_rewards[i, :gold_len[i]] = golden["rewards"][i]

    queue = list(range(S))
    slot_scen = [-1] * num_envs
    slot_t = [0] * num_envs
    active = [False] * num_envs
    done = [True] * num_envs
    obs = [None] * num_envs
    info = None
    t0 = time.time()
    while True:
        for i in range(num_envs):
            if done[i]:
                if queue:
                    si = queue.pop(0)
                    slot_scen[i], slot_t[i], active[i] = si, 0, True
                    sc = scenarios[si]
                    obs[i] = envs[i].reset(state=dict(sc["state"]), target=dict(sc["target"]))
                    pids[i].reset()
                    pids[i].set_reference(sc["target"]["roll"], sc["target"]["pitch"], sc["target"]["Va"])
                    if info is not None and info[i] is not None:
                        first_ref[si] = [info[i]["target"][k] for k in TARGET_KEYS]
                else:
                    active[i] = False
                done[i] = False
        if not queue and not any(active):
            break
        new_info = [None] * num_envs
        for i in range(num_envs):
            if not active[i]:
                continue
            if info is not None and info[i] is not None:
                pids[i].set_reference(phi=info[i]["target"]["roll"], theta=info[i]["target"]["pitch"],
                                      va=info[i]["target"]["Va"])
            a = pids[i].get_action(obs[i][i_phi], obs[i][i_th], obs[i][i_va], obs[i][i_om])
            o, rew, d, inf = envs[i].step(a)
            si = slot_scen[i]
            live_rewards[si, slot_t[i]] = rew
            slot_t[i] += 1
            obs[i] = o
            inf = dict(inf)
            inf["target"] = dict(inf["target"])   # SubprocVecEnv pickles info through the pipe: a snapshot
            new_info[i] = inf
            if d:
                done[i] = True
                live_len[si] = slot_t[i]
                live_final_y[si] = snapshot(envs[i])[0]
                live_metrics[si] = {k: inf[k] for k in ("rise_time", "settling_time", "overshoot", "total_error",
                                                        "avg_error", "control_variation", "success",
                                                        "success_time_frac", "end_error")}
                n = min(live_len[si], gold_len[si])
                print("scenario %3d len live %4d golden %4d  max|dr| %.2e  (%.0fs)" % (
                    si, live_len[si], gold_len[si],
                    np.abs(live_rewards[si, :n] - gold_rewards[si, :n]).max(), time.time() - t0), flush=True)
        # envs that are inactive keep their last info (the harness keeps stepping them; irrelevant here)
        info = [new_info[i] if new_info[i] is not None else (info[i] if info is not None else None)
                for i in range(num_envs)]
    out = dict(init_state=init_state, init_target=init_target, first_ref=first_ref, live_rewards=live_rewards,
               live_len=live_len, live_final_y=live_final_y, gold_rewards=gold_rewards, gold_len=gold_len)
    out.update(pack_metrics(live_metrics))
    for name in ("success", "rise_time", "overshoot", "settling_time", "control_variation"):
        keys = list(golden[name].keys())
        out["gold_" + name] = np.array([[np.nan if v is None else float(v) for v in golden[name][k]] for k in keys],
                                       dtype=np.float64).T
        out["gold_" + name + "_keys"] = np.array(keys)
    np.savez_compressed(os.path.join(HERE, "pid_none.npz"), **out)


def gen_gae():
    """Reference RolloutBuffer GAE on seeded data (buffers.py:304-333), incl. the mixed f32/f64 quirk."""
    import torch
    from stable_baselines3.common.buffers import RolloutBuffer
    out = {}
    for tag, (T, N, p_done) in {"a": (64, 16, 0.05), "b": (256, 8, 0.01), "c": (5, 3, 0.3)}.items():
        rs = np.random.RandomState(T * 1000 + N)
        obs_space = refshim.Box(-np.ones(14), np.ones(14), dtype=np.float32)
        act_space = refshim.Box(-np.ones(3), np.ones(3), dtype=np.float32)
        buf = RolloutBuffer(T, obs_space, act_space, device="cpu", gae_lambda=0.95, gamma=0.99, n_envs=N)
        rew = rs.standard_normal((T, N)).astype(np.float32)
        val = rs.standard_normal((T, N)).astype(np.float32)
        dones = (rs.uniform(size=(T, N)) < p_done)
        last_val = rs.standard_normal(N).astype(np.float32)
        last_done = (rs.uniform(size=N) < 0.2)
        for t in range(T):
            buf.add(np.zeros((N, 14), np.float32), np.zeros((N, 3), np.float32), rew[t], dones[t].astype(np.float32),
                    torch.as_tensor(val[t]), torch.zeros(N))
        buf.compute_returns_and_advantage(torch.as_tensor(last_val), dones=last_done)
        out.update({tag + "_rew": rew, tag + "_val": val, tag + "_done": dones.astype(np.float32),
                    tag + "_last_val": last_val, tag + "_last_done": last_done,
                    tag + "_adv": buf.advantages.copy(), tag + "_ret": buf.returns.copy()})
    np.savez_compressed(os.path.join(HERE, "gae.npz"), **out)


def gen_vecnorm():
    """Reference VecNormalize (+ RunningMeanStd) and RolloutBuffer.add / swap_and_flatten driven by a scripted VecEnv
    (vec_normalize.py:106-219, running_mean_std.py:19-39, buffers.py:51-64, 292-302): raw observations / rewards /
    dones in, normalised observations / rewards, running moments and buffer contents out."""
    import torch
    from stable_baselines3.common.buffers import RolloutBuffer
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv
    from stable_baselines3.common.vec_env.vec_normalize import VecNormalize
    T, N, D = 24, 6, 14
    rs = np.random.RandomState(2024)
    obs_seq = (rs.standard_normal((T + 1, N, D)) * rs.uniform(0.1, 30, D) + rs.uniform(-5, 5, D)).astype(np.float32)
    rew_seq = (rs.standard_normal((T, N)) * 3 - 1).astype(np.float32)
    done_seq = rs.uniform(size=(T, N)) < 0.15
    obs_space = refshim.Box(-np.inf * np.ones(D), np.inf * np.ones(D), dtype=np.float32)
    act_space = refshim.Box(-np.ones(3), np.ones(3), dtype=np.float32)

    class Scripted(VecEnv):
        def __init__(self):
            VecEnv.__init__(self, N, obs_space, act_space)
            self.t = 0

        def reset(self):
            self.t = 0
            return obs_seq[0].copy()

        def step_async(self, actions):
            pass

        def step_wait(self):
            t = self.t
            self.t += 1
            return obs_seq[t + 1].copy(), rew_seq[t].copy(), done_seq[t].copy(), [{} for _ in range(N)]

        def close(self): pass
        def get_attr(self, *a, **k): return [None] * N
        def set_attr(self, *a, **k): pass
        def env_method(self, *a, **k): return [None] * N
        def seed(self, seed=None): return [None] * N

    venv = VecNormalize(Scripted(), training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0,
                        gamma=0.99, epsilon=1e-8)
    buf = RolloutBuffer(T, obs_space, act_space, device="cpu", gae_lambda=0.95, gamma=0.99, n_envs=N)
    acts = rs.uniform(-1, 1, (T, N, 3)).astype(np.float32)
    vals = rs.standard_normal((T, N)).astype(np.float32)
    logp = rs.standard_normal((T, N)).astype(np.float32)
    last_obs = venv.reset()
    last_dones = np.zeros(N, dtype=bool)
    nobs, nrew = [last_obs.copy()], []
    for t in range(T):
        venv.step_async(acts[t])
        o, r, d, _ = venv.step_wait()
        buf.add(last_obs, acts[t], r, last_dones, torch.as_tensor(vals[t]), torch.as_tensor(logp[t]))   # on_policy_algorithm.py:178
        last_obs, last_dones = o, d
        nobs.append(o.copy()); nrew.append(np.asarray(r, dtype=np.float64).copy())
    out = dict(obs_seq=obs_seq, rew_seq=rew_seq, done_seq=done_seq, acts=acts, vals=vals, logp=logp,
               norm_obs=np.stack(nobs), norm_rew=np.stack(nrew),
               obs_mean=venv.obs_rms.mean, obs_var=venv.obs_rms.var, obs_count=np.float64(venv.obs_rms.count),
               ret_mean=np.float64(venv.ret_rms.mean), ret_var=np.float64(venv.ret_rms.var),
               ret_count=np.float64(venv.ret_rms.count), ret=venv.ret.copy(),
               buf_obs=buf.observations.copy(), buf_act=buf.actions.copy(), buf_rew=buf.rewards.copy(),
               buf_done=buf.dones.copy(), buf_val=buf.values.copy(), buf_logp=buf.log_probs.copy(),
               flat_obs=RolloutBuffer.swap_and_flatten(buf.observations.copy()))
    np.savez_compressed(os.path.join(HERE, "vecnorm.npz"), **out)


def gen_ppo_update():
    """Three optimiser steps of the reference PPO.train() (ppo/ppo.py:133-240: evaluate_actions, advantage
    normalisation, clipped surrogate, value MSE, entropy bonus, clip_grad_norm_ 0.5, Adam lr 3e-4 eps 1e-5) on a
    hand-filled rollout buffer, one full-batch minibatch per call: weights before and after every step."""
    import torch
    from stable_baselines3 import PPO
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv
    T, N, D = 32, 16, 14
    obs_space = refshim.Box(-np.inf * np.ones(D), np.inf * np.ones(D), dtype=np.float32)
    act_space = refshim.Box(-np.ones(3), np.ones(3), dtype=np.float32)

    class Scripted(VecEnv):
        def __init__(self):
            VecEnv.__init__(self, N, obs_space, act_space)

        def reset(self): return np.zeros((N, D), np.float32)
        def step_async(self, actions): pass
        def step_wait(self): return np.zeros((N, D), np.float32), np.zeros(N, np.float32), np.zeros(N, bool), [{} for _ in range(N)]
        def close(self): pass
        def get_attr(self, *a, **k): return [None] * N
        def set_attr(self, *a, **k): pass
        def env_method(self, *a, **k): return [None] * N
        def seed(self, seed=None): return [None] * N

    torch.manual_seed(0)
    m = PPO("MlpPolicy", Scripted(), n_steps=T, batch_size=T * N, n_epochs=1, ent_coef=0.01, vf_coef=0.5,
            max_grad_norm=0.5, learning_rate=3e-4, clip_range=0.2, device="cpu", verbose=0)
    with torch.no_grad():                                  # spread the ratios on both sides of the clip range
        m.policy.action_net.weight.mul_(40.0)
        m.policy.log_std.copy_(torch.tensor([-0.3, 0.1, -0.8]))
    rb = m.rollout_buffer
    rs = np.random.RandomState(7)
    rb.observations[:] = rs.standard_normal(rb.observations.shape)
    rb.actions[:] = rs.uniform(-1, 1, rb.actions.shape)
    rb.values[:] = rs.standard_normal(rb.values.shape)
    rb.log_probs[:] = rs.standard_normal(rb.log_probs.shape) * 0.3 - 3.0
    rb.advantages[:] = rs.standard_normal(rb.advantages.shape) * 2 + 0.5
    rb.returns[:] = rs.standard_normal(rb.returns.shape)
    rb.full, rb.pos = True, T
    m._current_progress_remaining = 1.0
    names = {"log_std": "log_std", "mlp_extractor.policy_net.0": "pi.0", "mlp_extractor.policy_net.2": "pi.2",
             "mlp_extractor.value_net.0": "vf.0", "mlp_extractor.value_net.2": "vf.2", "action_net": "action_net",
             "value_net": "value_net"}

    def weights(tag):
        o = {}
        for k, v in m.policy.state_dict().items():
            base, _, leaf = k.rpartition(".")
            mine = names[k] if k in names else names[base] + "." + leaf
            o["%s/%s" % (tag, mine)] = v.detach().numpy().copy()
        return o

    out = dict(obs=rb.observations.reshape(T * N, D).copy(), act=rb.actions.reshape(T * N, 3).copy(),
               old_values=rb.values.reshape(-1).copy(), old_log_prob=rb.log_probs.reshape(-1).copy(),
               adv=rb.advantages.reshape(-1).copy(), ret=rb.returns.reshape(-1).copy())
    out.update(weights("w0"))
    with torch.no_grad():
        v, lp, ent = m.policy.evaluate_actions(torch.as_tensor(out["obs"]), torch.as_tensor(out["act"]))
    out.update(eval_values=v.numpy().reshape(-1), eval_log_prob=lp.numpy(), eval_entropy=ent.numpy())
    for k in (1, 2, 3):
        m.train()
        out.update(weights("w%d" % k))
    np.savez_compressed(os.path.join(HERE, "ppo_update.npz"), **out)


def gen_sac_update():
    """Three gradient steps of the reference SAC.train() (sac/sac.py:177-269: squashed-Gaussian actor, twin critics,
    automatic entropy coefficient, polyak 0.005, Adam lr 3e-4) on a hand-filled single-env replay buffer: the sampled row
    indices, the unit normals behind both policy samples of every step, and all weights before / after each step.  A
    second part records ReplayBuffer.sample with a VecNormalize env: normalisation at SAMPLE time (buffers.py:245-254)."""
    import torch
    from stable_baselines3 import SAC
    from stable_baselines3.common.vec_env import VecNormalize
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv
    D, B, CAP = 14, 64, 200
    obs_space = refshim.Box(-np.inf * np.ones(D), np.inf * np.ones(D), dtype=np.float32)
    act_space = refshim.Box(-np.ones(3), np.ones(3), dtype=np.float32)

    class Scripted(VecEnv):
        def __init__(self):
            VecEnv.__init__(self, 1, obs_space, act_space)

        def reset(self): return np.zeros((1, D), np.float32)
        def step_async(self, actions): pass
        def step_wait(self): return np.zeros((1, D), np.float32), np.zeros(1, np.float32), np.zeros(1, bool), [{}]
        def close(self): pass
        def get_attr(self, *a, **k): return [None]
        def set_attr(self, *a, **k): pass
        def env_method(self, *a, **k): return [None]
        def seed(self, seed=None): return [None]

    torch.manual_seed(0)
    m = SAC("MlpPolicy", Scripted(), buffer_size=CAP, batch_size=B, learning_rate=3e-4, gamma=0.99, tau=0.005,
            policy_kwargs=dict(net_arch=[64, 64]), device="cpu", verbose=0)     # 2 x 64 keeps the fixture small
    rs = np.random.RandomState(11)
    rb = m.replay_buffer
    for k in range(150):
        rb.add(rs.standard_normal((1, D)).astype(np.float32), rs.standard_normal((1, D)).astype(np.float32),
               rs.uniform(-1, 1, (1, 3)).astype(np.float32), rs.standard_normal(1).astype(np.float32),
               (rs.uniform(size=1) < 0.1).astype(np.float32))
    out = dict(rb_obs=rb.observations[:150, 0].copy(), rb_next_obs=rb.next_observations[:150, 0].copy(),
               rb_act=rb.actions[:150, 0].copy(), rb_rew=rb.rewards[:150, 0].copy(), rb_done=rb.dones[:150, 0].copy())

    def weights(tag):
        o = {}
        for k, v in m.policy.state_dict().items():
            o["%s/%s" % (tag, k)] = v.detach().numpy().copy()
        o["%s/log_ent_coef" % tag] = m.log_ent_coef.detach().numpy().copy()
        return o

    picked = []
    inner = rb._get_samples
    rb._get_samples = lambda batch_inds, env=None: (picked.append(np.array(batch_inds)), inner(batch_inds, env=env))[1]
    rec = []
    alp = m.actor.action_log_prob
    m.actor.action_log_prob = lambda obs: (lambda r: (rec.append((r[0].detach().numpy().copy(), r[1].detach().numpy().copy())), r)[1])(alp(obs))
    m._current_progress_remaining = 1.0
    out.update(weights("w0"))
    for k in (1, 2, 3):
        np.random.seed(50 + k)
        torch.manual_seed(100 + k)
        m.train(gradient_steps=1, batch_size=B)
        # the two policy samples of the step draw their unit normals from the default generator, in this order
        torch.manual_seed(100 + k)
        out["eps_pi_%d" % k], out["eps_next_%d" % k] = torch.randn(B, 3).numpy(), torch.randn(B, 3).numpy()
        out["idx_%d" % k] = picked[-1]
        out["actions_pi_%d" % k], out["log_prob_%d" % k] = rec[-2]
        out.update(weights("w%d" % k))
    # sample-time normalisation
    venv = VecNormalize(Scripted(), clip_obs=5.0, clip_reward=3.0)
    venv.obs_rms.update(rs.standard_normal((300, D)) * 2 + 0.5)
    venv.ret_rms.update(rs.standard_normal(300) * 4)
    idx = rs.randint(0, 150, size=40)
    smp = inner(idx, env=venv)
    out.update(norm_idx=idx, norm_obs=smp.observations.numpy(), norm_next_obs=smp.next_observations.numpy(),
               norm_rew=smp.rewards.numpy().reshape(-1), norm_done=smp.dones.numpy().reshape(-1), norm_act=smp.actions.numpy(),
               norm_obs_mean=venv.obs_rms.mean, norm_obs_var=venv.obs_rms.var, norm_ret_var=np.float64(venv.ret_rms.var))
    np.savez_compressed(os.path.join(HERE, "sac_update.npz"), **out)


def gen_curriculum():
    """set_curriculum_level (fixed_wing.py:334-412) on the default config and on the dev config: initial-state
    ranges of the simulator and target ranges, for several levels."""
    out = {}
    dev = os.path.join(os.path.dirname(refshim.GYM_CONFIG), "fixed_wing_config_dev.json")
    for tag, path in (("default", refshim.GYM_CONFIG), ("dev", dev)):
        env = FixedWingAircraft(path, config_kw={}, sim_config_kw={"turbulence": False})
        for level in (0.0, 0.25, 0.6, 1.0):
            env.set_curriculum_level(level)
            s = env.simulator.state
            lo = [float(s[k].init_min) if s[k].init_min is not None else np.nan for k in STATE_KEYS]
            hi = [float(s[k].init_max) if s[k].init_max is not None else np.nan for k in STATE_KEYS]
            tp = env._target_props_init["states"]
            conv = lambda st, v: np.radians(v) if tp[st].get("convert_to_radians", False) else v
            key = "%s_%g" % (tag, level)
            out[key + "_init_lo"], out[key + "_init_hi"] = np.array(lo), np.array(hi)
            for f in ("low", "high", "delta"):
                out[key + "_tgt_" + f] = np.array([conv(st, tp[st][f]) for st in TARGET_KEYS], dtype=np.float64)
        out[tag + "_obs_low"], out[tag + "_obs_high"] = env.observation_space.low, env.observation_space.high
        out[tag + "_act_low"], out[tag + "_act_high"] = env.action_space.low, env.action_space.high
    np.savez_compressed(os.path.join(HERE, "curriculum.npz"), **out)


def gen_dryden():
    """Reference Dryden output for injected noise, for the gym parameterisation (dt<-2000, b<-0.01, h<-2.1;
    pyfly.py:781-783 vs dryden.py:52) and for raw pyfly (sim_length 300), all three intensities."""
    from pyfly.pyfly import Wind
    out = {}
    for L in (2000, 300):
        for inten in ("light", "moderate", "severe"):
            w = Wind(turbulence=True, mag_min=-8, mag_max=8, b=2.1, turbulence_intensity=inten, sim_length=L, dt=0.01)
            rs = np.random.RandomState(L + len(inten))
            noise = rs.standard_normal((4, L))
            w.reset([0.0, 0.0, 0.0], noise)
            w.get_turbulence_linear(0)
            tag = "L%d_%s" % (L, inten)
            out[tag + "_noise"] = noise
            out[tag + "_lin"] = np.array(w.dryden.vel_lin)
            out[tag + "_ang"] = np.array(w.dryden.vel_ang)
    # two blocks of turbulence_sim_length = 300: the second simulate() call restarts lsim from the last state
    w = Wind(turbulence=True, mag_min=-8, mag_max=8, b=2.1, turbulence_intensity="moderate", sim_length=300, dt=0.01)
    noise = np.random.RandomState(600).standard_normal((4, 600))
    w.reset([0.0, 0.0, 0.0], noise)
    w.get_turbulence_linear(0)
    w.get_turbulence_linear(300)
    out["blocks_noise"], out["blocks_lin"], out["blocks_ang"] = noise, np.array(w.dryden.vel_lin), np.array(w.dryden.vel_ang)
    np.savez_compressed(os.path.join(HERE, "dryden.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    jobs = {"angular": gen_angular, "integrator": gen_integrator, "waypoint": gen_waypoint, "targets": gen_targets, "resample": gen_resample, "reward": gen_reward, "cnn": gen_cnn, "params": gen_params, "traj": gen_traj, "turb": gen_turb, "turb_moderate": gen_turb_moderate, "model": gen_model, "fail": gen_fail, "full": gen_full,
            "pid": gen_pid, "gae": gen_gae, "vecnorm": gen_vecnorm, "ppo_update": gen_ppo_update, "sac_update": gen_sac_update, "curriculum": gen_curriculum, "dryden": gen_dryden}
    for name, fn in jobs.items():
        if what in (name, "all"):
            t0 = time.time()
            fn()
            print("[%s] done in %.1fs" % (name, time.time() - t0), flush=True)
