"""Pins the oracle (oracle/fw_oracle.c): the C restatement must reproduce
  (1) fixtures recorded by running the unmodified reference (tests/golden/make_golden.py), and
  (2) the reference's own golden PID evaluation (examples/evaluations/eval_res_PID_none.npy on
      examples/test_sets/test_set_wind_none_step20-20-3.npy), stored in tests/golden/pid_none.npz.
CPU only.  These are the tests that entitle the GPU parity tests to use the oracle as their checker."""
import numpy as np
import pytest

from conftest import TRAJ_CASES, close_or_both_nan, golden_metric_rows, load_golden
from oracle import fw_oracle as O
from oracle import pid as P
from tum_adlr_deep_reinforcement_learning_b200.config import build_config


def _rel(a, b):
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


@pytest.mark.parametrize("name,cfg_kw,sim_kw,f32", TRAJ_CASES, ids=[c[0] for c in TRAJ_CASES])
def test_oracle_reproduces_live_reference_trajectories(name, cfg_kw, sim_kw, f32):
    g = load_golden(name)
    cfg = build_config(config_kw=cfg_kw, sim_config_kw=sim_kw)
    E = g["actions"].shape[0]
    rows, eps = golden_metric_rows(g) if "m_success" in g.files else ([], [])
    for ep in range(E):
        env = O.OracleEnv(cfg)
        noise = g["noise"][ep] if "noise" in g.files else None
        obs = env.reset(g["init_state"][ep], g["init_target"][ep], noise)
        assert np.abs(obs - g["obs0"][ep]).max() < 1e-12
        s = env.get()
        assert np.abs(s["y"] - g["y0"][ep]).max() < 1e-14
        if noise is not None:
            ref = np.concatenate([g["turb_lin"][ep], g["turb_ang"][ep]])
            assert np.abs(env.turbulence() - ref).max() < 1e-15
        for t in range(int(g["n_valid"][ep])):
            obs, rew, done, term = env.step(g["actions"][ep, t], f32)
            s = env.get()
            assert done == bool(g["done"][ep, t]) and term == int(g["term"][ep, t]), (ep, t)
            assert s["nfev"] == int(g["nfev"][ep, t]), (ep, t, "RK45 step-size decisions differ")
            if term < 10:
                assert _rel(s["y"], g["y"][ep, t]).max() < 1e-9, (ep, t)
                assert _rel(s["vab"], g["vab"][ep, t]).max() < 1e-9
            assert _rel(obs, g["obs"][ep, t]).max() < 1e-9, (ep, t)
            assert abs(rew - g["reward"][ep, t]) < 1e-9 * max(1.0, abs(g["reward"][ep, t]))
            assert _rel(s["target"], g["target"][ep, t]).max() < 1e-12
            assert _rel(s["cmd"], g["cmd"][ep, t]).max() < 1e-12
        if ep in eps:
            m, ret, ln, term = env.metrics()
            row = rows[eps.index(ep)]
            assert close_or_both_nan(m, row, 1e-7, 1e-9).all(), (ep, m, row)
            assert ln == int(g["n_valid"][ep])


@pytest.mark.parametrize("dist", ["gaussian", "uniform"])
def test_oracle_aircraft_parameter_randomisation(dist):
    """simulator.model (fixed_wing.py:748-813): with the parameters the live reference drew at every reset injected, the
    oracle flies the reference's trajectories (identical RK45 decisions); its own Philox draws follow the configured
    distribution: unlisted parameters and parameters whose original is 0 stay put, the clip holds (a negative original
    turns the relative clip inside out and pins the value at orig + clip, as np.clip does), Jx changes nothing."""
    from conftest import model_env_config
    g = load_golden("traj_model_" + dist)
    cfg = build_config(env_cfg=model_env_config(dist), sim_config_kw={"turbulence": True})
    assert cfg.model_on == 1 and cfg.model_uniform == int(dist == "uniform")
    for ep in range(g["actions"].shape[0]):
        env = O.OracleEnv(cfg)
        obs = env.reset(g["init_state"][ep], g["init_target"][ep], g["noise"][ep])
        assert np.abs(obs - g["obs0"][ep]).max() < 1e-12
        own = env.params()
        env.set_params(g["params"][ep])
        for t in range(int(g["n_valid"][ep])):
            obs, rew, done, term = env.step(g["actions"][ep, t])
            s = env.get()
            assert done == bool(g["done"][ep, t]) and s["nfev"] == int(g["nfev"][ep, t]), (ep, t)
            assert _rel(s["y"], g["y"][ep, t]).max() < 1e-9 and _rel(obs, g["obs"][ep, t]).max() < 1e-9, (ep, t)
        nominal = np.array([cfg.par_orig[i] for i in range(48)])
        on = np.array([bool(cfg.par_enabled[i]) for i in range(48)])
        assert np.array_equal(own[~on], nominal[~on])
        assert own[21] == nominal[21] == 0.0                       # C_D_q: listed, original 0 -> never drawn
        assert abs(own[45] - (nominal[45] + 0.05 * nominal[45])) < 1e-15 or dist == "uniform"    # C_n_r pinned by the clip
    # distribution of the oracle's own draws
    batch = O.OracleBatch(cfg, 4000)
    batch.reset()
    P_ = batch.params()
    i = 17                                                             # C_L_alpha: var 0.1 |orig|, clip 0.15 orig
    o = cfg.par_orig[i]
    if dist == "gaussian":
        assert abs(P_[:, i].mean() - o) < 0.01 * o and P_[:, i].min() >= o * 0.85 - 1e-12 and P_[:, i].max() <= o * 1.15 + 1e-12
        assert 0.08 * o < P_[:, i].std() < 0.1 * o                     # N(o, 0.1 o) clipped at 1.5 sigma
    else:
        assert abs(P_[:, i].mean() - o) < 0.01 * o and abs(P_[:, i].std() - 0.2 * o / np.sqrt(12)) < 0.005 * o
        assert P_[:, i].min() >= 0.9 * o and P_[:, i].max() <= 1.1 * o


def test_dryden_matches_reference_for_all_intensities():
    g = load_golden("dryden")
    for L in (2000, 300):
        for inten in ("light", "moderate", "severe"):
            tag = "L%d_%s" % (L, inten)
            cfg = build_config(config_kw={"steps_max": L},
                               sim_config_kw={"turbulence": True, "turbulence_intensity": inten})
            out = O.dryden(cfg, g[tag + "_noise"])
            ref = np.concatenate([g[tag + "_lin"], g[tag + "_ang"]])
            assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max(), tag
            assert np.all(out[:, 0] == 0)      # column 0 is exactly zero (x_0 = 0, D = 0)


def test_gae_bit_exact():
    g = load_golden("gae")
    for tag in ("a", "b", "c"):
        adv, ret = O.gae(g[tag + "_rew"], g[tag + "_val"], g[tag + "_done"], g[tag + "_last_val"],
                         g[tag + "_last_done"])
        assert np.array_equal(adv, g[tag + "_adv"]), tag
        assert np.array_equal(ret, g[tag + "_ret"]), tag


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    assert O.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def run_pid_scenarios(step_fn, reset_fn, g, idx):
    """Replays the reference's PID evaluation (evaluate_controller.py:155-215) for the scenarios in idx, all in
    lock step.  reset_fn(state21, target3) -> obs [n,14] f64; step_fn(actions [n,3] f64) -> obs, rew, done."""
    n = len(idx)
    pid = P.BatchPID(n)
    state = P.scenario_state21(g["init_state"][idx])
    tgt = g["init_target"][idx].copy()
    obs = reset_fn(state, tgt)
    pid.set_reference(tgt)
    first = g["first_ref"][idx]
    has_first = ~np.isnan(first[:, 0])
    pid.ref[has_first] = first[has_first]           # stale reference of the harness for the first action
    rewards = np.zeros((n, 1500))
    length = np.zeros(n, dtype=np.int64)
    alive = np.ones(n, bool)
    for t in range(1500):
        a = pid.get_action(obs[:, 0], obs[:, 1], obs[:, 2], obs[:, 3:6])
        obs, rew, done = step_fn(a)
        rewards[alive, t] = rew[alive]
        newly = alive & done
        length[newly] = t + 1
        alive &= ~done
        pid.set_reference(obs[:, 6:9])                # info["target"] == the target entries of the observation
        if not alive.any():
            break
    return rewards, length


def test_oracle_reproduces_reference_golden_pid_evaluation():
    g = load_golden("pid_none")
    idx = np.arange(100)
    cfg = build_config(config_kw=P.PID_EVAL_CONFIG_KW, sim_config_kw=P.PID_EVAL_SIM_KW)
    envs = [O.OracleEnv(cfg) for _ in idx]
    fin = {}

    def reset_fn(state, tgt):
        return np.stack([e.reset(state[i], tgt[i]) for i, e in enumerate(envs)])

    def step_fn(a):
        out = []
        for i, e in enumerate(envs):
            if i in fin:
                out.append(fin[i])
                continue
            o, r, d, term = e.step(a[i])
            if d:
                fin[i] = (o, r, True)
            out.append((o, r, d))
        return (np.stack([x[0] for x in out]), np.array([x[1] for x in out]), np.array([x[2] for x in out]))

    rewards, length = run_pid_scenarios(step_fn, reset_fn, g, idx)
    # (a) against our own run of the reference in this container: tight
    assert np.array_equal(length, g["live_len"])
    for s in idx:
        n = length[s]
        assert np.abs(rewards[s, :n] - g["live_rewards"][s, :n]).max() < 1e-8, s
    # (b) against the reference's own golden file (produced under the pinned scipy 1.6 / numpy 1.19 stack):
    # episode lengths exact, reward traces to 1e-5 (the live reference itself agrees to 8.5e-6).  The last six
    # golden traces carry extra entries appended after the scenario ended (harness artefact): compare the prefix.
    for s in idx:
        n = length[s]
        if s < 94:
            assert n == g["gold_len"][s], s
        assert np.abs(rewards[s, :n] - g["gold_rewards"][s, :n]).max() < 2e-5, s
    # (c) integer metrics: the golden file lists them in completion order -> compare as multisets
    m = np.stack([envs[i].metrics()[0] for i in idx])
    for name, lo, keys in (("settling_time", 3, ("roll", "pitch", "Va", "all")), ("rise_time", 0, ("roll", "pitch", "Va")),
                           ("success", 17, ("roll", "pitch", "Va", "all"))):
        gk = [str(k) for k in g["gold_%s_keys" % name]]
        cols = [lo + keys.index(k) for k in gk]
        ours = np.sort(np.nan_to_num(m[:, cols], nan=-1), axis=0)
        gold = np.sort(np.nan_to_num(g["gold_" + name], nan=-1), axis=0)
        assert np.array_equal(ours, gold), name
    assert np.abs(np.sort(m[:, 16]) - np.sort(g["gold_control_variation"][:, 0])).max() < 1e-4


def test_oracle_general_observation_layout_cnn_config():
    """examples/models/cnn_controller/fixed_wing_config.json: 5 x 12 observation matrix with history rows, relative
    targets and the init_noise offset (pinned to u = 0.25 in the fixture run and here)."""
    from conftest import cnn_env_config
    g = load_golden("traj_cnn_obs")
    cfg = build_config(env_cfg=cnn_env_config(), sim_config_kw={"turbulence": False}, obs_init_noise=0.25)
    assert cfg.obs_generic == 1 and (cfg.obs_len, cfg.obs_n) == (5, 12)
    for ep in range(g["actions"].shape[0]):
        env = O.OracleEnv(cfg)
        obs = env.reset(g["init_state"][ep], g["init_target"][ep])
        assert obs.shape == (60,) and np.abs(obs - g["obs0"][ep]).max() < 1e-12
        for t in range(int(g["n_valid"][ep])):
            obs, rew, done, term = env.step(g["actions"][ep, t])
            assert _rel(obs, g["obs"][ep, t]).max() < 1e-9, (ep, t)
            assert abs(rew - g["reward"][ep, t]) < 1e-9 and done == bool(g["done"][ep, t])


def test_oracle_general_reward_engine():
    """Potential-form reward with three terms and every factor class (tests/golden/make_golden.py:gen_reward), f64 and
    float32 actions.  ONE env object across all episodes, like the fixture run: `goal_achieved` is never cleared by
    reset (fixed_wing.py:81), so the success bonus fires once in the env's lifetime."""
    from conftest import rich_reward_env_config
    cfg = build_config(env_cfg=rich_reward_env_config(), sim_config_kw={"turbulence": False})
    assert cfg.rew_generic == 1 and cfg.rew_n == 11 and cfg.rew_potential == 1
    env = O.OracleEnv(cfg)
    for name, f32 in (("traj_reward_rich", False), ("traj_reward_rich_f32", True)):
        g = load_golden(name)
        for ep in range(g["actions"].shape[0]):
            env.reset(g["init_state"][ep], g["init_target"][ep])
            for t in range(int(g["n_valid"][ep])):
                obs, rew, done, term = env.step(g["actions"][ep, t], f32)
                assert abs(rew - g["reward"][ep, t]) < 1e-9 * max(1.0, abs(g["reward"][ep, t])), (name, ep, t)
                assert done == bool(g["done"][ep, t])


def test_oracle_error_integrals_and_strided_observation_rows():
    """`integrator` observation entries, `int_error` reward factors (fixed_wing.py:1003-1012, 1165-1180) and
    observation.step = 2 against live-reference runs: window 4 with rows at lags 1 / 3 / 5, and the window of 0 that every
    config of the reference tree carries (the reward integral then runs over the whole history).  ONE env object across
    the episodes: the reset observation reads the error history of the episode that just ended (None on the first)."""
    from conftest import INTEGRATOR_CASES, integrator_env_config
    for name, W, L, step in INTEGRATOR_CASES:
        g = load_golden(name)
        cfg = build_config(env_cfg=integrator_env_config(W, L, step), sim_config_kw={"turbulence": False}, obs_init_noise=0.25)
        assert cfg.obs_generic == 1 and cfg.rew_generic == 1 and (cfg.obs_len, cfg.obs_n, cfg.obs_step) == (L, 15, step)
        env = O.OracleEnv(cfg)
        for ep in range(g["actions"].shape[0]):
            obs = env.reset(g["init_state"][ep], g["init_target"][ep])
            assert np.abs(obs - g["obs0"][ep]).max() < 1e-12, (name, ep)
            for t in range(int(g["n_valid"][ep])):
                obs, rew, done, term = env.step(g["actions"][ep, t])
                assert _rel(obs, g["obs"][ep, t]).max() < 1e-9, (name, ep, t)
                assert abs(rew - g["reward"][ep, t]) < 1e-9 * max(1.0, abs(g["reward"][ep, t])), (name, ep, t)
                assert done == bool(g["done"][ep, t])


def test_oracle_attitude_angular_targets():
    """Target class attitude_angular (fixed_wing.py:671-675, 1455-1460, 1558-1642): omega_p/q/r as derived target states
    in observations, error and goal rewards, the success streak and the per-state metrics, against a live-reference run
    (env RNG replaced by fixed draws, u = 0.625: the rate targets sampled at reset derive from the SAMPLED attitude
    targets before the injected ones override them)."""
    from conftest import angular_env_config, angular_metric_rows, close_or_both_nan
    g = load_golden("traj_angular")
    cfg = build_config(env_cfg=angular_env_config(), sim_config_kw={"turbulence": False}, rng_u_override=0.625)
    assert cfg.ang_on == 1 and cfg.rew_generic == 1 and cfg.obs_generic == 1 and cfg.obs_n == 16
    env = O.OracleEnv(cfg)
    row = 0
    for ep in range(g["actions"].shape[0]):
        obs = env.reset(g["init_state"][ep], g["init_target"][ep])
        assert np.abs(obs - g["obs0"][ep]).max() < 1e-12, ep
        assert np.abs(env.angular()[0] - g["target0"][ep, 3:]).max() < 1e-12, ep
        for t in range(int(g["n_valid"][ep])):
            obs, rew, done, term = env.step(g["actions"][ep, t])
            assert _rel(obs, g["obs"][ep, t]).max() < 1e-9, (ep, t)
            assert abs(rew - g["reward"][ep, t]) < 1e-9 * max(1.0, abs(g["reward"][ep, t])), (ep, t)
            assert _rel(env.angular()[0], g["target"][ep, t, 3:]).max() < 1e-9, (ep, t)
            assert done == bool(g["done"][ep, t]) and (not done or term == int(g["term"][ep, t]))
        if done:
            ours = angular_metric_rows(env.metrics()[0], env.angular()[1])
            assert close_or_both_nan(ours, g["metrics52"][row], 1e-9, 1e-12).all(), (ep, ours, g["metrics52"][row])
            row += 1
    assert row == len(g["metrics52"])


def test_oracle_moving_target_classes():
    """Target classes linear / sinusoidal with Va compensate on a sinusoidal pitch target, against a live-reference run
    whose env-level RNG was replaced by fixed draws u (tests/golden/make_golden.py:gen_targets)."""
    from conftest import moving_targets_env_config
    g = load_golden("traj_moving_targets")
    for tag in ("a", "b"):
        cfg = build_config(env_cfg=moving_targets_env_config(), sim_config_kw={"turbulence": False},
                           rng_u_override=float(g[tag + "_u"]))
        assert list(cfg.tgt_class) == [2, 3, 1]
        for ep in range(3):
            env = O.OracleEnv(cfg)
            obs = env.reset(g[tag + "_init_state"][ep])
            assert np.abs(env.get()["target"] - g[tag + "_target0"][ep]).max() < 1e-12
            assert np.abs(obs - g[tag + "_obs0"][ep]).max() < 1e-12
            for t in range(150):
                obs, rew, done, term = env.step(g[tag + "_actions"][ep, t])
                assert _rel(env.get()["target"], g[tag + "_target"][ep, t]).max() < 1e-11, (tag, ep, t)
                assert _rel(obs, g[tag + "_obs"][ep, t]).max() < 1e-9
                assert abs(rew - g[tag + "_reward"][ep, t]) < 1e-9


def test_oracle_target_resampling_against_reference():
    """resample_every = 37 and on_success = "new" (5-step streak) with the default target classes, against a
    live-reference run whose env-level RNG was replaced by fixed draws u (make_golden.py:gen_resample): target
    trajectories incl. the Va-compensate law restarting after every resample, observations, rewards."""
    from conftest import resample_env_config
    g = load_golden("traj_resample")
    for tag in ("a", "b"):
        cfg = build_config(env_cfg=resample_env_config(), sim_config_kw={"turbulence": False},
                           rng_u_override=float(g[tag + "_u"]))
        assert cfg.resample_every == 37 and cfg.on_success == 2 and cfg.streak_req == 5
        jumps = 0
        for ep in range(3):
            env = O.OracleEnv(cfg)
            obs = env.reset(g[tag + "_init_state"][ep])
            assert np.abs(env.get()["target"] - g[tag + "_target0"][ep]).max() < 1e-12
            assert np.abs(obs - g[tag + "_obs0"][ep]).max() < 1e-12
            prev = g[tag + "_target0"][ep]
            for t in range(150):
                if not np.isfinite(g[tag + "_reward"][ep, t]):
                    break
                obs, rew, done, term = env.step(g[tag + "_actions"][ep, t])
                assert _rel(env.get()["target"], g[tag + "_target"][ep, t]).max() < 1e-11, (tag, ep, t)
                assert _rel(obs, g[tag + "_obs"][ep, t]).max() < 1e-9
                assert abs(rew - g[tag + "_reward"][ep, t]) < 1e-9
                jumps += int(abs(g[tag + "_target"][ep, t, 2] - prev[2]) > 0.05)
                prev = g[tag + "_target"][ep, t]
        assert jumps >= 6                                  # the fixture does exercise resampling


def test_dryden_block_restart_matches_reference():
    """pyfly re-simulates turbulence every turbulence_sim_length samples (300 for raw pyfly / the waypoint env); lsim then
    restarts from T[0] > 0 and decays the carried state over [0, T[0]] first."""
    g = load_golden("dryden")
    cfg = build_config(env_kind="waypoint", sim_config_kw={"turbulence": True, "turbulence_intensity": "moderate"})
    assert cfg.turb_block_len == 300
    out = O.dryden(cfg, g["blocks_noise"])
    ref = np.concatenate([g["blocks_lin"], g["blocks_ang"]])
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.all(out[:, 300] == 0) and np.abs(out[:, 301]).max() > 0


def test_oracle_waypoint_env_head():
    """FixedWingAircraft_simple (magpie/magpy/simple_train.py:197-702) run in place by make_golden.py:gen_waypoint:
    observations, rewards and the teleports from leg to leg."""
    g = load_golden("traj_waypoint")
    cfg = build_config(env_kind="waypoint", sim_config_kw={"turbulence": False})
    assert cfg.steps_max == 500 and cfg.scale_actions == 0 and cfg.va_con_max == 0
    for t in range(3):
        env = O.OracleEnv(cfg)
        env.set_waypoint_tasks(g["tasks"], t)
        obs = env.reset()
        assert obs.shape == (12,) and np.abs(obs - g["obs0_%d" % t]).max() < 1e-12
        for k in range(120):
            obs, rew, done, term = env.step(g["actions_%d" % t][k])
            assert _rel(obs, g["obs_%d" % t][k]).max() < 1e-9, (t, k)
            assert abs(rew - g["reward_%d" % t][k]) < 1e-9 and done == bool(g["done_%d" % t][k])


def test_oracle_against_the_live_reference_on_fresh_episodes():
    """Where the reference can be imported (the build container, or its mirror baseline/_ref), tools/oracle_sweep.py draws
    NEW episodes in nine categories (calm / windy / out-of-range initial states, clipped and f32 actions, the three
    turbulence intensities) and then under EVERY config file the reference ships (general observation layouts, reward
    variants), steps the live reference (fixed_wing.py:483-652) and the oracle side by side and applies the checks of
    the fixture test above: exact done / termination / RK45 RHS count, <= 1e-9 relative elsewhere.  In its own
    process: the reference needs the fabricated gym / matplotlib modules of oracle/refshim."""
    import os
    import subprocess
    import sys
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference libraries neither mounted nor mirrored under baseline/_ref")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "oracle_sweep.py"), "2", "50", "31337"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and "MISMATCHES: 0" in res.stdout, res.stdout[-2000:]
