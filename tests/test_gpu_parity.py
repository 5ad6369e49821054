"""GPU parity tests proper: libfwb200.so (through the C ABI) against
  (1) fixtures recorded from the live reference (tests/golden/*.npz, made by tests/golden/make_golden.py), and
  (2) the C oracle (oracle/fw_oracle.c) on seeded inputs.
Bars: fp64 exact mode within 1e-9 relative per step (north_star); done / termination codes / RHS-evaluation
counts / integer metrics bit-exact."""
import numpy as np
import pytest

from conftest import TRAJ_CASES, close_or_both_nan, golden_metric_rows, load_golden

pytestmark = pytest.mark.gpu

RTOL_F64 = 1e-9


def _rel(a, b):
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def _run_golden(name, cfg_kw, sim_kw, f32, torch, B):
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    g = load_golden(name)
    E, T = g["actions"].shape[:2]
    cfg = build_config(config_kw=cfg_kw, sim_config_kw=sim_kw)
    env = B(E, cfg=cfg)
    env.enable_f64_outputs()
    noise = torch.as_tensor(g["noise"]) if "noise" in g.files else None
    env.reset(state=g["init_state"], target=g["init_target"], noise=noise)
    out = dict(obs0=env.obs64.cpu().numpy(), y0=env.get_field(bt.FIELD_Y).cpu().numpy())
    rec = {k: [] for k in ("y", "obs", "rew", "done", "term", "nfev", "vab", "tgt", "cmd")}
    metrics = {}
    alive = np.ones(E, bool)
    for t in range(T):
        a = torch.as_tensor(g["actions"][:, t], dtype=torch.float32 if f32 else torch.float64).cuda().contiguous()
        env.step(a, auto_reset=False)
        term, m, ret, ln = env.episode_info()
        rec["y"].append(env.get_field(bt.FIELD_Y).cpu().numpy())
        rec["obs"].append(env.obs64.cpu().numpy())
        rec["rew"].append(env.rew64.cpu().numpy())
        rec["done"].append(env.done.cpu().numpy().astype(bool))
        rec["term"].append(term.cpu().numpy())
        rec["nfev"].append(env.get_field(bt.FIELD_NFEV).cpu().numpy()[:, 0])
        rec["vab"].append(env.get_field(bt.FIELD_VAB).cpu().numpy())
        rec["tgt"].append(env.get_field(bt.FIELD_TARGET).cpu().numpy())
        rec["cmd"].append(env.get_field(bt.FIELD_CMD).cpu().numpy())
        d = rec["done"][-1]
        for ep in np.where(d & alive)[0]:
            metrics[int(ep)] = m[ep].cpu().numpy()
        alive &= ~d
    env.close()
    out.update({k: np.stack(v, axis=1) for k, v in rec.items()})
    out["metrics"] = metrics
    return g, out


@pytest.mark.parametrize("name,cfg_kw,sim_kw,f32", TRAJ_CASES, ids=[c[0] for c in TRAJ_CASES])
def test_reference_fixture_trajectories(name, cfg_kw, sim_kw, f32, cuda_device):
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.batched import BatchedFixedWing
    g, o = _run_golden(name, cfg_kw, sim_kw, f32, torch, BatchedFixedWing)
    E, T = g["actions"].shape[:2]
    assert np.abs(o["y0"] - g["y0"]).max() < 1e-14
    assert np.abs(o["obs0"] - g["obs0"]).max() < 1e-12
    worst = {}
    n_steps = n_same_nfev = 0
    for ep in range(E):
        nv = int(g["n_valid"][ep])
        sl = slice(0, nv)
        ok = g["term"][ep, sl] < 10          # steps where the simulator step succeeded
        # flags: bit exact
        assert np.array_equal(o["done"][ep, sl], g["done"][ep, sl].astype(bool)), (name, ep)
        assert np.array_equal(o["term"][ep, sl], g["term"][ep, sl]), (name, ep)
        for key, ref, got in (("y", g["y"][ep, sl][ok], o["y"][ep, sl][ok]),
                              ("vab", g["vab"][ep, sl][ok], o["vab"][ep, sl][ok]),
                              ("obs", g["obs"][ep, sl], o["obs"][ep, sl]),
                              ("rew", g["reward"][ep, sl], o["rew"][ep, sl]),
                              ("tgt", g["target"][ep, sl], o["tgt"][ep, sl]),
                              ("cmd", g["cmd"][ep, sl], o["cmd"][ep, sl])):
            if ref.size:
                worst[key] = max(worst.get(key, 0.0), float(_rel(got, ref).max()))
        n_steps += nv
        n_same_nfev += int((o["nfev"][ep, sl] == g["nfev"][ep, sl]).sum())
    print("\n[%s] worst rel err %s ; identical RHS-evaluation counts %d/%d" % (
        name, {k: "%.2e" % v for k, v in worst.items()}, n_same_nfev, n_steps))
    for k, v in worst.items():
        assert v < RTOL_F64, (name, k, v)
    assert n_same_nfev == n_steps, "adaptive step-size decisions diverged from the reference"
    if "m_success" in g.files:
        rows, eps = golden_metric_rows(g)
        for row, ep in zip(rows, eps):
            got = o["metrics"][ep]
            assert close_or_both_nan(got, row, 1e-7, 1e-9).all(), (name, ep, got, row)
            for lo, hi in ((0, 7), (17, 21)):     # rise/settling times and success flags are integers: exact
                assert np.array_equal(np.nan_to_num(got[lo:hi], nan=-1), np.nan_to_num(row[lo:hi], nan=-1))


DEV_CONFIG_KW = {"action": {"scale_space": False}, "observation": {"noise": {"mean": 0, "var": 0.1}},
                 "target": {"states": {0: {"bound": 3}, 1: {"bound": 3}}}}      # fixed_wing_config_dev.json


@pytest.mark.parametrize("config_kw,amp", [(None, 1.5), (DEV_CONFIG_KW, 0.4)], ids=["default", "dev_config_obs_noise"])
def test_batch_against_oracle_random_policy(config_kw, amp, cuda_device):
    """4096 envs x 60 steps, Philox resets, turbulence on, float32 actions: CUDA vs C oracle, state-for-state.
    The second case is the reference's fixed_wing_config_dev.json: unscaled actions, observation noise (Philox)."""
    import torch
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    n, T = 4096, 60
    cfg = build_config(config_kw=config_kw, sim_config_kw={"turbulence": True}, seed=1234)
    env = bt.BatchedFixedWing(n, cfg=cfg)
    env.enable_f64_outputs()
    env.reset()
    ob = O.OracleBatch(cfg, n)
    obs_ref = ob.reset().copy()
    assert _rel(env.obs64.cpu().numpy(), obs_ref).max() < 1e-12
    rs = np.random.RandomState(0)
    same = tot = 0
    for t in range(T):
        a = rs.uniform(-amp, amp, (n, 3)).astype(np.float32)
        env.step(torch.as_tensor(a).cuda(), auto_reset=True)
        o_ref, r_ref, d_ref = ob.step(a)
        assert np.array_equal(env.done.cpu().numpy(), d_ref)
        assert _rel(env.obs64.cpu().numpy(), o_ref).max() < RTOL_F64, t
        assert _rel(env.rew64.cpu().numpy(), r_ref).max() < RTOL_F64, t
        # step-size decisions: RHS-evaluation counts of the step that just ran (envs that auto-reset keep the count of
        # the step that ended the episode on both sides)
        nf = env.get_field(bt.FIELD_NFEV).cpu().numpy()
        nf_ref, na_ref = ob.counters()
        same += int((nf[:, 0] == nf_ref).sum() + 0)
        tot += n
    print("\n[oracle batch] identical RHS-evaluation counts on %d/%d env-steps" % (same, tot))
    assert same == tot
    env.close()


def test_gae_bit_exact_vs_reference_fixture(cuda_device):
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.batched import gae
    g = load_golden("gae")
    for tag in ("a", "b", "c"):
        dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
        adv, ret = gae(dev(g[tag + "_rew"]), dev(g[tag + "_val"]), dev(g[tag + "_done"]), dev(g[tag + "_last_val"]),
                       dev(g[tag + "_last_done"].astype(np.uint8)))
        assert np.array_equal(adv.cpu().numpy(), g[tag + "_adv"]), tag
        assert np.array_equal(ret.cpu().numpy(), g[tag + "_ret"]), tag


def test_gae_streaming_kernel_bit_exact_vs_oracle(cuda_device):
    """The cp.async streaming GAE kernel (strips of 32 / 16 / 8 env columns, tiles of 32 / 64 rows, ragged last tile) and
    the plain fallback, against the oracle's restatement of buffers.py:304-333 — every bit of advantages and returns."""
    import torch
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200.batched import gae
    rs = np.random.RandomState(5)
    for T, N in ((300, 1000), (131, 9472), (70, 20480), (2, 64), (1, 64), (65, 1001), (2048, 512)):
        rew, val = rs.standard_normal((T, N)).astype(np.float32), rs.standard_normal((T, N)).astype(np.float32)
        done = (rs.uniform(size=(T, N)) < 0.02).astype(np.float32)
        lv, ld = rs.standard_normal(N).astype(np.float32), (rs.uniform(size=N) < 0.3).astype(np.uint8)
        adv, ret = gae(*(torch.as_tensor(x).cuda() for x in (rew, val, done, lv, ld)))
        adv_ref, ret_ref = O.gae(rew, val, done, lv, ld)
        assert np.array_equal(adv.cpu().numpy(), adv_ref), (T, N)
        assert np.array_equal(ret.cpu().numpy(), ret_ref), (T, N)


def test_general_observation_layout_cnn_config(cuda_device):
    """The reference's CNN-controller config (5 x 12 observation matrix: history rows, relative targets, action
    windows shifted in time) on the CUDA path against the live-reference fixture and, with Philox init_noise and
    observation noise switched on, against the oracle."""
    import torch
    from conftest import cnn_env_config
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    g = load_golden("traj_cnn_obs")
    E, T = g["actions"].shape[:2]
    cfg = build_config(env_cfg=cnn_env_config(), sim_config_kw={"turbulence": False}, obs_init_noise=0.25)
    env = bt.BatchedFixedWing(E, cfg=cfg)
    assert env.obs_dim == 60
    env.enable_f64_outputs()
    env.reset(state=g["init_state"], target=g["init_target"])
    assert np.abs(env.obs64.cpu().numpy() - g["obs0"]).max() < 1e-12
    for t in range(T):
        env.step(torch.as_tensor(g["actions"][:, t]).cuda().contiguous(), auto_reset=False)
        assert _rel(env.obs64.cpu().numpy(), g["obs"][:, t]).max() < RTOL_F64, t
        assert _rel(env.rew64.cpu().numpy(), g["reward"][:, t]).max() < RTOL_F64
    env.close()
    # random resets / Philox init_noise / observation noise / auto-reset: CUDA vs oracle
    ecfg = cnn_env_config()
    ecfg["observation"]["noise"] = {"mean": 0.0, "var": 0.05}
    ecfg["steps_max"] = 12
    cfg = build_config(env_cfg=ecfg, sim_config_kw={"turbulence": True}, seed=9)
    n = 512
    env = bt.BatchedFixedWing(n, cfg=cfg)
    env.enable_f64_outputs()
    env.reset()
    ob = O.OracleBatch(cfg, n)
    assert _rel(env.obs64.cpu().numpy(), ob.reset()).max() < 1e-12
    rs = np.random.RandomState(3)
    for t in range(30):
        a = rs.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
        env.step(torch.as_tensor(a).cuda(), auto_reset=True)
        o_ref, r_ref, d_ref = ob.step(a)
        assert np.array_equal(env.done.cpu().numpy(), d_ref)
        assert _rel(env.obs64.cpu().numpy(), o_ref).max() < RTOL_F64, t
    env.close()


def test_general_reward_engine(cuda_device):
    """The general reward engine on the CUDA path: (a) the live-reference fixture replayed on a ONE-env handle across
    episodes (goal_achieved persists like in the reference), (b) 1024 envs with auto-reset against the oracle."""
    import torch
    from conftest import rich_reward_env_config
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    cfg = build_config(env_cfg=rich_reward_env_config(), sim_config_kw={"turbulence": False})
    env = bt.BatchedFixedWing(1, cfg=cfg)
    env.enable_f64_outputs()
    for name, f32 in (("traj_reward_rich", False), ("traj_reward_rich_f32", True)):
        g = load_golden(name)
        for ep in range(g["actions"].shape[0]):
            env.reset(state=g["init_state"][ep:ep + 1], target=g["init_target"][ep:ep + 1])
            for t in range(int(g["n_valid"][ep])):
                a = torch.as_tensor(g["actions"][ep:ep + 1, t], dtype=torch.float32 if f32 else torch.float64).cuda()
                env.step(a.contiguous(), auto_reset=False)
                r = float(env.rew64.cpu().numpy()[0])
                assert abs(r - g["reward"][ep, t]) < 1e-9 * max(1.0, abs(g["reward"][ep, t])), (name, ep, t)
    env.close()
    cfg = build_config(env_cfg=rich_reward_env_config(), sim_config_kw={"turbulence": True}, seed=21)
    n = 1024
    env = bt.BatchedFixedWing(n, cfg=cfg)
    env.enable_f64_outputs()
    env.reset()
    ob = O.OracleBatch(cfg, n)
    ob.reset()
    rs = np.random.RandomState(8)
    for t in range(170):                      # > 2 episodes of 80 steps: prev_shaping / goal_achieved across resets
        a = rs.uniform(-1.6, 1.6, (n, 3)).astype(np.float32)
        env.step(torch.as_tensor(a).cuda(), auto_reset=True)
        o_ref, r_ref, d_ref = ob.step(a)
        assert np.array_equal(env.done.cpu().numpy(), d_ref), t
        assert _rel(env.rew64.cpu().numpy(), r_ref).max() < RTOL_F64, t
    env.close()


def test_error_integrals_and_strided_observation_rows(cuda_device):
    """`integrator` observation entries, `int_error` reward factors (fixed_wing.py:1003-1012, 1165-1180) and
    observation.step = 2 on the CUDA path: (a) the live-reference fixtures replayed on a ONE-env handle across episodes —
    the reset observation of an integrator entry reads the error history of the episode that just ended (error * window
    on the very first reset); (b) 512 envs with Philox resets, turbulence and auto-reset against the oracle, where the
    value of the ended episode is added to the precomputed next-episode observation."""
    import torch
    from conftest import INTEGRATOR_CASES, integrator_env_config
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    for name, W, L, step in INTEGRATOR_CASES:
        g = load_golden(name)
        cfg = build_config(env_cfg=integrator_env_config(W, L, step), sim_config_kw={"turbulence": False}, obs_init_noise=0.25)
        env = bt.BatchedFixedWing(1, cfg=cfg)
        assert env.obs_dim == L * 15
        env.enable_f64_outputs()
        for ep in range(g["actions"].shape[0]):
            env.reset(state=g["init_state"][ep:ep + 1], target=g["init_target"][ep:ep + 1])
            assert np.abs(env.obs64.cpu().numpy()[0] - g["obs0"][ep]).max() < 1e-12, (name, ep)
            for t in range(int(g["n_valid"][ep])):
                env.step(torch.as_tensor(g["actions"][ep:ep + 1, t]).cuda().contiguous(), auto_reset=False)
                assert _rel(env.obs64.cpu().numpy()[0], g["obs"][ep, t]).max() < RTOL_F64, (name, ep, t)
                r = float(env.rew64.cpu().numpy()[0])
                assert abs(r - g["reward"][ep, t]) < 1e-9 * max(1.0, abs(g["reward"][ep, t])), (name, ep, t)
        env.close()
        ecfg = integrator_env_config(W, L, step)
        ecfg["steps_max"] = 14
        cfg = build_config(env_cfg=ecfg, sim_config_kw={"turbulence": True}, seed=33)
        n = 512
        env = bt.BatchedFixedWing(n, cfg=cfg)
        env.enable_f64_outputs()
        env.reset()
        ob = O.OracleBatch(cfg, n)
        assert _rel(env.obs64.cpu().numpy(), ob.reset()).max() < 1e-12
        rs = np.random.RandomState(5)
        for t in range(40):                       # almost three episodes of 14 steps, plus early failures
            a = rs.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
            env.step(torch.as_tensor(a).cuda(), auto_reset=True)
            o_ref, r_ref, d_ref = ob.step(a)
            assert np.array_equal(env.done.cpu().numpy(), d_ref), (name, t)
            assert _rel(env.obs64.cpu().numpy(), o_ref).max() < RTOL_F64, (name, t)
            assert _rel(env.rew64.cpu().numpy(), r_ref).max() < RTOL_F64, (name, t)
        env.close()


def test_attitude_angular_targets(cuda_device):
    """Target class attitude_angular on the CUDA path (omega_p/q/r as derived target states: observations, error and goal
    rewards, success streak, per-state metrics): (a) the live-reference fixture on a ONE-env handle, (b) 512 envs with
    Philox resets, turbulence, on_success = "new" resampling and auto-reset against the oracle, metrics included."""
    import torch
    from conftest import angular_env_config, angular_metric_rows, close_or_both_nan
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    g = load_golden("traj_angular")
    cfg = build_config(env_cfg=angular_env_config(), sim_config_kw={"turbulence": False}, rng_u_override=0.625)
    env = bt.BatchedFixedWing(1, cfg=cfg)
    assert env.obs_dim == 16
    env.enable_f64_outputs()
    row = 0
    for ep in range(g["actions"].shape[0]):
        env.reset(state=g["init_state"][ep:ep + 1], target=g["init_target"][ep:ep + 1])
        assert np.abs(env.obs64.cpu().numpy()[0] - g["obs0"][ep]).max() < 1e-12, ep
        assert np.abs(env.get_field(bt.FIELD_ATARGET).cpu().numpy()[0] - g["target0"][ep, 3:]).max() < 1e-12
        for t in range(int(g["n_valid"][ep])):
            env.step(torch.as_tensor(g["actions"][ep:ep + 1, t]).cuda().contiguous(), auto_reset=False)
            assert _rel(env.obs64.cpu().numpy()[0], g["obs"][ep, t]).max() < RTOL_F64, (ep, t)
            r = float(env.rew64.cpu().numpy()[0])
            assert abs(r - g["reward"][ep, t]) < 1e-9 * max(1.0, abs(g["reward"][ep, t])), (ep, t)
            assert _rel(env.get_field(bt.FIELD_ATARGET).cpu().numpy()[0], g["target"][ep, t, 3:]).max() < RTOL_F64, (ep, t)
            assert bool(env.done.cpu().numpy()[0]) == bool(g["done"][ep, t])
        if bool(g["done"][ep, int(g["n_valid"][ep]) - 1]):
            term, m28, ret, ln = env.episode_info()
            ours = angular_metric_rows(m28.cpu().numpy()[0], env.episode_info_angular().cpu().numpy()[0])
            assert close_or_both_nan(ours, g["metrics52"][row], 1e-9, 1e-12).all(), (ep, ours, g["metrics52"][row])
            row += 1
    env.close()
    ecfg = angular_env_config()
    ecfg["steps_max"] = 25
    ecfg["target"]["on_success"] = "new"
    cfg = build_config(env_cfg=ecfg, sim_config_kw={"turbulence": True}, seed=77)
    n = 512
    env = bt.BatchedFixedWing(n, cfg=cfg)
    env.enable_f64_outputs()
    env.reset()
    ob = O.OracleBatch(cfg, n)
    assert _rel(env.obs64.cpu().numpy(), ob.reset()).max() < 1e-12
    rs = np.random.RandomState(6)
    for t in range(60):
        a = rs.uniform(-1.4, 1.4, (n, 3)).astype(np.float32)
        env.step(torch.as_tensor(a).cuda(), auto_reset=True)
        o_ref, r_ref, d_ref = ob.step(a)
        assert np.array_equal(env.done.cpu().numpy(), d_ref), t
        assert _rel(env.obs64.cpu().numpy(), o_ref).max() < RTOL_F64, t
        assert _rel(env.rew64.cpu().numpy(), r_ref).max() < RTOL_F64, t
        if d_ref.any():
            am = env.episode_info_angular().cpu().numpy()
            for i in np.flatnonzero(d_ref)[:8]:
                ref = ob.env_angular(int(i))[1]
                assert close_or_both_nan(am[i], ref, 1e-9, 1e-12).all(), (t, i, am[i], ref)
        if t == 30:                 # checkpoint: the state blob carries the rate targets' rings and metrics
            blob = env.get_state()
            env.close()
            env = bt.BatchedFixedWing(n, cfg=cfg)
            env.enable_f64_outputs()
            env.set_state(blob)
    env.close()


def test_attitude_angular_targets_with_the_default_observation_and_reward(cuda_device):
    """The default observation row and reward family plus three rate targets (one without a bound): the general head
    instantiation with obs_generic = 0, against the oracle — done flags (the bounded rate targets gate the success
    streak), observations, rewards and the rate targets' metrics."""
    import torch
    from conftest import close_or_both_nan
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config, default_env_config
    ecfg = default_env_config()
    for name, b in zip(("omega_p", "omega_q", "omega_r"), (0.6, None, 0.5)):
        st = {"name": name, "class": "attitude_angular"}
        if b is not None:
            st["bound"] = b
        ecfg["target"]["states"].append(st)
    ecfg["steps_max"] = 20
    cfg = build_config(env_cfg=ecfg, sim_config_kw={"turbulence": True}, seed=11)
    assert (cfg.obs_generic, cfg.rew_generic, cfg.ang_on) == (0, 1, 1)
    n = 512
    env = bt.BatchedFixedWing(n, cfg=cfg)
    env.enable_f64_outputs()
    env.reset()
    ob = O.OracleBatch(cfg, n)
    assert np.abs(env.obs64.cpu().numpy() - ob.reset()).max() < 1e-12
    rs = np.random.RandomState(1)
    for t in range(50):
        a = rs.uniform(-1.3, 1.3, (n, 3)).astype(np.float32)
        env.step(torch.as_tensor(a).cuda())
        o_ref, r_ref, d_ref = ob.step(a)
        assert np.array_equal(env.done.cpu().numpy(), d_ref), t
        assert _rel(env.obs64.cpu().numpy(), o_ref).max() < RTOL_F64, t
        assert _rel(env.rew64.cpu().numpy(), r_ref).max() < RTOL_F64, t
        if d_ref.any():
            am = env.episode_info_angular().cpu().numpy()
            for i in np.flatnonzero(d_ref)[:4]:
                assert close_or_both_nan(am[i], ob.env_angular(int(i))[1], 1e-9, 1e-12).all(), (t, i)
    env.close()


def test_moving_target_classes(cuda_device):
    """linear / sinusoidal targets on the CUDA path: the live-reference fixture (fixed draws), then Philox sampling with
    on_success = "new" resampling against the oracle."""
    import copy
    import torch
    from conftest import moving_targets_env_config
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    g = load_golden("traj_moving_targets")
    for tag in ("a", "b"):
        cfg = build_config(env_cfg=moving_targets_env_config(), sim_config_kw={"turbulence": False},
                           rng_u_override=float(g[tag + "_u"]))
        env = bt.BatchedFixedWing(3, cfg=cfg)
        env.enable_f64_outputs()
        env.reset(state=g[tag + "_init_state"])
        assert np.abs(env.get_field(bt.FIELD_TARGET).cpu().numpy() - g[tag + "_target0"]).max() < 1e-12
        for t in range(150):
            env.step(torch.as_tensor(g[tag + "_actions"][:, t]).cuda().contiguous(), auto_reset=False)
            assert _rel(env.get_field(bt.FIELD_TARGET).cpu().numpy(), g[tag + "_target"][:, t]).max() < 1e-11, (tag, t)
            assert _rel(env.obs64.cpu().numpy(), g[tag + "_obs"][:, t]).max() < RTOL_F64
            assert _rel(env.rew64.cpu().numpy(), g[tag + "_reward"][:, t]).max() < RTOL_F64
        env.close()
    ecfg = copy.deepcopy(moving_targets_env_config())
    ecfg["steps_max"] = 60
    ecfg["target"].update(on_success="new", success_streak_req=5, success_streak_fraction=0.6)
    for s_, b in zip(ecfg["target"]["states"], (80, 50, 15)):
        s_["bound"] = b
    cfg = build_config(env_cfg=ecfg, sim_config_kw={"turbulence": True}, seed=77)
    n = 512
    env = bt.BatchedFixedWing(n, cfg=cfg)
    env.enable_f64_outputs()
    env.reset()
    ob = O.OracleBatch(cfg, n)
    assert _rel(env.obs64.cpu().numpy(), ob.reset()).max() < 1e-12
    rs = np.random.RandomState(5)
    for t in range(130):
        a = rs.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
        env.step(torch.as_tensor(a).cuda(), auto_reset=True)
        o_ref, r_ref, d_ref = ob.step(a)
        assert np.array_equal(env.done.cpu().numpy(), d_ref), t
        assert _rel(env.obs64.cpu().numpy(), o_ref).max() < RTOL_F64, t
        assert _rel(env.rew64.cpu().numpy(), r_ref).max() < RTOL_F64, t
    env.close()


def test_target_resampling_against_reference(cuda_device):
    """resample_every / on_success = "new" on the CUDA path against the live-reference fixture (fixed draws u)."""
    import torch
    from conftest import resample_env_config
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    g = load_golden("traj_resample")
    for tag in ("a", "b"):
        cfg = build_config(env_cfg=resample_env_config(), sim_config_kw={"turbulence": False},
                           rng_u_override=float(g[tag + "_u"]))
        env = bt.BatchedFixedWing(3, cfg=cfg)
        env.enable_f64_outputs()
        env.reset(state=g[tag + "_init_state"])
        assert np.abs(env.get_field(bt.FIELD_TARGET).cpu().numpy() - g[tag + "_target0"]).max() < 1e-12
        assert np.isfinite(g[tag + "_reward"]).all()       # no episode of the fixture ends early
        for t in range(150):
            env.step(torch.as_tensor(g[tag + "_actions"][:, t]).cuda().contiguous(), auto_reset=False)
            assert _rel(env.get_field(bt.FIELD_TARGET).cpu().numpy(), g[tag + "_target"][:, t]).max() < 1e-11, (tag, t)
            assert _rel(env.obs64.cpu().numpy(), g[tag + "_obs"][:, t]).max() < RTOL_F64
            assert _rel(env.rew64.cpu().numpy(), g[tag + "_reward"][:, t]).max() < RTOL_F64
        env.close()


def test_full_size_batch_properties(cuda_device):
    """BASELINE.json's C3 size (65 536 envs on one GPU, turbulence, auto-reset), checked through properties that do not
    depend on the size: (1) every env's trajectory is independent of the batch it runs in — a 192-env slice of the big
    batch equals, bit for bit, a 192-env handle with the same global env ids, which in turn matches the C oracle;
    (2) the persistent-lane work queue hands envs to lanes in a run-dependent order, yet two runs agree bit for bit;
    (3) RHS evaluations = 2 + 6 x attempts for every env-step."""
    import torch
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    n, T, k0, m = 65536, 45, 40000, 192
    kw = dict(config_kw={"steps_max": 17}, sim_config_kw={"turbulence": True}, seed=77)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = [(torch.rand(n, 3, device="cuda", generator=g) * 3 - 1.5).contiguous() for _ in range(T)]
    sums = []
    for run in range(2):
        big = bt.BatchedFixedWing(n, cfg=build_config(**kw))
        big.reset()
        small = ob = None
        if run == 0:
            cfg_s = build_config(env_id_offset=k0, **kw)
            small = bt.BatchedFixedWing(m, cfg=cfg_s)
            small.enable_f64_outputs()
            small.reset()
            ob = O.OracleBatch(cfg_s, m)
            ob.reset()
            assert torch.equal(big.obs[k0:k0 + m], small.obs)
        acc = torch.zeros((), dtype=torch.float64, device="cuda")
        ends = 0
        for t in range(T):
            big.step(acts[t])
            acc += big.obs.double().sum() + big.rew.double().sum() * 3.0 + big.done.double().sum() * 7.0
            nf = big.get_field(bt.FIELD_NFEV)
            assert bool((nf[:, 0] == 2 + 6 * nf[:, 1]).all())
            ends += int(big.done.sum())
            if small is not None:
                a = acts[t][k0:k0 + m].contiguous()
                small.step(a)
                assert torch.equal(big.obs[k0:k0 + m], small.obs) and torch.equal(big.rew[k0:k0 + m], small.rew)
                assert torch.equal(big.done[k0:k0 + m], small.done), t
                o_ref, r_ref, d_ref = ob.step(a.cpu().numpy())
                assert np.array_equal(small.done.cpu().numpy(), d_ref)
                assert _rel(small.obs64.cpu().numpy(), o_ref).max() < RTOL_F64, t
        assert ends >= 2 * n                       # every env went through at least two auto-resets
        sums.append(float(acc))
        big.close()
    assert sums[0] == sums[1]


@pytest.mark.parametrize("dist", ["gaussian", "uniform"])
def test_aircraft_parameter_randomisation(dist, cuda_device):
    """simulator.model (FixedWingAircraft.sample_simulator_parameters, fixed_wing.py:748-813) on the CUDA path:
    (a) with the parameters the LIVE reference drew at each reset injected (FW_FIELD_PARAMS), the reference's
        trajectories are reproduced with identical RK45 decisions (tests/golden/traj_model_*.npz);
    (b) the device's own Philox draws equal the oracle's, parameter for parameter, through auto-resets, and the two then
        fly the same trajectories; a handle without the block refuses FW_FIELD_PARAMS."""
    import torch
    from conftest import model_env_config
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200._lib import FwError
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    g = load_golden("traj_model_" + dist)
    E, T = g["actions"].shape[:2]
    cfg = build_config(env_cfg=model_env_config(dist), sim_config_kw={"turbulence": True})
    env = bt.BatchedFixedWing(E, cfg=cfg)
    env.enable_f64_outputs()
    env.reset(state=g["init_state"], target=g["init_target"], noise=g["noise"])
    assert np.abs(env.obs64.cpu().numpy() - g["obs0"]).max() < 1e-12
    env.set_field(bt.FIELD_PARAMS, g["params"])
    assert np.array_equal(env.get_field(bt.FIELD_PARAMS).cpu().numpy(), g["params"])
    alive = np.ones(E, bool)
    for t in range(T):
        env.step(torch.as_tensor(g["actions"][:, t]).cuda().contiguous(), auto_reset=False)
        y, obs = env.get_field(bt.FIELD_Y).cpu().numpy(), env.obs64.cpu().numpy()
        nf = env.get_field(bt.FIELD_NFEV).cpu().numpy()[:, 0]
        for ep in np.flatnonzero(alive):
            if t >= int(g["n_valid"][ep]):
                alive[ep] = False
                continue
            assert nf[ep] == int(g["nfev"][ep, t]), (ep, t)
            assert (np.abs(y[ep] - g["y"][ep, t]) / np.maximum(1, np.abs(g["y"][ep, t]))).max() < 1e-9, (ep, t)
            assert (np.abs(obs[ep] - g["obs"][ep, t]) / np.maximum(1, np.abs(g["obs"][ep, t]))).max() < 1e-9, (ep, t)
    env.close()
    # (b) own sampling against the oracle, across episode ends
    n = 512
    cfg2 = build_config(env_cfg=model_env_config(dist), sim_config_kw={"turbulence": True}, config_kw={"steps_max": 12}, seed=21)
    env = bt.BatchedFixedWing(n, cfg=cfg2)
    env.enable_f64_outputs()
    env.reset()
    ob = O.OracleBatch(cfg2, n)
    ob.reset()
    rs = np.random.RandomState(3)
    for t in range(30):
        pr = ob.params()
        assert (np.abs(env.get_field(bt.FIELD_PARAMS).cpu().numpy() - pr) / np.maximum(1e-30, np.abs(pr))).max() < 1e-12, t
        a = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
        env.step(torch.as_tensor(a).cuda())
        o_ref, r_ref, d_ref = ob.step(a)
        assert np.array_equal(env.done.cpu().numpy(), d_ref), t
        assert (np.abs(env.obs64.cpu().numpy() - o_ref) / np.maximum(1.0, np.abs(o_ref))).max() < 1e-9, t
    first = ob.params()
    assert len(np.unique(first[:, 17])) > 0.8 * n                      # C_L_alpha: every env its own draw
    blob = env.get_state()
    twin = bt.BatchedFixedWing(n, cfg=cfg2)
    twin.set_state(blob)
    assert torch.equal(twin.get_field(bt.FIELD_PARAMS), env.get_field(bt.FIELD_PARAMS))     # the blob carries them
    env.close(); twin.close()
    plain = bt.BatchedFixedWing(4, cfg=build_config())
    with pytest.raises(FwError):
        plain.get_field(bt.FIELD_PARAMS)
    plain.close()
