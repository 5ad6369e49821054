"""CPU tests of the host-side logic, including the N>1 path under gloo (world_size 2): gradient all-reduce,
normaliser-moment merge, env sharding by global env id, rollout-buffer layout."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tum_adlr_deep_reinforcement_learning_b200.buffers import RolloutBuffer, RunningMeanStd
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
from tum_adlr_deep_reinforcement_learning_b200.ppo import ActorCritic, allreduce_gradients


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = ActorCritic()
    rs = np.random.RandomState(100 + rank)
    obs = torch.as_tensor(rs.standard_normal((64, 14)), dtype=torch.float32)
    a, v, lp = net(obs)
    (lp.mean() + v.mean()).backward()
    local = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None]).clone()
    flat = allreduce_gradients(list(net.parameters()), dist, world)
    rms = RunningMeanStd((14,), device="cpu")
    data = rs.standard_normal((200 + 50 * rank, 14)) * (1 + rank) + rank
    rms.update(torch.as_tensor(data))
    rms.sync(dist)
    q.put((rank, local.numpy(), flat.numpy(), rms.mean.numpy(), rms.var.numpy(), float(rms.count), data))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_and_moment_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, f0, m0, v0, c0, d0), (_, l1, f1, m1, v1, c1, d1) = res
    assert np.allclose(f0, (l0 + l1) / 2, atol=1e-7) and np.array_equal(f0, f1)
    both = np.concatenate([d0, d1])
    assert np.allclose(m0, both.mean(0), atol=1e-4) and np.allclose(v0, both.var(0), rtol=1e-3)
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)
    assert abs(c0 - (len(both) + 2e-4) / 2) < 1e-9


def test_running_mean_std_matches_reference_formula():
    rs = np.random.RandomState(0)
    rms = RunningMeanStd((3,), device="cpu")
    chunks = [rs.standard_normal((n, 3)) * 3 + 1 for n in (5, 17, 64)]
    for c in chunks:
        rms.update(torch.as_tensor(c))
    # running_mean_std.py:19-39 evaluated with numpy
    mean, var, count = np.zeros(3), np.ones(3), 1e-4
    for c in chunks:
        bm, bv, bc = c.mean(0), c.var(0), c.shape[0]
        delta = bm - mean
        tot = count + bc
        new_mean = mean + delta * bc / tot
        m2 = var * count + bv * bc + np.square(delta) * count * bc / tot
        mean, var, count = new_mean, m2 / tot, tot
    assert np.allclose(rms.mean.numpy(), mean, atol=1e-12) and np.allclose(rms.var.numpy(), var, atol=1e-12)


def test_rollout_buffer_flatten_is_env_major_like_swap_and_flatten():
    T, N = 4, 3
    buf = RolloutBuffer(T, N, device="cpu")
    for t in range(T):
        obs = torch.arange(N, dtype=torch.float32).reshape(N, 1).repeat(1, 14) + 10 * t
        buf.add(obs, torch.zeros(N, 3), torch.full((N,), float(t)), torch.zeros(N), torch.zeros(N), torch.zeros(N))
    assert buf.full
    flat = buf.flat(buf.observations)[:, 0].numpy()
    # swap_and_flatten (buffers.py:51-64): index = env * T + t
    expect = np.array([n + 10 * t for n in range(N) for t in range(T)], dtype=np.float32)
    assert np.array_equal(flat, expect)
    g = torch.Generator().manual_seed(0)
    seen = torch.cat([b.observations[:, 0] for b in buf.get(5, generator=g)])
    assert sorted(seen.tolist()) == sorted(expect.tolist())


def test_env_sharding_is_independent_of_rank_count():
    """Envs are keyed by GLOBAL env id in the Philox counter: two shards of 8 envs (offsets 0, 8) reset to exactly
    the same states as one batch of 16 — the property that makes N-GPU results independent of N."""
    from oracle import fw_oracle as O
    cfg = build_config(seed=42)
    whole = O.OracleBatch(cfg, 16).reset().copy()
    parts = []
    for off in (0, 8):
        c = build_config(seed=42, env_id_offset=off)
        parts.append(O.OracleBatch(c, 8).reset().copy())
    assert np.array_equal(whole, np.concatenate(parts))
    assert len({tuple(r) for r in whole}) == 16          # every env has its own stream


def test_replay_buffer_ring_semantics():
    """Multi-env ring insert (the widening of common/buffers.py:146-256 to N envs per step): wrap-around, size,
    uniform sampling only from filled slots."""
    from tum_adlr_deep_reinforcement_learning_b200.sac import ReplayBuffer
    buf = ReplayBuffer(10, obs_dim=2, action_dim=1, device="cpu")
    for k in range(4):                                   # 4 inserts of 3 transitions into capacity 10
        base = 3 * k
        obs = torch.arange(base, base + 3, dtype=torch.float32).reshape(3, 1).repeat(1, 2)
        buf.add(obs, obs + 100, torch.zeros(3, 1), obs[:, 0], torch.tensor([0, 0, 1]))
        assert buf.size() == min(10, base + 3)
    assert buf.full and buf.pos == 2
    # slots 0,1 were overwritten by transitions 10, 11; the rest still hold 2..9
    assert buf.observations[:, 0].tolist() == [10, 11, 2, 3, 4, 5, 6, 7, 8, 9]
    assert buf.next_observations[0, 0].item() == 110 and buf.dones[1].item() == 1.0
    torch.manual_seed(0)
    o, a, no, d, r = buf.sample(256)
    assert o.shape == (256, 2) and set(o[:, 0].tolist()) <= set(range(2, 12))
    assert torch.equal(no, o + 100) and torch.equal(r, o[:, 0])
    small = ReplayBuffer(100, obs_dim=2, action_dim=1, device="cpu")
    small.add(torch.ones(3, 2), torch.ones(3, 2), torch.zeros(3, 1), torch.ones(3), torch.zeros(3))
    assert small.size() == 3 and small.sample(64)[0].eq(1).all()


def test_split_k_linear_gradients_match_nn_linear():
    """ppo._LinearSplitK (weight gradient as a chunked bmm + sum, bias gradient in two stages) against
    torch.nn.functional.linear on CPU tensors, for batch sizes with and without a power-of-two chunking."""
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.ppo import _LinearSplitK
    torch.manual_seed(0)
    for B, din, dout in ((8192, 14, 64), (4096, 64, 3), (1000, 64, 1)):
        x = torch.randn(B, din, dtype=torch.float64, requires_grad=True)
        w = torch.randn(dout, din, dtype=torch.float64, requires_grad=True)
        b = torch.randn(dout, dtype=torch.float64, requires_grad=True)
        up = torch.randn(B, dout, dtype=torch.float64)
        (_LinearSplitK.apply(x, w, b) * up).sum().backward()
        got = [t.grad.clone() for t in (x, w, b)]
        for t in (x, w, b):
            t.grad = None
        (torch.nn.functional.linear(x, w, b) * up).sum().backward()
        for g, t in zip(got, (x, w, b)):
            assert torch.allclose(g, t.grad, rtol=1e-12, atol=1e-10)


def test_episode_end_info_builder_layout():
    """vec_env._done_info: the compiled dict display carries every reference key with the METRIC_LAYOUT offsets."""
    import numpy as np
    from tum_adlr_deep_reinforcement_learning_b200 import vec_env as V
    from tum_adlr_deep_reinforcement_learning_b200.config import METRIC_LAYOUT
    row = [float(i) for i in range(28)] + [-12.5, 77.0, 11.0]            # metrics | return | length | term code
    tob = np.arange(14, dtype=np.float32)
    info = V._done_info(row, tob, 3.25)
    assert info["termination"] == "omega_q" or isinstance(info["termination"], (str, int))
    assert info["episode"] == {"r": -12.5, "l": 77, "t": 3.25} and info["terminal_observation"] is tob
    for name, off, keys in METRIC_LAYOUT:
        assert list(info[name]) == list(keys)
        for q, k in enumerate(keys):
            assert info[name][k] == (bool(row[off + q]) if name == "success" else row[off + q])


def _vecnorm_fixture():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vecnorm.npz"))


def test_vecnormalize_and_rollout_buffer_match_the_reference():
    """buffers.DeviceVecNormalize / RunningMeanStd / RolloutBuffer (CPU tensors) against the live reference's
    VecNormalize + RolloutBuffer on a scripted env (tests/golden/vecnorm.npz): reset semantics (the reset feeds zeros
    to ret_rms and leaves obs_rms alone), normalised observations / rewards, running moments, buffer rows (previous
    observation and done flags are stored), swap_and_flatten order.  The reference takes the batch moments of float32
    observations in float32, hence the 1e-5 tolerances."""
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.buffers import DeviceVecNormalize
    g = _vecnorm_fixture()
    T, N, D = g["rew_seq"].shape[0], g["rew_seq"].shape[1], g["obs_seq"].shape[2]
    nm = DeviceVecNormalize(N, obs_dim=D, device="cpu", gamma=0.99)
    buf = RolloutBuffer(T, N, obs_dim=D, device="cpu")
    last_obs = nm.reset(torch.as_tensor(g["obs_seq"][0]))
    assert np.allclose(last_obs.numpy(), g["norm_obs"][0], rtol=1e-6, atol=1e-6)
    last_dones = torch.zeros(N)
    for t in range(T):
        o, r = nm.step(torch.as_tensor(g["obs_seq"][t + 1]), torch.as_tensor(g["rew_seq"][t]),
                       torch.as_tensor(g["done_seq"][t]))
        buf.add(last_obs, torch.as_tensor(g["acts"][t]), r, last_dones, torch.as_tensor(g["vals"][t]),
                torch.as_tensor(g["logp"][t]))
        last_obs, last_dones = o, torch.as_tensor(g["done_seq"][t]).float()
        assert np.allclose(o.numpy(), g["norm_obs"][t + 1], rtol=2e-5, atol=2e-5), t
        assert np.allclose(r.numpy(), g["norm_rew"][t], rtol=2e-5, atol=1e-6), t
    assert np.allclose(nm.obs_rms.mean.numpy(), g["obs_mean"], rtol=1e-5, atol=1e-5)
    assert np.allclose(nm.obs_rms.var.numpy(), g["obs_var"], rtol=1e-5)
    assert np.isclose(float(nm.obs_rms.count), float(g["obs_count"])) and np.isclose(float(nm.ret_rms.count), float(g["ret_count"]))
    assert np.isclose(float(nm.ret_rms.mean), float(g["ret_mean"]), rtol=1e-6) and np.isclose(float(nm.ret_rms.var), float(g["ret_var"]), rtol=1e-6)
    assert np.allclose(nm.ret.numpy(), g["ret"], rtol=1e-6, atol=1e-6)
    for mine, ref in ((buf.observations, "buf_obs"), (buf.actions, "buf_act"), (buf.rewards, "buf_rew"),
                      (buf.dones, "buf_done"), (buf.values, "buf_val"), (buf.log_probs, "buf_logp")):
        assert np.allclose(mine.numpy(), g[ref], rtol=2e-5, atol=2e-5), ref
    assert np.allclose(buf.flat(buf.observations).numpy(), g["flat_obs"], rtol=2e-5, atol=2e-5)


def _load_ppo_fixture_into(policy, g, tag):
    import torch
    with torch.no_grad():
        for name, p in policy.named_parameters():
            p.copy_(torch.as_tensor(g["%s/%s" % (tag, name)]).to(p.device))


def _ppo_fixture_batch(g, device):
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.buffers import RolloutBufferSamples
    t = lambda k: torch.as_tensor(g[k]).to(device)
    return RolloutBufferSamples(t("obs"), t("act"), t("old_values"), t("old_log_prob"), t("adv"), t("ret"))


def test_ppo_policy_and_update_match_the_reference_on_cpu():
    """ppo.ActorCritic.evaluate_actions and three PPO minibatch updates (tensor-op loss, clip_grad_norm_, torch Adam)
    against the live reference's ActorCriticPolicy / PPO.train() (tests/golden/ppo_update.npz)."""
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.ppo import ActorCritic
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo_update.npz"))
    pol = ActorCritic()
    _load_ppo_fixture_into(pol, g, "w0")
    batch = _ppo_fixture_batch(g, "cpu")
    with torch.no_grad():
        v, lp, ent = pol.evaluate_actions(batch.observations, batch.actions)
    assert np.allclose(v.numpy(), g["eval_values"], rtol=1e-5, atol=1e-5)
    assert np.allclose(lp.numpy(), g["eval_log_prob"], rtol=1e-5, atol=1e-4)
    assert np.allclose(ent.numpy(), g["eval_entropy"], rtol=1e-6)
    opt = torch.optim.Adam(pol.parameters(), lr=3e-4, eps=1e-5)
    for k in (1, 2, 3):
        values, log_prob, entropy = pol.evaluate_actions(batch.observations, batch.actions)
        adv = (batch.advantages - batch.advantages.mean()) / (batch.advantages.std() + 1e-8)
        ratio = torch.exp(log_prob - batch.old_log_prob)
        policy_loss = -torch.min(adv * ratio, adv * torch.clamp(ratio, 0.8, 1.2)).mean()
        loss = policy_loss + 0.01 * (-entropy.mean()) + 0.5 * torch.nn.functional.mse_loss(batch.returns, values)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(pol.parameters(), 0.5)
        opt.step()
        for name, p in pol.named_parameters():
            ref = g["w%d/%s" % (k, name)]
            assert np.abs(p.detach().numpy() - ref).max() <= 2e-6 + 1e-5 * np.abs(ref - g["w0/" + name]).max(), (k, name)


def test_vecnormalize_checkpoint_round_trips_with_the_reference_format(tmp_path):
    """DeviceVecNormalize.save writes what the fork's VecNormalize.load reads (vec_normalize.py:222-243) and
    DeviceVecNormalize.load reads what the fork's VecNormalize.save writes: statistics and settings survive both ways."""
    import torch
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference libraries neither mounted nor mirrored under baseline/_ref")
    refshim.install()
    import gym
    from stable_baselines3.common.vec_env import DummyVecEnv, VecNormalize
    from tum_adlr_deep_reinforcement_learning_b200.buffers import DeviceVecNormalize

    class Toy(gym.Env):
        observation_space = gym.spaces.Box(low=-np.ones(5, np.float32), high=np.ones(5, np.float32))
        action_space = gym.spaces.Box(low=-np.ones(2, np.float32), high=np.ones(2, np.float32))

        def reset(self):
            return np.zeros(5, np.float32)

        def step(self, a):
            return np.zeros(5, np.float32), 0.0, False, {}

    rs = np.random.RandomState(0)
    mine = DeviceVecNormalize(3, obs_dim=5, device="cpu", clip_obs=7.0, gamma=0.97)
    mine.reset(torch.zeros(3, 5))
    for _ in range(20):
        mine.step(torch.as_tensor(rs.standard_normal((3, 5)), dtype=torch.float32),
                  torch.as_tensor(rs.standard_normal(3), dtype=torch.float32), torch.zeros(3, dtype=torch.uint8))
    mine.save(str(tmp_path / "mine.pkl"))
    ref = VecNormalize.load(str(tmp_path / "mine.pkl"), DummyVecEnv([Toy for _ in range(3)]))
    assert np.array_equal(ref.obs_rms.mean, mine.obs_rms.mean.numpy()) and np.array_equal(ref.obs_rms.var, mine.obs_rms.var.numpy())
    assert ref.obs_rms.count == float(mine.obs_rms.count) and ref.ret_rms.var == float(mine.ret_rms.var)
    assert ref.clip_obs == 7.0 and ref.gamma == 0.97 and ref.training and ref.num_envs == 3
    o = rs.standard_normal((3, 5)).astype(np.float32)
    assert np.allclose(ref.normalize_obs(o), mine.normalize_obs(torch.as_tensor(o)).numpy(), atol=1e-6)
    # the other way round
    ref2 = VecNormalize(DummyVecEnv([Toy for _ in range(3)]), clip_reward=4.0)
    ref2.obs_rms.update(rs.standard_normal((50, 5)))
    ref2.ret_rms.update(rs.standard_normal(50))
    ref2.save(str(tmp_path / "ref.pkl"))
    back = DeviceVecNormalize(3, obs_dim=5, device="cpu").load(str(tmp_path / "ref.pkl"))
    assert np.array_equal(back.obs_rms.mean.numpy(), ref2.obs_rms.mean) and float(back.ret_rms.var) == float(ref2.ret_rms.var)
    assert float(back.obs_rms.count) == ref2.obs_rms.count and back.clip_reward == 4.0


def _sac_from_fixture(g, tag, device):
    """sac.SAC (2 x 64 networks) with the reference SACPolicy weights of snapshot `tag` (tests/golden/sac_update.npz)."""
    import types
    import torch
    from tum_adlr_deep_reinforcement_learning_b200 import sac as S

    env = types.SimpleNamespace(device=torch.device(device), num_envs=1, sim=types.SimpleNamespace(obs_dim=14))
    real_mlp = S._mlp
    S._mlp = lambda inp, out, hidden=(64, 64): real_mlp(inp, out, (64, 64))
    try:
        algo = S.SAC(env, buffer_size=256, batch_size=64, normalize=False, use_cuda_graph=False)
    finally:
        S._mlp = real_mlp
    _sac_load(algo, g, tag)
    return algo


def _sac_ref_weights(g, tag):
    """name -> array for (actor.net, critic.qs, critic_target.qs, log_ent_coef) in sac.py's layout: the reference keeps
    mu / log_std as two Linear heads (sac/policies.py:95-107), sac.Actor one Linear whose output is chunked in two."""
    t = lambda k: g["%s/%s" % (tag, k)]
    out = {}
    for i, j in ((0, 0), (2, 2)):
        out["actor.net.%d.weight" % j], out["actor.net.%d.bias" % j] = t("actor.latent_pi.%d.weight" % i), t("actor.latent_pi.%d.bias" % i)
    out["actor.net.4.weight"] = np.concatenate([t("actor.mu.weight"), t("actor.log_std.weight")])
    out["actor.net.4.bias"] = np.concatenate([t("actor.mu.bias"), t("actor.log_std.bias")])
    for net in ("critic", "critic_target"):
        for q in (0, 1):
            for layer in (0, 2, 4):
                for leaf in ("weight", "bias"):
                    out["%s.qs.%d.%d.%s" % (net, q, layer, leaf)] = t("%s.qf%d.%d.%s" % (net, q, layer, leaf))
    out["log_ent_coef"] = t("log_ent_coef")
    return out


def _sac_named(algo):
    import itertools
    return dict(itertools.chain((("actor." + k, v) for k, v in algo.actor.named_parameters()),
                                (("critic." + k, v) for k, v in algo.critic.named_parameters()),
                                (("critic_target." + k, v) for k, v in algo.critic_target.named_parameters()),
                                [("log_ent_coef", algo.log_ent_coef)]))


def _sac_load(algo, g, tag):
    import torch
    ref = _sac_ref_weights(g, tag)
    with torch.no_grad():
        for k, p in _sac_named(algo).items():
            p.copy_(torch.as_tensor(ref[k]).reshape(p.shape))


def _sac_check_steps(algo, g, device, tol_scale=1.0):
    import torch
    dev = torch.device(device)
    rows = {k: torch.as_tensor(g["rb_" + k]).to(dev) for k in ("obs", "next_obs", "act", "rew", "done")}
    w0 = _sac_ref_weights(g, "w0")
    for k in (1, 2, 3):
        idx = torch.as_tensor(g["idx_%d" % k]).to(dev)
        batch = (rows["obs"][idx], rows["act"][idx], rows["next_obs"][idx], rows["done"][idx], rows["rew"][idx])
        noise = (torch.as_tensor(g["eps_pi_%d" % k]).to(dev), torch.as_tensor(g["eps_next_%d" % k]).to(dev))
        if k == 1:
            with torch.no_grad():
                a_pi, logp = algo.actor(batch[0], eps=noise[0])
            assert np.allclose(a_pi.cpu().numpy(), g["actions_pi_1"], atol=2e-6)
            assert np.allclose(logp.cpu().numpy(), g["log_prob_1"].reshape(-1), atol=2e-5)
        algo.train_step(batch=batch, noise=noise)
        ref = _sac_ref_weights(g, "w%d" % k)
        for name, p in _sac_named(algo).items():
            got = p.detach().cpu().numpy().reshape(ref[name].shape)
            step = np.abs(ref[name] - w0[name]).max()
            assert np.abs(got - ref[name]).max() <= tol_scale * (2e-6 + 2e-3 * step), (k, name, np.abs(got - ref[name]).max(), step)


def test_sac_update_matches_the_reference_on_cpu():
    """sac.SAC.train_step against three gradient steps of the live reference's SAC.train() (sac/sac.py:177-269;
    tests/golden/sac_update.npz: same sampled rows, same unit normals): actor, both critics, the polyak-averaged targets and
    the entropy coefficient after every step."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sac_update.npz"))
    _sac_check_steps(_sac_from_fixture(g, "w0", "cpu"), g, "cpu")


def test_replay_ring_semantics_on_cpu():
    """ReplayBuffer (tensor path): packed rows, wrap-around insert, and sample-time normalisation equal to the live
    reference's ReplayBuffer._get_samples under a VecNormalize env (buffers.py:245-254)."""
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.buffers import DeviceVecNormalize
    from tum_adlr_deep_reinforcement_learning_b200.sac import ReplayBuffer
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sac_update.npz"))
    rb = ReplayBuffer(150, device="cpu")
    t = lambda k: torch.as_tensor(g["rb_" + k])
    for lo in range(0, 150, 50):
        rb.add(t("obs")[lo:lo + 50], t("next_obs")[lo:lo + 50], t("act")[lo:lo + 50], t("rew")[lo:lo + 50], t("done")[lo:lo + 50])
    assert rb.full and rb.pos == 0 and int(rb.size_dev) == 150 and int(rb.head_dev) == 0
    assert torch.equal(rb.observations, t("obs")) and torch.equal(rb.dones, t("done"))
    norm = DeviceVecNormalize(1, device="cpu", clip_obs=5.0, clip_reward=3.0)
    norm.load_state_dict({"obs_mean": g["norm_obs_mean"], "obs_var": g["norm_obs_var"], "obs_count": 1.0,
                          "ret_mean": 0.0, "ret_var": g["norm_ret_var"], "ret_count": 1.0})
    obs, act, nxt, done, rew = rb.sample(40, norm=norm, indices=torch.as_tensor(g["norm_idx"]))
    assert np.allclose(obs.numpy(), g["norm_obs"], atol=1e-6) and np.allclose(nxt.numpy(), g["norm_next_obs"], atol=1e-6)
    assert np.allclose(rew.numpy(), g["norm_rew"], atol=1e-6) and np.array_equal(done.numpy(), g["norm_done"])
    assert np.array_equal(act.numpy(), g["norm_act"])
    rb.add(t("obs")[:7] + 1, t("next_obs")[:7], t("act")[:7], t("rew")[:7], t("done")[:7])      # wraps over the oldest rows
    assert rb.pos == 7 and torch.equal(rb.observations[:7], t("obs")[:7] + 1) and torch.equal(rb.observations[7:], t("obs")[7:])
