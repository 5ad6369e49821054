"""Multi-env SAC with the replay buffer in HBM (BASELINE.json config C5: 1024 envs, turbulence on).

The reference's SAC is single-env (`assert env.num_envs == 1`, stable_baselines3/common/off_policy_algorithm.py:387;
`assert n_envs == 1`, common/buffers.py:173).  This module keeps its algorithm and defaults
  SAC.train                      stable_baselines3/sac/sac.py:177-269   (twin critics, polyak 0.005, auto entropy
                                 coefficient with target entropy -|A|, lr 3e-4, gamma 0.99)
  ReplayBuffer.add / sample      stable_baselines3/common/buffers.py:146-256
  SACPolicy (squashed Gaussian)  stable_baselines3/sac/policies.py (net_arch [256, 256], ReLU, log_std in [-20, 2])
and widens the data path to N envs per step: a ring `[capacity, ...]` of device tensors, one batched insert of N
transitions per env step, uniform index sampling on the device.  Like the reference (SB3 0.10) the stored next
observation of a finished episode is the reset observation and `done` masks the bootstrap, time limits included.
"""
import math
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

from .buffers import DeviceVecNormalize

LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0


class ReplayBuffer:
    """The replay ring in HBM behind fw_replay_insert / fw_replay_sample (csrc/fw_replay.cu): one packed row per
    transition, ORIGINAL observations and rewards stored, normalisation with the current VecNormalize statistics at
    sample time (buffers.py:245-254), ring head / fill level / Philox sample counter on the device (CUDA-graph safe).
    On a CPU device (tests) the same layout and semantics run as tensor ops."""

    def __init__(self, capacity, obs_dim=14, action_dim=3, device="cuda", seed=0):
        self.capacity, self.obs_dim, self.action_dim = int(capacity), int(obs_dim), int(action_dim)
        self.device = torch.device(device)
        d = self.device
        self.row_floats = (2 * self.obs_dim + self.action_dim + 2 + 3) // 4 * 4
        self.rows = torch.zeros(self.capacity, self.row_floats, dtype=torch.float32, device=d)
        self.head_dev = torch.zeros((), dtype=torch.long, device=d)
        self.size_dev = torch.zeros((), dtype=torch.long, device=d)
        self.calls_dev = torch.zeros((), dtype=torch.long, device=d)
        self.seed = int(seed)
        self.pos, self.full = 0, False               # host mirror of the ring head (advance_host)
        self.last_indices = None
        self._c = None

    # views of the packed rows, reference names (buffers.py:176-184)
    @property
    def observations(self):
        return self.rows[:, :self.obs_dim]

    @property
    def next_observations(self):
        return self.rows[:, self.obs_dim:2 * self.obs_dim]

    @property
    def actions(self):
        return self.rows[:, 2 * self.obs_dim:2 * self.obs_dim + self.action_dim]

    @property
    def rewards(self):
        return self.rows[:, 2 * self.obs_dim + self.action_dim]

    @property
    def dones(self):
        return self.rows[:, 2 * self.obs_dim + self.action_dim + 1]

    def size(self):
        return self.capacity if self.full else self.pos

    def advance_host(self, n):
        self.pos += n
        if self.pos >= self.capacity:
            self.full = True
            self.pos -= self.capacity

    def _struct(self):
        if self._c is None:
            from . import _lib
            self._c = _lib.FwReplay(rows=self.rows.data_ptr(), head_dev=self.head_dev.data_ptr(),
                                    size_dev=self.size_dev.data_ptr(), sample_calls_dev=self.calls_dev.data_ptr(),
                                    capacity=self.capacity, obs_dim=self.obs_dim, act_dim=self.action_dim,
                                    row_floats=self.row_floats)
        return self._c

    def add(self, obs, next_obs, action, reward, done, advance_host=True):
        """Batched insert of n transitions (one per env) at the ring head, wrapping around the end.  advance_host=False
        when the call is being captured into / replayed from a CUDA graph: the caller then mirrors every replay with
        advance_host(n)."""
        n = obs.shape[0]
        assert n <= self.capacity
        if self.device.type == "cuda":
            import ctypes
            from . import _lib
            t = [obs.contiguous(), next_obs.contiguous(), action.contiguous(), reward.contiguous(),
                 done.to(torch.uint8).contiguous()]
            assert all(x.is_cuda for x in t) and all(x.dtype == torch.float32 for x in t[:4])
            self._keep = t
            _lib.check(_lib.lib().fw_replay_insert(ctypes.byref(self._struct()), *[ctypes.c_void_p(x.data_ptr()) for x in t],
                                                   n, ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       "fw_replay_insert")
        else:
            idx = (self.head_dev + torch.arange(n)) % self.capacity
            row = torch.zeros(n, self.row_floats)
            D, A = self.obs_dim, self.action_dim
            row[:, :D], row[:, D:2 * D], row[:, 2 * D:2 * D + A] = obs, next_obs, action
            row[:, 2 * D + A], row[:, 2 * D + A + 1] = reward, done.to(torch.float32)
            self.rows.index_copy_(0, idx, row)
            self.head_dev.add_(n).remainder_(self.capacity)
            self.size_dev.add_(n).clamp_(max=self.capacity)
        if advance_host:
            self.advance_host(n)

    def sample(self, batch_size, norm=None, indices=None):
        """(observations, actions, next_observations, dones, rewards) of batch_size uniformly drawn transitions;
        `norm` (DeviceVecNormalize) normalises observations and rewards with its CURRENT statistics.  `indices`
        (tests) replaces the draw."""
        D, A = self.obs_dim, self.action_dim
        if self.device.type == "cuda" and indices is None:
            import ctypes
            from . import _lib
            dev = self.device
            out = [torch.empty(batch_size, D, device=dev), torch.empty(batch_size, A, device=dev),
                   torch.empty(batch_size, D, device=dev), torch.empty(batch_size, device=dev),
                   torch.empty(batch_size, device=dev)]
            self.last_indices = torch.empty(batch_size, dtype=torch.long, device=dev)
            nm = _lib.FwReplayNorm()
            if norm is not None:
                nm.obs_mean, nm.obs_var = norm.obs_rms.mean.data_ptr(), norm.obs_rms.var.data_ptr()
                nm.ret_var = norm.ret_rms.var.data_ptr()
                nm.clip_obs, nm.clip_reward, nm.epsilon = norm.clip_obs, norm.clip_reward, norm.epsilon
                nm.norm_obs, nm.norm_reward = int(norm.norm_obs), int(norm.norm_reward)
            p = lambda x: ctypes.c_void_p(x.data_ptr())
            _lib.check(_lib.lib().fw_replay_sample(ctypes.byref(self._struct()), ctypes.byref(nm), batch_size, self.seed,
                                                   p(out[0]), p(out[1]), p(out[2]), p(out[3]), p(out[4]),
                                                   p(self.last_indices),
                                                   ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "fw_replay_sample")
            return tuple(out)
        if indices is None:
            indices = torch.randint(0, self.size(), (batch_size,), device=self.device)
        self.last_indices = indices
        r = self.rows[indices]
        obs, nxt, rew = r[:, :D], r[:, D:2 * D], r[:, 2 * D + A]
        if norm is not None:
            obs, nxt, rew = norm.normalize_obs(obs), norm.normalize_obs(nxt), norm.normalize_reward(rew)
        return obs, r[:, 2 * D:2 * D + A], nxt, r[:, 2 * D + A + 1], rew


def _mlp(inp, out, hidden=(256, 256)):
    from .ppo import SplitKLinear          # nn.Linear whose weight gradient is a split-K bmm for large batches
    layers, d = [], inp
    for h in hidden:
        layers += [SplitKLinear(d, h), nn.ReLU()]
        d = h
    layers.append(SplitKLinear(d, out))
    return nn.Sequential(*layers)


class Actor(nn.Module):
    def __init__(self, obs_dim=14, action_dim=3, hidden=(256, 256)):
        super().__init__()
        self.net = _mlp(obs_dim, 2 * action_dim, hidden)
        self.action_dim = action_dim

    def forward(self, obs, deterministic=False, eps=None):
        """`eps`: the unit normals of the reparameterised sample (tests inject the reference's draws)."""
        mean, log_std = self.net(obs).chunk(2, dim=-1)
        log_std = log_std.clamp(LOG_STD_MIN, LOG_STD_MAX)
        std = log_std.exp()
        u = mean if deterministic else mean + std * (torch.randn_like(mean) if eps is None else eps)
        a = torch.tanh(u)
        # log prob of the squashed Gaussian (SquashedDiagGaussianDistribution, common/distributions.py)
        logp = (-0.5 * ((u - mean) / std).pow(2) - log_std - 0.5 * math.log(2 * math.pi)).sum(-1)
        logp = logp - torch.log(1 - a.pow(2) + 1e-6).sum(-1)
        return a, logp


class Critic(nn.Module):
    def __init__(self, obs_dim=14, action_dim=3, hidden=(256, 256), n_critics=2):
        super().__init__()
        self.qs = nn.ModuleList([_mlp(obs_dim + action_dim, 1, hidden) for _ in range(n_critics)])

    def forward(self, obs, act):
        x = torch.cat([obs, act], dim=-1)
        return [q(x).squeeze(-1) for q in self.qs]


class SAC:
    def __init__(self, env, buffer_size=1_000_000, batch_size=4096, gradient_steps=2, learning_starts=10_000,
                 learning_rate=3e-4, gamma=0.99, tau=0.005, target_entropy="auto", normalize=True, seed=0,
                 use_cuda_graph=True, allow_eager_fallback=False):
        self.env, self.device, self.n_envs = env, env.device, env.num_envs
        self.allow_eager_fallback = bool(allow_eager_fallback)
        self.batch_size, self.gradient_steps, self.learning_starts = batch_size, gradient_steps, learning_starts
        self.gamma, self.tau = gamma, tau
        torch.manual_seed(seed)
        od = int(getattr(getattr(env, "sim", None), "obs_dim", 14))
        self.actor = Actor(obs_dim=od).to(self.device)
        self.critic = Critic(obs_dim=od).to(self.device)
        self.critic_target = Critic(obs_dim=od).to(self.device)
        self.critic_target.load_state_dict(self.critic.state_dict())
        self.log_ent_coef = torch.zeros(1, device=self.device, requires_grad=True)     # ent_coef "auto", init 1.0
        self.target_entropy = -3.0 if target_entropy == "auto" else float(target_entropy)
        self.use_cuda_graph = bool(use_cuda_graph) and self.device.type == "cuda"
        cap = self.use_cuda_graph
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), lr=learning_rate, capturable=cap)
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=learning_rate, capturable=cap)
        self.ent_opt = torch.optim.Adam([self.log_ent_coef], lr=learning_rate, capturable=cap)
        self._train_graph = None
        self._graph_stats = None
        self._env_graph = None
        self._env_warm = 0
        self.buffer = ReplayBuffer(buffer_size, obs_dim=od, device=self.device, seed=seed)
        self._last_obs_raw = None
        self.norm = DeviceVecNormalize(self.n_envs, obs_dim=od, device=self.device, gamma=gamma, norm_obs=normalize,
                                       norm_reward=normalize)
        self.num_timesteps = 0
        self._last_obs = None
        self.logs = []
        self.ep_ret_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        self.ep_count = torch.zeros((), dtype=torch.float64, device=self.device)
        self._run_ret = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)

    def _env_step_body(self, use_actor, capturable=False):
        with torch.no_grad():
            if use_actor:
                actions, _ = self.actor(self._last_obs)
            else:
                actions = torch.rand(self.n_envs, 3, device=self.device) * 2 - 1      # uniform warm-up (sac.py learning_starts)
        obs_raw, rew_raw, done = self.env.step_tensor(actions.contiguous())
        d = done.bool()
        self._run_ret.add_(rew_raw.to(torch.float64))
        self.ep_ret_sum.add_((self._run_ret * d).sum())
        self.ep_count.add_(d.sum())
        self._run_ret.masked_fill_(d, 0.0)
        obs, _ = self.norm.step(obs_raw, rew_raw, done)
        # the ring keeps the ORIGINAL observation and reward (off_policy_algorithm.py:430-436); they are normalised with
        # the statistics current when a batch is drawn (buffers.py:245-254)
        self.buffer.add(self._last_obs_raw, obs_raw, actions, rew_raw, done, advance_host=not capturable)
        self._last_obs_raw.copy_(obs_raw)
        self._last_obs.copy_(obs)

    def _env_step(self):
        """One env step of all envs + replay insert.  Once the actor drives the envs the step (actor forward, the
        simulator kernels, normaliser, ring insert: ~100 launches) is replayed as one CUDA graph."""
        use_actor = self.num_timesteps >= self.learning_starts
        if use_actor and self.use_cuda_graph and self._env_graph is not False:
            if self._env_graph is None:
                if self._env_warm < 2:                              # two eager steps first (allocator, lazy init)
                    self._env_warm += 1
                    self._env_step_body(True, capturable=True)
                    self.buffer.advance_host(self.n_envs)
                    self.num_timesteps += self.n_envs
                    return
                self.env.sim.join()
                torch.cuda.synchronize(self.device)
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._env_step_body(True, capturable=True)
                    self._env_graph = g
                except Exception as e:
                    if not getattr(self, "allow_eager_fallback", False):
                        raise RuntimeError("SAC: CUDA-graph capture of the env step failed (set allow_eager_fallback or "
                                           "use_cuda_graph=False to run eagerly)") from e
                    self.logs.append({"env_cuda_graph_disabled": repr(e)})
                    self._env_graph = False
                    torch.cuda.synchronize(self.device)
            if self._env_graph:
                self._env_graph.replay()
                self.buffer.advance_host(self.n_envs)
                self.num_timesteps += self.n_envs
                return
        self._env_step_body(use_actor)
        self.num_timesteps += self.n_envs

    def train_step_graphed(self):
        """train_step replayed as one CUDA graph: the eager step is ~150 small launches driven from python (4.5 ms);
        captured it is bound by the kernels themselves.  The first calls run eagerly (optimizer state, allocator)."""
        if not self.use_cuda_graph:
            return self.train_step()
        if self._train_graph is None:
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self.train_step(capturable=True)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    stats = self.train_step(capturable=True)
                    self._graph_stats = tuple(x.clone() for x in stats)
                self._train_graph = g
            except Exception as e:
                if not getattr(self, "allow_eager_fallback", False):
                    raise RuntimeError("SAC: CUDA-graph capture of the gradient step failed (set allow_eager_fallback or "
                                       "use_cuda_graph=False to run eagerly)") from e
                self.logs.append({"cuda_graph_disabled": repr(e)})
                self.use_cuda_graph = False
                torch.cuda.synchronize(self.device)
                return self.train_step()
        self._train_graph.replay()
        return self._graph_stats

    def train_step(self, capturable=False, batch=None, noise=(None, None)):
        """One gradient step of sac.py:196-256 on a device-sampled batch.  `batch` / `noise` (tests): a given
        (obs, actions, next_obs, dones, rewards) tuple and the unit normals of the two policy samples."""
        obs, act, next_obs, done, rew = batch if batch is not None else self.buffer.sample(self.batch_size, norm=self.norm)
        a_pi, logp = self.actor(obs, eps=noise[0])
        ent_coef = self.log_ent_coef.exp().detach()
        ent_loss = -(self.log_ent_coef * (logp + self.target_entropy).detach()).mean()
        self.ent_opt.zero_grad(set_to_none=not capturable)
        ent_loss.backward()
        self.ent_opt.step()
        with torch.no_grad():
            na, nlogp = self.actor(next_obs, eps=noise[1])
            q_next = torch.min(*self.critic_target(next_obs, na)) - ent_coef * nlogp
            target = rew + (1 - done) * self.gamma * q_next
        q1, q2 = self.critic(obs, act)
        critic_loss = 0.5 * (F.mse_loss(q1, target) + F.mse_loss(q2, target))
        self.critic_opt.zero_grad(set_to_none=not capturable)
        critic_loss.backward()
        self.critic_opt.step()
        q_pi = torch.min(*self.critic(obs, a_pi))
        actor_loss = (ent_coef * logp - q_pi).mean()
        self.actor_opt.zero_grad(set_to_none=not capturable)
        actor_loss.backward()
        self.actor_opt.step()
        with torch.no_grad():                                   # polyak update, two launches for all tensors
            src, dst = list(self.critic.parameters()), list(self.critic_target.parameters())
            torch._foreach_mul_(dst, 1 - self.tau)
            torch._foreach_add_(dst, src, alpha=self.tau)
        return critic_loss.detach(), actor_loss.detach(), ent_coef.reshape(())

    def learn(self, total_timesteps, log_every=50, callback=None):
        if self._last_obs is None:
            raw = self.env.reset_tensor()
            self._last_obs_raw = raw.clone()
            self._last_obs = self.norm.reset(raw).clone()
            self._t_start, self._it = time.time(), 0
        while self.num_timesteps < total_timesteps:
            self._env_step()
            stats = None
            if self.num_timesteps >= self.learning_starts and self.buffer.size() >= self.batch_size:
                for _ in range(self.gradient_steps):
                    stats = self.train_step_graphed()
            self._it += 1
            if self._it % log_every == 0:
                r, c = self.ep_ret_sum.item(), self.ep_count.item()
                self.ep_ret_sum.zero_(); self.ep_count.zero_()
                row = {"iteration": self._it, "timesteps": self.num_timesteps,
                       "fps": self.num_timesteps / (time.time() - self._t_start),
                       "ep_rew_mean": r / c if c else float("nan"), "episodes": int(c)}
                if stats is not None:
                    row.update(critic_loss=float(stats[0]), actor_loss=float(stats[1]), ent_coef=float(stats[2]))
                self.logs.append(row)
                if callback is not None:
                    callback(row)
        return self
