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
    def __init__(self, capacity, obs_dim=14, action_dim=3, device="cuda"):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        d = self.device
        self.observations = torch.zeros(self.capacity, obs_dim, dtype=torch.float32, device=d)
        self.next_observations = torch.zeros(self.capacity, obs_dim, dtype=torch.float32, device=d)
        self.actions = torch.zeros(self.capacity, action_dim, dtype=torch.float32, device=d)
        self.rewards = torch.zeros(self.capacity, dtype=torch.float32, device=d)
        self.dones = torch.zeros(self.capacity, dtype=torch.float32, device=d)
        self.pos = 0
        self.full = False

    def size(self):
        return self.capacity if self.full else self.pos

    def add(self, obs, next_obs, action, reward, done):
        """Batched insert of n transitions (one per env) at the ring head, wrapping around the end."""
        n = obs.shape[0]
        assert n <= self.capacity
        first = min(n, self.capacity - self.pos)
        for dst, src in ((self.observations, obs), (self.next_observations, next_obs), (self.actions, action),
                         (self.rewards, reward), (self.dones, done.to(torch.float32))):
            dst[self.pos:self.pos + first].copy_(src[:first])
            if first < n:
                dst[:n - first].copy_(src[first:])
        self.pos += n
        if self.pos >= self.capacity:
            self.full = True
            self.pos -= self.capacity

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, self.size(), (batch_size,), device=self.device, generator=generator)
        return (self.observations[idx], self.actions[idx], self.next_observations[idx], self.dones[idx],
                self.rewards[idx])


def _mlp(inp, out, hidden=(256, 256)):
    layers, d = [], inp
    for h in hidden:
        layers += [nn.Linear(d, h), nn.ReLU()]
        d = h
    layers.append(nn.Linear(d, out))
    return nn.Sequential(*layers)


class Actor(nn.Module):
    def __init__(self, obs_dim=14, action_dim=3, hidden=(256, 256)):
        super().__init__()
        self.net = _mlp(obs_dim, 2 * action_dim, hidden)
        self.action_dim = action_dim

    def forward(self, obs, deterministic=False):
        mean, log_std = self.net(obs).chunk(2, dim=-1)
        log_std = log_std.clamp(LOG_STD_MIN, LOG_STD_MAX)
        std = log_std.exp()
        u = mean if deterministic else mean + std * torch.randn_like(mean)
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
                 learning_rate=3e-4, gamma=0.99, tau=0.005, target_entropy="auto", normalize=True, seed=0):
        self.env, self.device, self.n_envs = env, env.device, env.num_envs
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
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), lr=learning_rate)
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=learning_rate)
        self.ent_opt = torch.optim.Adam([self.log_ent_coef], lr=learning_rate)
        self.buffer = ReplayBuffer(buffer_size, obs_dim=od, device=self.device)
        self.norm = DeviceVecNormalize(self.n_envs, obs_dim=od, device=self.device, gamma=gamma, norm_obs=normalize,
                                       norm_reward=normalize)
        self.num_timesteps = 0
        self._last_obs = None
        self.logs = []
        self.ep_ret_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        self.ep_count = torch.zeros((), dtype=torch.float64, device=self.device)
        self._run_ret = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)

    def _env_step(self):
        with torch.no_grad():
            if self.num_timesteps < self.learning_starts:
                actions = torch.rand(self.n_envs, 3, device=self.device) * 2 - 1      # uniform warm-up (sac.py learning_starts)
            else:
                actions, _ = self.actor(self._last_obs)
        obs_raw, rew_raw, done = self.env.step_tensor(actions.contiguous())
        d = done.bool()
        self._run_ret.add_(rew_raw.to(torch.float64))
        self.ep_ret_sum.add_((self._run_ret * d).sum())
        self.ep_count.add_(d.sum())
        self._run_ret.masked_fill_(d, 0.0)
        obs, rew = self.norm.step(obs_raw, rew_raw, done)
        self.buffer.add(self._last_obs, obs, actions, rew, done)
        self._last_obs = obs.clone()
        self.num_timesteps += self.n_envs

    def train_step(self):
        """One gradient step of sac.py:196-256 on a device-sampled batch."""
        obs, act, next_obs, done, rew = self.buffer.sample(self.batch_size)
        a_pi, logp = self.actor(obs)
        ent_coef = self.log_ent_coef.exp().detach()
        ent_loss = -(self.log_ent_coef * (logp + self.target_entropy).detach()).mean()
        self.ent_opt.zero_grad(set_to_none=True)
        ent_loss.backward()
        self.ent_opt.step()
        with torch.no_grad():
            na, nlogp = self.actor(next_obs)
            q_next = torch.min(*self.critic_target(next_obs, na)) - ent_coef * nlogp
            target = rew + (1 - done) * self.gamma * q_next
        q1, q2 = self.critic(obs, act)
        critic_loss = 0.5 * (F.mse_loss(q1, target) + F.mse_loss(q2, target))
        self.critic_opt.zero_grad(set_to_none=True)
        critic_loss.backward()
        self.critic_opt.step()
        q_pi = torch.min(*self.critic(obs, a_pi))
        actor_loss = (ent_coef * logp - q_pi).mean()
        self.actor_opt.zero_grad(set_to_none=True)
        actor_loss.backward()
        self.actor_opt.step()
        with torch.no_grad():
            for p, pt in zip(self.critic.parameters(), self.critic_target.parameters()):
                pt.mul_(1 - self.tau).add_(p, alpha=self.tau)
        return critic_loss.detach(), actor_loss.detach(), ent_coef

    def learn(self, total_timesteps, log_every=50, callback=None):
        if self._last_obs is None:
            self._last_obs = self.norm.reset(self.env.reset_tensor()).clone()
            self._t_start, self._it = time.time(), 0
        while self.num_timesteps < total_timesteps:
            self._env_step()
            stats = None
            if self.num_timesteps >= self.learning_starts and self.buffer.size() >= self.batch_size:
                for _ in range(self.gradient_steps):
                    stats = self.train_step()
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
