"""On-device rollout storage + GAE for PPO, and the running-moment normaliser of VecNormalize.

Mirrors the reference's forked SB3:
  RolloutBuffer.reset / add / compute_returns_and_advantage / get   (stable_baselines3/common/buffers.py:259-389)
  RunningMeanStd.update (Chan parallel update)                       (common/running_mean_std.py:19-39)
  VecNormalize.step_wait / normalize_obs / normalize_reward          (common/vec_env/vec_normalize.py:106-178)
Everything is a CUDA tensor: time-major [T, N, ...] float32 arrays that never leave HBM; GAE runs in the
hand-written fw_gae kernel (bit-exact with the reference's mixed f32/f64 arithmetic); minibatches are
device-side gathers of a device-side permutation instead of host numpy slices + H2D copies.
"""
from collections import namedtuple

import torch

from . import batched as bt

RolloutBufferSamples = namedtuple("RolloutBufferSamples", ["observations", "actions", "old_values", "old_log_prob",
                                                           "advantages", "returns"])


class RolloutBuffer:
    def __init__(self, buffer_size, n_envs, obs_dim=14, action_dim=3, device="cuda", gae_lambda=0.95, gamma=0.99):
        self.buffer_size, self.n_envs = int(buffer_size), int(n_envs)
        self.obs_dim, self.action_dim = obs_dim, action_dim
        self.device = torch.device(device)
        self.gae_lambda, self.gamma = gae_lambda, gamma
        T, N, d = self.buffer_size, self.n_envs, self.device
        self.observations = torch.zeros(T, N, obs_dim, dtype=torch.float32, device=d)
        self.actions = torch.zeros(T, N, action_dim, dtype=torch.float32, device=d)
        self.rewards = torch.zeros(T, N, dtype=torch.float32, device=d)
        self.dones = torch.zeros(T, N, dtype=torch.float32, device=d)
        self.values = torch.zeros(T, N, dtype=torch.float32, device=d)
        self.log_probs = torch.zeros(T, N, dtype=torch.float32, device=d)
        self.advantages = torch.zeros(T, N, dtype=torch.float32, device=d)
        self.returns = torch.zeros(T, N, dtype=torch.float32, device=d)
        self.pos = 0
        self.full = False

    def reset(self):
        self.pos = 0
        self.full = False

    def add(self, obs, action, reward, done, value, log_prob):
        """`done` is the "episode ended before obs" flag of the PREVIOUS step, exactly like the reference stores
        `self._last_dones` (on_policy_algorithm.py:178-180)."""
        t = self.pos
        self.observations[t].copy_(obs)
        self.actions[t].copy_(action)
        self.rewards[t].copy_(reward)
        self.dones[t].copy_(done)
        self.values[t].copy_(value.reshape(-1))
        self.log_probs[t].copy_(log_prob.reshape(-1))
        self.pos += 1
        if self.pos == self.buffer_size:
            self.full = True

    def compute_returns_and_advantage(self, last_values, dones):
        """GAE(lambda) without time-limit bootstrapping, as buffers.py:304-333.  `dones`: the done vector returned by
        the final env.step (bool / uint8)."""
        adv, ret = bt.gae(self.rewards, self.values, self.dones, last_values.reshape(-1).to(torch.float32),
                          dones.reshape(-1).to(torch.uint8), self.gamma, self.gae_lambda)
        self.advantages.copy_(adv)
        self.returns.copy_(ret)

    def flat(self, x):
        """swap_and_flatten (buffers.py:51-64): [T, N, ...] -> [N*T, ...] in env-major order."""
        shape = x.shape
        return x.transpose(0, 1).reshape(shape[0] * shape[1], *shape[2:])

    def get(self, batch_size=None, generator=None):
        assert self.full, "rollout buffer is not full"
        total = self.buffer_size * self.n_envs
        perm = torch.randperm(total, device=self.device, generator=generator)
        flat = [self.flat(x) for x in (self.observations, self.actions, self.values, self.log_probs, self.advantages,
                                       self.returns)]
        if batch_size is None:
            batch_size = total
        for start in range(0, total, batch_size):
            idx = perm[start:start + batch_size]
            yield RolloutBufferSamples(*(f.index_select(0, idx) for f in flat))


class RunningMeanStd:
    """running_mean_std.py:6-39 on the device, float64 statistics.  `sync(group)` all-reduces the moments so that every
    data-parallel rank normalises identically."""

    def __init__(self, shape=(), device="cuda", epsilon=1e-4):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = torch.tensor(float(epsilon), dtype=torch.float64, device=device)

    def update(self, x):
        x = x.to(torch.float64)
        self.update_from_moments(x.mean(dim=0), x.var(dim=0, unbiased=False), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m_2 = self.var * self.count + batch_var * batch_count + delta.square() * self.count * batch_count / tot
        # in place: the tensors keep their addresses, so the update can live inside a captured CUDA graph
        self.mean.copy_(new_mean)
        self.var.copy_(m_2 / tot)
        self.count.copy_(tot)

    def sync(self, dist):
        """Merge the per-rank moments with one all-reduce of (count, count*mean, count*(var + mean^2))."""
        packed = torch.cat([self.count.reshape(1), (self.count * self.mean).reshape(-1),
                            (self.count * (self.var + self.mean.square())).reshape(-1)])
        dist.all_reduce(packed)
        k = self.mean.numel()
        cnt = packed[0]
        mean = (packed[1:1 + k] / cnt).reshape(self.mean.shape)
        ex2 = (packed[1 + k:] / cnt).reshape(self.mean.shape)
        world = dist.get_world_size()
        self.mean.copy_(mean)
        self.var.copy_((ex2 - mean.square()).clamp_min(0))
        self.count.copy_(cnt / world)


class DeviceVecNormalize:
    """VecNormalize (vec_normalize.py:13-243) over device tensors: running obs / return statistics, clipped
    normalisation, `ret = ret * gamma + r`, `ret[done] = 0`."""

    def __init__(self, n_envs, obs_dim=14, device="cuda", norm_obs=True, norm_reward=True, clip_obs=10.0,
                 clip_reward=10.0, gamma=0.99, epsilon=1e-8, training=True):
        self.obs_rms = RunningMeanStd((obs_dim,), device)
        self.ret_rms = RunningMeanStd((), device)
        self.ret = torch.zeros(n_envs, dtype=torch.float64, device=device)
        self.norm_obs, self.norm_reward = norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.training = training

    def normalize_obs(self, obs):
        if not self.norm_obs:
            return obs
        o = (obs.to(torch.float64) - self.obs_rms.mean) / torch.sqrt(self.obs_rms.var + self.epsilon)
        return o.clamp(-self.clip_obs, self.clip_obs).to(torch.float32)

    def normalize_reward(self, rew):
        if not self.norm_reward:
            return rew
        r = rew.to(torch.float64) / torch.sqrt(self.ret_rms.var + self.epsilon)
        return r.clamp(-self.clip_reward, self.clip_reward).to(torch.float32)

    def reset(self, obs):
        """vec_normalize.py:209-219: the return accumulators restart and, in training, the (all-zero) returns are fed
        to ret_rms once; the observation statistics are NOT updated by a reset."""
        self.ret.zero_()
        if self.training:
            self.ret_rms.update(self.ret)
        return self.normalize_obs(obs)

    def step(self, obs, rew, done):
        """vec_normalize.py:106-127 order: update return, update obs stats, normalise, zero finished returns."""
        self.ret.mul_(self.gamma).add_(rew.to(torch.float64))
        if self.training:
            self.obs_rms.update(obs)
            self.ret_rms.update(self.ret)
        out = self.normalize_obs(obs), self.normalize_reward(rew)
        self.ret.masked_fill_(done.bool(), 0.0)
        return out

    def sync(self, dist):
        self.obs_rms.sync(dist)
        self.ret_rms.sync(dist)

    def state_dict(self):
        return {"obs_mean": self.obs_rms.mean, "obs_var": self.obs_rms.var, "obs_count": self.obs_rms.count,
                "ret_mean": self.ret_rms.mean, "ret_var": self.ret_rms.var, "ret_count": self.ret_rms.count,
                "ret": self.ret}

    def load_state_dict(self, sd):
        for rms, pre in ((self.obs_rms, "obs"), (self.ret_rms, "ret")):
            rms.mean.copy_(torch.as_tensor(sd[pre + "_mean"], dtype=torch.float64))
            rms.var.copy_(torch.as_tensor(sd[pre + "_var"], dtype=torch.float64))
            rms.count.copy_(torch.as_tensor(sd[pre + "_count"], dtype=torch.float64))
        if "ret" in sd and tuple(torch.as_tensor(sd["ret"]).shape) == tuple(self.ret.shape):
            self.ret.copy_(torch.as_tensor(sd["ret"], dtype=torch.float64))

    # ---- the reference's checkpoint format (vec_normalize.py:222-243): a pickle of the VecNormalize object without its
    # venv.  save() writes exactly that when stable_baselines3 is importable — the fork's VecNormalize.load(path, venv)
    # then reads our statistics — and load() reads a file written by the fork's VecNormalize.save (or by save()).
    def save(self, path):
        import pickle
        fields = dict(clip_obs=self.clip_obs, clip_reward=self.clip_reward, gamma=self.gamma, epsilon=self.epsilon,
                      training=self.training, norm_obs=self.norm_obs, norm_reward=self.norm_reward)
        stats = {k: v.detach().cpu().numpy().copy() for k, v in self.state_dict().items()}
        try:
            from stable_baselines3.common.running_mean_std import RunningMeanStd as RefRms
            from stable_baselines3.common.vec_env import VecNormalize as RefVecNormalize
        except Exception:
            with open(path, "wb") as f:
                pickle.dump({"format": "fwb200.DeviceVecNormalize", **fields, **stats}, f)
            return
        import numpy as np
        obj = RefVecNormalize.__new__(RefVecNormalize)
        rms = {}
        for pre, shape in (("obs", stats["obs_mean"].shape), ("ret", ())):
            r = RefRms(shape=shape)
            r.mean, r.var = stats[pre + "_mean"].astype(np.float64), stats[pre + "_var"].astype(np.float64)
            r.count = float(stats[pre + "_count"])
            rms[pre] = r
        # the attributes VecNormalize.__getstate__ keeps (venv, class_attributes and ret are rebuilt by set_venv)
        obj.__dict__.update(dict(fields, obs_keys=None, obs_spaces=None, obs_rms=rms["obs"], ret_rms=rms["ret"],
                                 old_obs=np.array([]), old_reward=np.array([]), venv=None, class_attributes={},
                                 ret=np.zeros(0), num_envs=int(self.ret.shape[0])))
        with open(path, "wb") as f:
            pickle.dump(obj, f)

    def load(self, path):
        import pickle
        with open(path, "rb") as f:
            obj = pickle.load(f)                 # a reference pickle needs stable_baselines3 importable, as in the reference
        if isinstance(obj, dict):
            src = obj
            stats = obj
        else:
            src = obj.__dict__
            stats = {"obs_mean": obj.obs_rms.mean, "obs_var": obj.obs_rms.var, "obs_count": obj.obs_rms.count,
                     "ret_mean": obj.ret_rms.mean, "ret_var": obj.ret_rms.var, "ret_count": obj.ret_rms.count}
        for k in ("clip_obs", "clip_reward", "gamma", "epsilon", "training", "norm_obs", "norm_reward"):
            setattr(self, k, src[k])
        self.load_state_dict(stats)
        return self


FW_ROLLOUT_BLOCKS, FW_PPO_SCRATCH_DOUBLES = 64, 9 + 7 * 592          # include/fwb200.h


def rollout_scratch_doubles(obs_dim):
    """FW_ROLLOUT_SCRATCH(obs_dim) of include/fwb200.h: allocate it ZEROED (torch.zeros), it holds a ticket counter."""
    return 3 * obs_dim + 4 + FW_ROLLOUT_BLOCKS * (2 * obs_dim + 5)


def fused_post_step(norm, buf, obs_raw, rew_raw, done, actions, values, log_probs, last_obs, last_dones, run_ret,
                    run_len, ep_stats, scratch):
    """One step of rollout glue through fw_rollout_post_step (csrc/fw_ppo.cu): `norm` (DeviceVecNormalize) statistics
    and return accumulators are updated, row `buf.pos` of `buf` (RolloutBuffer) receives the PREVIOUS observation /
    done flags with this step's action, value, log-prob and normalised reward, `last_obs` / `last_dones` roll forward,
    `run_ret` / `run_len` / `ep_stats` carry the Monitor-style episode totals.  All CUDA tensors; returns the tensors
    that must stay alive until the launches have run."""
    import ctypes
    from . import _lib
    t = buf.pos
    tens = dict(obs_raw=obs_raw, rew_raw=rew_raw, done=done, actions=actions.contiguous(),
                values=values.reshape(-1).contiguous(), log_probs=log_probs.reshape(-1).contiguous(),
                last_obs=last_obs, last_dones=last_dones, ret=norm.ret, obs_mean=norm.obs_rms.mean,
                obs_var=norm.obs_rms.var, obs_count=norm.obs_rms.count, ret_mean=norm.ret_rms.mean,
                ret_var=norm.ret_rms.var, ret_count=norm.ret_rms.count, run_ret=run_ret, run_len=run_len,
                ep_stats=ep_stats, buf_obs=buf.observations[t], buf_actions=buf.actions[t], buf_rewards=buf.rewards[t],
                buf_dones=buf.dones[t], buf_values=buf.values[t], buf_log_probs=buf.log_probs[t], scratch=scratch)
    assert all(v.is_cuda and v.is_contiguous() for v in tens.values())
    assert scratch.numel() >= rollout_scratch_doubles(obs_raw.shape[1]), "scratch: torch.zeros(rollout_scratch_doubles(obs_dim))"
    p = _lib.FwRolloutPost(**{k: v.data_ptr() for k, v in tens.items()}, n=obs_raw.shape[0], obs_dim=obs_raw.shape[1],
                           act_dim=actions.shape[1], gamma=norm.gamma, clip_obs=norm.clip_obs,
                           clip_reward=norm.clip_reward, epsilon=norm.epsilon, norm_obs=int(norm.norm_obs),
                           norm_reward=int(norm.norm_reward), training=int(norm.training))
    _lib.check(_lib.lib().fw_rollout_post_step(ctypes.byref(p), ctypes.c_void_p(
        torch.cuda.current_stream(obs_raw.device).cuda_stream)), "fw_rollout_post_step")
    buf.pos += 1
    if buf.pos == buf.buffer_size:
        buf.full = True
    return tens
