"""PPO over the batched CUDA simulator: rollout collection, GAE and the clipped-surrogate update all stay in HBM.

Re-hosts the reference's on-policy loop on device tensors:
  OnPolicyAlgorithm.collect_rollouts / learn   (stable_baselines3/common/on_policy_algorithm.py:123-246)
  PPO.train                                    (stable_baselines3/ppo/ppo.py:133-240; defaults :69-79)
  ActorCriticPolicy, MlpPolicy                 (stable_baselines3/common/policies.py:331-637; net_arch pi/vf [64,64], tanh,
                                               state-independent log_std initialised to 0)
The policy MLP stays in PyTorch (north_star).  Data parallel: one process per GPU, envs sharded by global env id, no
traffic on the env step; per optimiser step ONE all-reduce over the flattened gradient, per rollout one small
all-reduce of the VecNormalize moments.  Advantage normalisation is per minibatch and rank-local as in ppo.py:170.
"""
import math
import os
import time

import torch
import torch.nn as nn

from .buffers import DeviceVecNormalize, RolloutBuffer
from .vec_env import FixedWingVecEnv


class _LinearSplitK(torch.autograd.Function):
    """y = x W^T + b whose weight gradient is a split-K batched GEMM.  For a [B, 64] activation with B = 32 768 the
    weight gradient g^T x has a 64 x 64 (or smaller) output and K = B: cuBLAS runs it in ONE thread block (95 us per
    layer, 60 % of the whole PPO minibatch update, tools/ppo_update_prof.py); cut into S row chunks it is one bmm launch
    over S blocks plus a tiny sum."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        B = x.shape[0]
        # row chunks: as many as keep the partial sums [S, out, in] small (2 M elements) and divide the batch
        S = 128
        while S > 1 and (B % S or S * weight.numel() > (1 << 21)):
            S //= 2
        gi = g @ weight if ctx.needs_input_grad[0] else None
        gs = g.view(S, B // S, g.shape[1])
        gw = torch.bmm(gs.transpose(1, 2), x.view(S, B // S, x.shape[1])).sum(0)
        gb = gs.sum(1).sum(0)
        return gi, gw, gb


class SplitKLinear(nn.Linear):
    """nn.Linear (same parameters, same state_dict) with the split-K weight gradient for large training batches."""

    def forward(self, x):
        if x.is_cuda and x.dim() == 2 and x.shape[0] >= 4096 and torch.is_grad_enabled() and x.is_contiguous():
            return _LinearSplitK.apply(x, self.weight, self.bias)
        return nn.functional.linear(x, self.weight, self.bias)


class ActorCritic(nn.Module):
    def __init__(self, obs_dim=14, action_dim=3, hidden=(64, 64), log_std_init=0.0):
        super().__init__()

        def mlp():
            layers, d = [], obs_dim
            for h in hidden:
                layers += [SplitKLinear(d, h), nn.Tanh()]
                d = h
            return nn.Sequential(*layers), d

        self.pi, d_pi = mlp()
        self.vf, d_vf = mlp()
        self.action_net = SplitKLinear(d_pi, action_dim)
        self.value_net = SplitKLinear(d_vf, 1)
        self.log_std = nn.Parameter(torch.full((action_dim,), float(log_std_init)))
        # SB3 orthogonal init: sqrt(2) for the extractors, 0.01 for the action head, 1 for the value head
        for m in list(self.pi) + list(self.vf):
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, gain=math.sqrt(2))
                nn.init.zeros_(m.bias)
        nn.init.orthogonal_(self.action_net.weight, gain=0.01)
        nn.init.zeros_(self.action_net.bias)
        nn.init.orthogonal_(self.value_net.weight, gain=1.0)
        nn.init.zeros_(self.value_net.bias)

    def _dist(self, obs):
        mean = self.action_net(self.pi(obs))
        return mean, self.log_std.expand_as(mean)

    @staticmethod
    def _log_prob(actions, mean, log_std):
        var = torch.exp(2 * log_std)
        return (-((actions - mean) ** 2) / (2 * var) - log_std - 0.5 * math.log(2 * math.pi)).sum(dim=1)

    def forward(self, obs, deterministic=False):
        mean, log_std = self._dist(obs)
        actions = mean if deterministic else mean + torch.exp(log_std) * torch.randn_like(mean)
        values = self.value_net(self.vf(obs)).squeeze(-1)
        return actions, values, self._log_prob(actions, mean, log_std)

    def evaluate_actions(self, obs, actions):
        mean, log_std = self._dist(obs)
        values = self.value_net(self.vf(obs)).squeeze(-1)
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + log_std).sum(dim=1)
        return values, self._log_prob(actions, mean, log_std), entropy

    def predict_values(self, obs):
        return self.value_net(self.vf(obs)).squeeze(-1)


class FusedPPOLoss(torch.autograd.Function):
    """loss, policy_loss, value_loss = FusedPPOLoss.apply(mean, values, log_std, actions, old_log_prob, advantages,
    returns, clip_range, ent_coef, vf_coef): the block of ppo.py:163-207 between the network outputs and the scalar
    loss, forward and backward, in three launches of fw_ppo_loss (csrc/fw_ppo.cu) instead of ~75 elementwise kernels.
    The networks stay in PyTorch: autograd receives d loss / d mean, d values, d log_std from the kernel."""

    @staticmethod
    def forward(ctx, mean, values, log_std, actions, old_log_prob, advantages, returns, clip_range, ent_coef, vf_coef):
        import ctypes
        from . import _lib
        args = [t.detach().contiguous() for t in (mean, values, log_std, actions, old_log_prob, advantages, returns)]
        assert all(t.is_cuda and t.dtype == torch.float32 for t in args)
        B = args[0].shape[0]
        assert args[0].shape == (B, 3) and args[3].shape == (B, 3) and all(t.shape == (B,) for t in (args[1], args[4], args[5], args[6]))
        dev = args[0].device
        g_mean, g_val, g_ls = torch.empty_like(args[0]), torch.empty_like(args[1]), torch.empty_like(args[2])
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        from .buffers import FW_PPO_SCRATCH_DOUBLES
        scratch = torch.empty(FW_PPO_SCRATCH_DOUBLES, dtype=torch.float64, device=dev)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().fw_ppo_loss(*[ptr(t) for t in args], B, float(clip_range), float(ent_coef), float(vf_coef),
                                          ptr(scratch), ptr(g_mean), ptr(g_val), ptr(g_ls), ptr(losses),
                                          ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fw_ppo_loss")
        ctx.save_for_backward(g_mean, g_val, g_ls)
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g_loss, _g_losses):
        g_mean, g_val, g_ls = ctx.saved_tensors
        return g_loss * g_mean, g_loss * g_val, g_loss * g_ls, None, None, None, None, None, None, None


class FlatAdam:
    """Adam + gradient-norm clipping over ONE flat buffer (fw_adam_clip_step, csrc/fw_ppo.cu).  The module's parameters
    and their .grad become views of `flat` / `grad`, so zero_grad is one memset, the data-parallel all-reduce needs no
    gather/scatter, and clip + step is one launch instead of ~25 foreach kernels.  Same update rule as
    torch.optim.Adam(lr, betas=(0.9, 0.999), eps, weight_decay=0) after clip_grad_norm_(max_grad_norm)."""

    def __init__(self, module, lr, eps=1e-5, betas=(0.9, 0.999), max_grad_norm=0.5):
        self.params = [p for p in module.parameters()]
        dev = self.params[0].device
        assert dev.type == "cuda" and all(p.dtype == torch.float32 for p in self.params)
        self.flat = torch.cat([p.data.reshape(-1) for p in self.params]).contiguous()
        self.grad = torch.zeros_like(self.flat)
        off = 0
        for p in self.params:
            n = p.numel()
            p.data = self.flat[off:off + n].view_as(p)
            p.grad = self.grad[off:off + n].view_as(p)
            off += n
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr, self.eps, self.betas, self.max_grad_norm = float(lr), float(eps), betas, float(max_grad_norm)

    def zero_grad(self, set_to_none=False):
        self.grad.zero_()

    def enable_peer_allreduce(self, dist):
        """Data-parallel runs on one node: step() becomes fw_comm_allreduce_adam — gradient mean over all ranks (NVLink
        peer memory, ranks added in rank order), clip and Adam in ONE launch instead of ncclAllReduce + divide + step.
        Returns False (and changes nothing) when the peers' buffers cannot be mapped."""
        import ctypes
        from . import _lib
        world, rank = dist.get_world_size(), dist.get_rank()
        if world < 2 or world > 8:
            return False
        L = _lib.lib()
        h = ctypes.c_void_p()
        dev = self.flat.device
        rc = L.fw_comm_create(self.flat.numel(), world, rank, dev.index if dev.index is not None else 0, ctypes.byref(h))
        mine = ctypes.create_string_buffer(64)
        if rc == 0:
            rc = L.fw_comm_export(h, mine)
        # every rank must take the same branch: exchange (status, handle) and agree
        payload = torch.tensor([1 if rc == 0 else 0] + list(mine.raw), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(payload) for _ in range(world)]
        dist.all_gather(gathered, payload)
        ok = all(int(g[0]) == 1 for g in gathered)
        if ok:
            blob = b"".join(bytes(g[1:].cpu().tolist()) for g in gathered)
            rc = L.fw_comm_connect(h, ctypes.c_char_p(blob))
            flag = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item()) == 1
        if not ok:
            if h:
                L.fw_comm_destroy(h)
            return False
        self._comm = h
        return True

    def comm_error(self):
        """True when a peer failed to arrive inside fw_comm_allreduce_adam (checked by the trainer once per iteration)."""
        from . import _lib
        return bool(getattr(self, "_comm", None)) and _lib.lib().fw_comm_error(self._comm) == 1

    def step(self):
        import ctypes
        from . import _lib
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        if getattr(self, "_comm", None):
            _lib.check(_lib.lib().fw_comm_allreduce_adam(self._comm, ptr(self.flat), ptr(self.grad), ptr(self.exp_avg),
                                                         ptr(self.exp_avg_sq), ptr(self.step_count), self.flat.numel(),
                                                         self.lr, self.betas[0], self.betas[1], self.eps, self.max_grad_norm,
                                                         ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)),
                       "fw_comm_allreduce_adam")
            return
        _lib.check(_lib.lib().fw_adam_clip_step(ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                                ptr(self.step_count), self.flat.numel(), self.lr, self.betas[0],
                                                self.betas[1], self.eps, self.max_grad_norm, ctypes.c_void_p(
                                                    torch.cuda.current_stream(self.flat.device).cuda_stream)),
                   "fw_adam_clip_step")

    def state_dict(self):
        return {"flat_adam": True, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_count,
                "lr": self.lr, "eps": self.eps, "betas": self.betas, "max_grad_norm": self.max_grad_norm}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.step_count.copy_(sd["step"])
        self.lr, self.eps, self.betas = float(sd["lr"]), float(sd["eps"]), tuple(sd["betas"])


def allreduce_gradients(params, dist, world_size):
    """One all-reduce(SUM) over the flattened gradient (~10.5k fp32 = 42 KB: latency bound), then the mean."""
    grads = [p.grad for p in params if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat /= world_size
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat


class PPO:
    """learning_rate 3e-4, n_epochs 10, gamma 0.99, gae_lambda 0.95, clip_range 0.2, ent_coef 0, vf_coef 0.5,
    max_grad_norm 0.5 are the reference defaults (ppo.py:69-79).  n_steps / batch_size default to values that make
    sense at thousands of envs per GPU (SB3's 2048 / 64 are sized for a handful of envs)."""

    def __init__(self, env: FixedWingVecEnv, n_steps=32, batch_size=32768, n_epochs=10, learning_rate=3e-4, gamma=0.99,
                 gae_lambda=0.95, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, normalize=True,
                 seed=0, dist=None, use_cuda_graph=True, fused_loss=True, fused_rollout=True,
                 flat_optimizer=True, allow_eager_fallback=False):
        self.env = env
        # a failed CUDA-graph capture raises unless this is set: a silently eager run is ~10x slower and would make
        # every throughput figure taken from it meaningless
        self.allow_eager_fallback = bool(allow_eager_fallback)
        self.device = env.device
        self.n_envs = env.num_envs
        self.n_steps, self.batch_size, self.n_epochs = n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda, self.clip_range = gamma, gae_lambda, clip_range
        self.ent_coef, self.vf_coef, self.max_grad_norm = ent_coef, vf_coef, max_grad_norm
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        torch.manual_seed(seed)                      # identical initial weights on every rank
        self.obs_dim = int(getattr(getattr(env, "sim", None), "obs_dim", 14))    # 14 by default; general layouts, waypoint env
        self.policy = ActorCritic(obs_dim=self.obs_dim).to(self.device)
        torch.manual_seed(seed + 1000 * (self.rank + 1))   # different action noise per rank
        self.use_cuda_graph = bool(use_cuda_graph)
        # capture the NCCL gradient all-reduce inside the minibatch-update graph (data-parallel runs)
        self.graph_allreduce = os.environ.get("FWB200_PPO_GRAPH_ALLREDUCE", "1") != "0"
        self.fused_loss = bool(fused_loss)
        self.fused_rollout = bool(fused_rollout)
        self.two_stream_nets = os.environ.get("FWB200_PPO_TWO_STREAMS", "1") != "0"
        self._vf_stream = None
        self.flat_optimizer = bool(flat_optimizer) and self.device.type == "cuda"
        if self.flat_optimizer:
            self.optimizer = FlatAdam(self.policy, lr=learning_rate, eps=1e-5, max_grad_norm=max_grad_norm)
        else:
            self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=learning_rate, eps=1e-5,
                                              capturable=self.use_cuda_graph)
        # data parallel on one node: fuse the gradient all-reduce into the optimiser kernel over NVLink peer memory
        self.peer_allreduce = False
        if self.flat_optimizer and self.world > 1 and os.environ.get("FWB200_PPO_PEER_ALLREDUCE", "1") != "0":
            self.peer_allreduce = self.optimizer.enable_peer_allreduce(dist)
        self._rollout_graph = None
        self._train_graph = None
        self._eager_rollouts = 0
        self.buffer = RolloutBuffer(n_steps, self.n_envs, obs_dim=self.obs_dim, device=self.device,
                                    gae_lambda=gae_lambda, gamma=gamma)
        self.norm = DeviceVecNormalize(self.n_envs, obs_dim=self.obs_dim, device=self.device, gamma=gamma, norm_obs=normalize,
                                       norm_reward=normalize)
        self.num_timesteps = 0
        self._iteration = 0
        self._t_start = None
        self._last_obs = None
        self._last_dones = None
        # on-device episode statistics (Monitor-equivalent, no per-env python objects)
        self._ep_stats = torch.zeros(3, dtype=torch.float64, device=self.device)     # one buffer: fw_rollout_post_step
        self.ep_ret_sum, self.ep_len_sum, self.ep_count = self._ep_stats[0], self._ep_stats[1], self._ep_stats[2]
        from .buffers import rollout_scratch_doubles
        self._post_scratch = torch.zeros(rollout_scratch_doubles(self.obs_dim), dtype=torch.float64, device=self.device)
        self._run_ret = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)
        self._run_len = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)
        self.logs = []

    # ------------------------------------------------------------------ rollout
    def _setup(self):
        raw = self.env.reset_tensor()
        self._last_obs = self.norm.reset(raw).clone()
        self._last_dones = torch.zeros(self.n_envs, dtype=torch.float32, device=self.device)

    def _rollout_step(self, t):
        with torch.no_grad():
            if self.two_stream_nets and self._last_obs.is_cuda:
                # value head on the second stream, policy head + sampling on this one (see _loss)
                cur = torch.cuda.current_stream(self.device)
                if self._vf_stream is None:
                    self._vf_stream = torch.cuda.Stream(self.device)
                self._vf_stream.wait_stream(cur)
                with torch.cuda.stream(self._vf_stream):
                    values = self.policy.predict_values(self._last_obs)
                mean, log_std = self.policy._dist(self._last_obs)
                actions = mean + torch.exp(log_std) * torch.randn_like(mean)
                log_probs = self.policy._log_prob(actions, mean, log_std)
                cur.wait_stream(self._vf_stream)
                values.record_stream(cur)
            else:
                actions, values, log_probs = self.policy(self._last_obs)
        obs_raw, rew_raw, done = self.env.step_tensor(actions.contiguous())
        if self.fused_rollout and obs_raw.is_cuda:
            return self._fused_post_step(t, obs_raw, rew_raw, done, actions, values, log_probs)
        d = done.bool()
        self._run_ret.add_(rew_raw.to(torch.float64))
        self._run_len.add_(1.0)
        self.ep_ret_sum.add_((self._run_ret * d).sum())
        self.ep_len_sum.add_((self._run_len * d).sum())
        self.ep_count.add_(d.sum())
        self._run_ret.masked_fill_(d, 0.0)
        self._run_len.masked_fill_(d, 0.0)
        obs, rew = self.norm.step(obs_raw, rew_raw, done)
        self.buffer.add(self._last_obs, actions, rew, self._last_dones, values, log_probs)
        # static buffers (in-place) so that the whole rollout can be replayed as one CUDA graph
        self._last_obs.copy_(obs)
        self._last_dones.copy_(done)

    def _fused_post_step(self, t, obs_raw, rew_raw, done, actions, values, log_probs):
        """VecNormalize.step + RolloutBuffer.add + episode bookkeeping of one step in a few launches
        (buffers.fused_post_step -> fw_rollout_post_step) instead of ~70 tensor ops; float64 statistics."""
        from .buffers import fused_post_step
        self._post_keep = fused_post_step(self.norm, self.buffer, obs_raw, rew_raw, done, actions, values, log_probs,
                                          self._last_obs, self._last_dones, self._run_ret, self._run_len,
                                          self._ep_stats, self._post_scratch)

    def _rollout_body(self):
        self.buffer.reset()
        for t in range(self.n_steps):
            self._rollout_step(t)
        with torch.no_grad():
            last_values = self.policy.predict_values(self._last_obs)
        self.buffer.compute_returns_and_advantage(last_values, self._last_dones)
        self.env.sim.join()       # the simulator's side-stream work joins this stream (a graph capture must end joined)

    def collect_rollouts(self):
        """One rollout of n_steps over all envs + GAE.  With use_cuda_graph the ~40 small launches per step (policy
        MLP, the three simulator kernels, normaliser, bookkeeping) are captured once and replayed: the loop is launch
        bound otherwise (0.9 ms/step eager vs 0.1 ms of simulator work at 8192 envs)."""
        if self.use_cuda_graph and self._rollout_graph is None and self._eager_rollouts >= 1:
            # the first rollout ran eagerly (lazy initialisation, allocator warm-up); capture the second one
            self.env.sim.join()
            torch.cuda.synchronize(self.device)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._rollout_body()
                self._rollout_graph = g
            except Exception as e:
                if not self.allow_eager_fallback:
                    raise RuntimeError("PPO: CUDA-graph capture of the rollout failed (pass allow_eager_fallback=True "
                                       "or use_cuda_graph=False to run eagerly)") from e
                self.use_cuda_graph = False
                self._rollout_graph = None
                self.logs.append({"cuda_graph_disabled": repr(e)})
                torch.cuda.synchronize(self.device)
        if self._rollout_graph is not None:
            self._rollout_graph.replay()
            self.buffer.pos, self.buffer.full = self.n_steps, True
        else:
            self._rollout_body()
            self._eager_rollouts += 1
        self.num_timesteps += self.n_steps * self.n_envs * self.world
        if self.dist is not None:
            self.norm.sync(self.dist)

    # ------------------------------------------------------------------ update (ppo.py:133-240)
    def _loss(self, batch):
        """ppo.py:163-207.  On the GPU the part after the networks is the fused fw_ppo_loss kernel."""
        if self.fused_loss and batch.observations.is_cuda:
            # the policy and the value network are independent until the loss: the value branch runs on a second
            # stream (autograd replays its backward there as well), so that inside the captured update graph the two
            # chains of small, latency-bound kernels overlap
            cur = torch.cuda.current_stream(self.device)
            if self.two_stream_nets:
                if self._vf_stream is None:
                    self._vf_stream = torch.cuda.Stream(self.device)
                self._vf_stream.wait_stream(cur)
                with torch.cuda.stream(self._vf_stream):
                    values = self.policy.value_net(self.policy.vf(batch.observations)).squeeze(-1)
                mean = self.policy.action_net(self.policy.pi(batch.observations))
                cur.wait_stream(self._vf_stream)
                values.record_stream(cur)
            else:
                mean = self.policy.action_net(self.policy.pi(batch.observations))
                values = self.policy.value_net(self.policy.vf(batch.observations)).squeeze(-1)
            loss, parts = FusedPPOLoss.apply(mean, values, self.policy.log_std, batch.actions, batch.old_log_prob,
                                             batch.advantages, batch.returns, self.clip_range, self.ent_coef, self.vf_coef)
            return loss, parts[1], parts[2]
        values, log_prob, entropy = self.policy.evaluate_actions(batch.observations, batch.actions)
        adv = batch.advantages
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        ratio = torch.exp(log_prob - batch.old_log_prob)
        pl1 = adv * ratio
        pl2 = adv * torch.clamp(ratio, 1 - self.clip_range, 1 + self.clip_range)
        policy_loss = -torch.min(pl1, pl2).mean()
        value_loss = torch.nn.functional.mse_loss(batch.returns, values)
        entropy_loss = -entropy.mean()
        loss = policy_loss + self.ent_coef * entropy_loss + self.vf_coef * value_loss
        return loss, policy_loss, value_loss

    def _minibatch_update(self, batch):
        loss, policy_loss, value_loss = self._loss(batch)
        self.optimizer.zero_grad(set_to_none=False)
        loss.backward()
        if self.flat_optimizer:
            if self.dist is not None and self.world > 1 and not self.peer_allreduce:
                self.dist.all_reduce(self.optimizer.grad)      # the gradient already is one flat buffer
                self.optimizer.grad.div_(self.world)
            self.optimizer.step()                              # (peer all-reduce +) clip_grad_norm_ + Adam in one launch
            return policy_loss.detach(), value_loss.detach()
        if self.dist is not None and self.world > 1:
            allreduce_gradients(self._params, self.dist, self.world)
        torch.nn.utils.clip_grad_norm_(self._params, self.max_grad_norm)
        self.optimizer.step()
        return policy_loss.detach(), value_loss.detach()

    def train(self):
        self._params = [p for p in self.policy.parameters()]
        total = self.n_steps * self.n_envs
        bs = min(self.batch_size, total)
        flat = [self.buffer.flat(x) for x in (self.buffer.observations, self.buffer.actions, self.buffer.values,
                                              self.buffer.log_probs, self.buffer.advantages, self.buffer.returns)]
        graph_ok = self.use_cuda_graph and (self.world == 1 or self.graph_allreduce) and total % bs == 0
        if graph_ok and self._train_graph is None:
            from .buffers import RolloutBufferSamples
            self._mb = RolloutBufferSamples(*(torch.empty(bs, *f.shape[1:], dtype=f.dtype, device=self.device) for f in flat))
            for f, dst in zip(flat, self._mb):
                dst.copy_(f[:bs])
            self._pl = torch.zeros((), device=self.device)
            self._vl = torch.zeros((), device=self.device)
            # The warm-up below runs real optimiser steps on the first (unshuffled) rows, which the reference's
            # PPO.train never performs: snapshot weights and optimiser state and put them back afterwards, so that the
            # graphed and the eager path produce the same sequence of updates from the first iteration on.
            snap = self._snapshot_optimizer()
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                 # warm-up on a side stream as torch.cuda.graph requires
                for _ in range(2):
                    self._minibatch_update(self._mb)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    pl, vl = self._minibatch_update(self._mb)
                    self._pl.copy_(pl)
                    self._vl.copy_(vl)
                self._train_graph = g
            except Exception as e:
                if not self.allow_eager_fallback:
                    raise RuntimeError("PPO: CUDA-graph capture of the minibatch update failed (pass "
                                       "allow_eager_fallback=True or use_cuda_graph=False to run eagerly)") from e
                self.logs.append({"train_cuda_graph_disabled": repr(e)})
                self._train_graph = False
                torch.cuda.synchronize(self.device)
            self._restore_optimizer(snap)
        policy_loss = value_loss = None
        for epoch in range(self.n_epochs):
            perm = torch.randperm(total, device=self.device)
            for start in range(0, total, bs):
                idx = perm[start:start + bs]
                if graph_ok and self._train_graph:
                    for f, dst in zip(flat, self._mb):
                        torch.index_select(f, 0, idx, out=dst)
                    self._train_graph.replay()
                    policy_loss, value_loss = self._pl, self._vl
                else:
                    from .buffers import RolloutBufferSamples
                    batch = RolloutBufferSamples(*(f.index_select(0, idx) for f in flat))
                    policy_loss, value_loss = self._minibatch_update(batch)
        return dict(policy_loss=policy_loss, value_loss=value_loss, std=torch.exp(self.policy.log_std).mean().detach())

    def _snapshot_optimizer(self):
        if self.flat_optimizer:
            o = self.optimizer
            return [t.clone() for t in (o.flat, o.exp_avg, o.exp_avg_sq, o.step_count)]
        import copy
        return (copy.deepcopy(self.policy.state_dict()), copy.deepcopy(self.optimizer.state_dict()))

    def _restore_optimizer(self, snap):
        if self.flat_optimizer:
            o = self.optimizer
            for dst, src in zip((o.flat, o.exp_avg, o.exp_avg_sq, o.step_count), snap):
                dst.copy_(src)                       # in place: the captured graph holds these addresses
            return
        # torch.optim.Adam(capturable=True): state tensors are updated in place by the graph — restore in place too
        cur_p, cur_o = self.policy.state_dict(), self.optimizer.state_dict()
        for k, v in snap[0].items():
            cur_p[k].copy_(v)
        for pid, st in snap[1]["state"].items():
            for k, v in st.items():
                if torch.is_tensor(v):
                    cur_o["state"][pid][k].copy_(v)
        for pid in list(cur_o["state"].keys()):
            if pid not in snap[1]["state"]:          # state created by the warm-up: back to "never stepped"
                for k, v in cur_o["state"][pid].items():
                    if torch.is_tensor(v):
                        v.zero_()

    def learn(self, total_timesteps, log_interval=1, callback=None):
        if self._last_obs is None:
            self._setup()
        if getattr(self, "_t_start", None) is None:
            self._t_start = time.time()
        t0 = self._t_start
        while self.num_timesteps < total_timesteps:
            self.collect_rollouts()
            stats = self.train()
            self._iteration += 1
            it = self._iteration
            if it % log_interval == 0:
                packed = torch.stack([self.ep_ret_sum, self.ep_len_sum, self.ep_count])
                if self.dist is not None and self.world > 1:
                    self.dist.all_reduce(packed)
                r, l, c = packed.tolist()
                self.ep_ret_sum.zero_(); self.ep_len_sum.zero_(); self.ep_count.zero_()
                row = {"iteration": it, "timesteps": self.num_timesteps, "fps": self.num_timesteps / (time.time() - t0),
                       "ep_rew_mean": r / c if c else float("nan"), "ep_len_mean": l / c if c else float("nan"),
                       "episodes": int(c), **{k: float(v) for k, v in stats.items()}}
                self.logs.append(row)
                if callback is not None:
                    callback(row)
        if self.peer_allreduce and self.optimizer.comm_error():
            raise RuntimeError("fw_comm_allreduce_adam: a peer rank did not arrive within the time limit; the replicas are "
                               "no longer in step")
        return self

    # ------------------------------------------------------------------ checkpoint / resume
    def save(self, path, vecnormalize_path=None):
        """Everything a bit-identical continuation needs (SURVEY §5 "checkpoint / resume"; the reference keeps only the
        model zip, common/base_class.py:645, and the VecNormalize pickle, simple_train.py:733-751, and restarts the
        episodes): policy, optimiser, normaliser, the rollout carry (last observation / done flags, running episode
        totals), the simulator's whole env state (BatchedFixedWing.get_state) and the torch RNG states.
        `vecnormalize_path` additionally writes the normaliser in the reference's VecNormalize.save format."""
        torch.cuda.synchronize(self.device)
        opt = self.optimizer.state_dict()
        if self.flat_optimizer:
            opt = dict(opt, flat=self.optimizer.flat)
        torch.save({"policy": self.policy.state_dict(), "optimizer": opt, "normalizer": self.norm.state_dict(),
                    "num_timesteps": self.num_timesteps, "iteration": getattr(self, "_iteration", 0),
                    "last_obs": self._last_obs, "last_dones": self._last_dones, "run_ret": self._run_ret,
                    "run_len": self._run_len, "ep_stats": self._ep_stats,
                    "env_state": self.env.sim.get_state().cpu(), "curriculum_level": getattr(self.env, "curriculum_level", 1.0),
                    "rng_cpu": torch.get_rng_state(), "rng_cuda": torch.cuda.get_rng_state(self.device)}, path)
        if vecnormalize_path is not None:
            self.norm.save(vecnormalize_path)

    def load(self, path):
        sd = torch.load(path, map_location=self.device, weights_only=False)
        self.policy.load_state_dict(sd["policy"])
        opt = dict(sd["optimizer"])
        flat = opt.pop("flat", None)
        self.optimizer.load_state_dict(opt)
        if self.flat_optimizer and flat is not None:
            self.optimizer.flat.copy_(flat)              # the parameters are views of it
        self.norm.load_state_dict(sd["normalizer"])
        self.num_timesteps = sd["num_timesteps"]
        self._iteration = sd.get("iteration", 0)
        if sd.get("env_state") is not None:
            self.env.sim.set_state(sd["env_state"])
            n, dev = self.n_envs, self.device
            if self._last_obs is None:
                self._last_obs = torch.zeros(n, self.obs_dim, dtype=torch.float32, device=dev)
                self._last_dones = torch.zeros(n, dtype=torch.float32, device=dev)
            for dst, key in ((self._last_obs, "last_obs"), (self._last_dones, "last_dones"), (self._run_ret, "run_ret"),
                             (self._run_len, "run_len"), (self._ep_stats, "ep_stats")):
                dst.copy_(sd[key])                       # in place: captured graphs keep these addresses
            torch.set_rng_state(sd["rng_cpu"].cpu())
            torch.cuda.set_rng_state(sd["rng_cuda"].cpu(), self.device)
        return self
