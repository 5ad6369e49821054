"""FixedWingVecEnv — the SB3 `VecEnv` contract over the batched CUDA simulator, and the single-env
`FixedWingAircraft` gym-style API on top of a one-env batch.

Drop-in boundary: what `collect_rollouts` / `evaluate_policy` / the reference scripts call on
`VecNormalize(SubprocVecEnv([make_env]*N))`:
  reset(), step_async(), step_wait(), step(), close(), seed(), get_attr(), set_attr(), env_method(), num_envs,
  observation_space, action_space              (stable_baselines3/common/vec_env/base_vec_env.py:48-224)
  auto-reset + info["terminal_observation"]    (subproc_vec_env.py:26-31, dummy_vec_env.py:46-50)
  info["episode"] = {"r","l","t"}              (common/monitor.py:99-113)
  info["target"], info["termination"], the 9 metric dicts on done   (fixed_wing.py:519-626)
When stable_baselines3 is importable the classes here ARE `stable_baselines3.common.vec_env.VecEnv`s (derived, or
registered with the ABC if stable_baselines3 was imported after this module) with `gym.spaces.Box` spaces, so the fork's
`BaseAlgorithm._wrap_env` (common/base_class.py:173-177) and `VecNormalize` (vec_env/vec_normalize.py:28-39) take them
as they are — tests/test_gpu_dropin.py runs the fork's own `PPO.learn` on them.
There are no worker processes and no per-env Python objects: one CUDA launch steps every env.  Host copies happen
only at this numpy edge (pinned buffers); `step_tensor` skips them entirely.
"""
import sys
import time

import numpy as np
import torch

from . import batched as bt
from .config import (ANGULAR_TARGET_STATES, GOAL_STATES, METRIC_LAYOUT, TARGET_STATES, TERM_NAMES, action_space_bounds,
                     build_config, observation_space_bounds, resolve_configs)


class Box:
    """Minimal stand-in for gym.spaces.Box, used only when `gym` cannot be imported (it is not a dependency)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def seed(self, seed=None):
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low) & (np.abs(self.low) < 1e30), self.low, -1.0)
        hi = np.where(np.isfinite(self.high) & (np.abs(self.high) < 1e30), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s)" % (self.shape, self.dtype)


def make_box(low, high):
    """gym.spaces.Box (fixed_wing.py:245-258) when gym is importable — SB3 dispatches on isinstance(space, gym.spaces.Box)
    (common/preprocessing.py, on_policy_algorithm.py:158) — else the stand-in above."""
    low, high = np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32)
    try:
        import gym.spaces
        return gym.spaces.Box(low=low, high=high, dtype=np.float32)
    except Exception:
        return Box(low, high)


def _sb3_vecenv():
    try:
        from stable_baselines3.common.vec_env.base_vec_env import VecEnv
        return VecEnv
    except Exception:
        return None


_SB3VecEnv = _sb3_vecenv()


class _VecEnvSurface(_SB3VecEnv if _SB3VecEnv is not None else object):
    """The non-abstract part of stable_baselines3's VecEnv (base_vec_env.py:141-224), spelled out so that it is there
    whether the class derives from VecEnv (stable_baselines3 importable when this module is loaded) or is registered
    with the ABC afterwards."""

    metadata = {"render.modes": []}

    def _init_vecenv(self, num_envs, observation_space, action_space):
        self.num_envs = int(num_envs)
        self.observation_space = observation_space
        self.action_space = action_space
        register_with_sb3()                          # in case stable_baselines3 was imported after this module

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_images(self):
        raise NotImplementedError("rendering is out of scope (SURVEY §2 rows 15-16)")

    def render(self, mode="human"):
        return None

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._get_indices(indices)]

    @property
    def unwrapped(self):
        return self

    def getattr_depth_check(self, name, already_found):
        if hasattr(self, name) and already_found:
            return "%s.%s" % (type(self).__module__, type(self).__name__)
        return None

    def _get_indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, (int, np.integer)):
            return [int(indices)]
        return indices


class _SimulatorView:
    """What the scripts read through get_attr("simulator") (evaluate_controller.py:120): dt and state values."""

    def __init__(self, venv, index):
        self._venv, self._i = venv, index
        self.dt = venv.cfg.dt

    @property
    def state(self):
        return self._venv.state_dict(self._i)


_EMPTY_INFO = {}


def _make_done_info():
    """Compiles the episode-end info builder from METRIC_LAYOUT as ONE dict display (several times faster than a loop
    of dict(zip(...)) — at ten finished episodes per step the loop cost more than a kernel).  row: 28 metrics |
    return | length | termination code, as python floats."""
    parts = ['"termination": TERM_NAMES.get(int(row[30]), int(row[30]))']
    for name, off, keys in METRIC_LAYOUT:
        conv = "bool(row[%d])" if name == "success" else "row[%d]"
        parts.append('"%s": {%s}' % (name, ", ".join('"%s": %s' % (k, conv % (off + q)) for q, k in enumerate(keys))))
    parts.append('"terminal_observation": term_obs')
    parts.append('"episode": {"r": row[28], "l": int(row[29]), "t": now}')
    src = "def _done_info(row, term_obs, now):\n    return {%s}\n" % ", ".join(parts)
    ns = {"TERM_NAMES": TERM_NAMES}
    exec(src, ns)
    return ns["_done_info"]


_done_info = _make_done_info()


def register_with_sb3():
    """Registers the adapters with stable_baselines3's VecEnv ABC when that package was imported AFTER this module (when
    it was importable at load time they derive from it).  Returns True if they are VecEnvs now."""
    mod = sys.modules.get("stable_baselines3.common.vec_env.base_vec_env")
    if mod is None:
        return False
    for cls in (FixedWingVecEnv, WaypointVecEnv):
        if not issubclass(cls, mod.VecEnv):
            mod.VecEnv.register(cls)
    return True


class FixedWingVecEnv(_VecEnvSurface):
    """n_envs reference-semantics fixed-wing envs on one GPU behind the VecEnv API.

    info_mode: "compat" builds the reference's per-env info dict for every env every step (info["target"] always
    present, fixed_wing.py:626); "lazy" (default) builds dicts only for envs that finished and hands out one shared
    empty dict for the rest, in one persistent list that is valid until the next step — at tens of thousands of envs
    the dict loop (or even copying the list), not the simulator, bounds the step rate.
    copy_outputs: False (default) returns views of double-buffered pinned host arrays that stay valid until the
    second-next step() (enough for SB3's collect_rollouts); True returns fresh copies every step.
    """

    def __init__(self, num_envs, config_path=None, config_kw=None, sim_config_path=None, sim_config_kw=None,
                 device=0, seed=0, env_id_offset=0, precision="f64", integrator="rk45", rk4_substeps=4,
                 info_mode="lazy", copy_outputs=False):
        self._build_kw = dict(env_cfg=config_path, sim_cfg=sim_config_path, config_kw=config_kw,
                              sim_config_kw=sim_config_kw, precision=precision, integrator=integrator,
                              rk4_substeps=rk4_substeps, seed=seed, env_id_offset=env_id_offset)
        self.cfg = build_config(**self._build_kw)                       # the flattened POD (include/fwb200.h FwConfig)
        # the reference-format dicts `env.cfg` / `env.simulator.cfg` (what get_attr("cfg") hands to the scripts)
        self.env_config, self.sim_config = resolve_configs(config_path, sim_config_path, config_kw, sim_config_kw)
        self.sim = bt.BatchedFixedWing(int(num_envs), cfg=self.cfg, device=device)
        self.device = self.sim.device
        # the reference's own spaces for any layout (fixed_wing.py:92-258): per-entry bounds, [length, n] for
        # observation.shape "matrix" — the numpy edge hands out observations in that shape (views of the flat rows the
        # kernels write, which are the row-major matrix); the tensor fast path stays flat
        lo, hi = observation_space_bounds(self.env_config, self.sim_config)
        assert lo.size == self.sim.obs_dim, (lo.shape, self.sim.obs_dim)
        self._init_vecenv(num_envs, make_box(lo, hi), make_box(*action_space_bounds(self.env_config, self.sim_config)))
        self.curriculum_level = 1.0
        self._init_host_edge(info_mode, copy_outputs)

    def _init_host_edge(self, info_mode, copy_outputs):
        """Pinned staging buffers of the numpy edge (shared by the attitude and the waypoint env)."""
        self.info_mode = info_mode
        self.copy_outputs = copy_outputs
        self.training = True
        n, adim = self.num_envs, self.action_space.shape[0]
        self._act_pin = torch.zeros(n, adim, dtype=torch.float32).pin_memory()
        self._act_np = self._act_pin.numpy()
        self._act_dev = torch.zeros(n, adim, dtype=torch.float32, device=self.device)
        # double-buffered pinned outputs: with copy_outputs=False the arrays returned by step k stay valid until
        # step k+2 (SB3's collect_rollouts reads obs_k after step k+1 returns, on_policy_algorithm.py:163-180).
        # Each buffer mirrors the simulator's packed obs | rew | done layout: ONE device-to-host copy per step.
        self._out_pin = [torch.zeros(self.sim.out_nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._out = [bt.unpack_outputs(b, n, self.sim.obs_dim) for b in self._out_pin]
        self._info_np = ([b[self.sim.info_offset:].view(torch.float64).numpy() for b in self._out_pin]
                         if self.sim.info_cap else None)
        oshape = (n,) + tuple(self.observation_space.shape)
        self._out_np = [(o.numpy().reshape(oshape), r.numpy(), d.numpy().view(np.bool_)) for o, r, d in self._out]
        self._flip = 0
        # touch the done-path once so that CUDA's lazy module loading does not land on the first finished episode
        self.sim.episode_info_rows(torch.zeros(1, dtype=torch.long, device=self.device))
        self._lazy_infos = [_EMPTY_INFO] * n
        self._lazy_dirty = []
        self._waiting = False
        self._t_start = time.time()
        self.h2d_bytes_per_step = self._act_pin.numel() * 4
        self.d2h_bytes_per_step = self.sim.out_nbytes        # obs | rew | done | episode-end rows, one copy

    def pinned_actions(self, count=1):
        """`count` float32 [num_envs, 3] arrays in page-locked host memory.  step() copies an array of this kind to the
        device directly; any other array is first staged through an internal pinned buffer (one extra host copy)."""
        bufs = [torch.zeros(self.num_envs, 3, dtype=torch.float32).pin_memory() for _ in range(count)]
        self._user_pins = getattr(self, "_user_pins", []) + bufs        # keep the allocations alive
        return [b.numpy() for b in bufs]

    # ------------------------------------------------------------------ tensor fast path
    def reset_tensor(self):
        return self.sim.reset()

    def step_tensor(self, actions):
        """actions: [n,3] float32/float64 CUDA tensor.  Returns (obs, rew, done) CUDA tensors (views of buffers that
        the next call overwrites).  Auto-resets like a VecEnv; terminal rows are in `self.sim.term_obs`."""
        return self.sim.step(actions, auto_reset=True)

    # ------------------------------------------------------------------ VecEnv API
    def reset(self):
        obs = self.sim.reset()
        self._flip ^= 1
        self._out[self._flip][0].copy_(obs, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        out = self._out_np[self._flip][0]
        return out.copy() if self.copy_outputs else out

    def step_async(self, actions):
        if self._waiting:
            raise RuntimeError("step_async called while a step is pending (subproc_vec_env.py:112 contract)")
        a = np.asarray(actions)
        assert a.shape == (self.num_envs, 3), a.shape
        src = None
        if a.dtype == np.float32 and a.flags.c_contiguous:
            t = torch.from_numpy(a)
            if t.is_pinned():                     # e.g. an array from pinned_actions(): DMA straight from it
                src = t
        if src is None:
            self._act_np[...] = a                 # casts to float32 like DummyVecEnv buffers / SB3 policies
            src, a = self._act_pin, self._act_np
        self._act_dev.copy_(src, non_blocking=True)
        self.sim.step(self._act_dev, auto_reset=True)
        self._flip ^= 1
        self._out_pin[self._flip].copy_(self.sim.out_packed, non_blocking=True)
        # fixed_wing.py:494 asserts on NaN actions before stepping; here the check runs while the GPU works (max
        # propagates NaN) — the step it rejects has been launched, but the assertion is fatal either way
        assert not np.isnan(a.max()), "NaN action (fixed_wing.py:494)"
        self._waiting = True

    def step_wait(self):
        if not self._waiting:
            raise RuntimeError("step_wait without step_async")
        torch.cuda.current_stream(self.device).synchronize()
        self._waiting = False
        obs, rew, done = self._out_np[self._flip]
        infos = self._build_infos(done)
        if self.copy_outputs:
            obs, rew, done = obs.copy(), rew.copy(), done.copy()
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.sim.close()

    def seed(self, seed=None):
        """Env i is seeded seed + i (subproc_vec_env.py:120-123): here the Philox key is `seed` and the counter
        carries the global env id, which gives every env its own stream.  Like the reference's env.seed
        (fixed_wing.py:324-332) it re-keys the random draws of every reset from now on; running episodes continue."""
        seed = 0 if seed is None else int(seed)
        self._build_kw["seed"] = seed
        self._apply_config()
        return [seed + i for i in range(self.num_envs)]

    def _apply_config(self):
        """Push the reset-time part of the configuration (curriculum-scaled init / target ranges, seed) to the live
        handle (fw_set_config): env state, device pointers and any CUDA graph captured over the step stay valid."""
        self.cfg = build_config(**self._build_kw)
        self.sim.set_config(self.cfg)

    def _indices(self, indices):
        return list(self._get_indices(indices))

    def get_attr(self, attr_name, indices=None):
        idx = self._indices(indices)
        if attr_name == "simulator":
            return [_SimulatorView(self, i) for i in idx]
        if attr_name == "cfg":                       # the env's config dict (evaluate_controller.py:123)
            return [self.env_config for _ in idx]
        if attr_name == "fw_config":
            return [self.cfg for _ in idx]
        if attr_name == "target":
            t = self._targets_host()
            return [dict(zip(self._target_names(), map(float, t[i]))) for i in idx]
        if attr_name == "steps_count":
            c = self.sim.get_field(bt.FIELD_COUNTERS).cpu().numpy()
            return [int(c[i, 0]) for i in idx]
        return [getattr(self, attr_name) for _ in idx]

    def set_attr(self, attr_name, value, indices=None):
        if attr_name not in ("training", "info_mode", "copy_outputs"):
            raise AttributeError("cannot set %r on the batched env" % attr_name)
        setattr(self, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        idx = self._indices(indices)
        if method_name == "reset":
            return self._reset_indices(idx, *args, **kwargs)
        if method_name == "set_curriculum_level":
            return self._set_curriculum_level(*args, **kwargs)
        if method_name == "seed":
            return self.seed(*args, **kwargs)
        if method_name == "render":                 # plots are out of scope: accepted, no-op
            return [None for _ in idx]
        raise NotImplementedError("env_method(%r)" % method_name)

    def _set_curriculum_level(self, level):
        """fixed_wing.py:334-412: rescales init / target ranges; applies to all envs from their next reset."""
        assert 0 <= level <= 1
        self.curriculum_level = float(level)
        self._build_kw["curriculum_level"] = float(level)
        self._apply_config()
        return [None] * self.num_envs

    def _reset_indices(self, idx, state=None, target=None, turbulence_noise=None):
        """env_method("reset", indices=i, state={...}, target={...}) as used by evaluate_controller.py:174."""
        n = self.num_envs
        mask = np.zeros(n, np.uint8)
        mask[idx] = 1
        st = np.full((n, 21), np.nan)
        tg = np.full((n, 3), np.nan)
        if state is not None:
            st[idx] = state_dict_to_row(state)
        if target is not None:
            tg[idx] = [target.get(k, np.nan) for k in TARGET_STATES]
        noise = None
        if turbulence_noise is not None:
            noise = np.zeros((n, 4, turbulence_noise.shape[1]))
            noise[idx] = turbulence_noise
        if self.sim.obs64 is None:
            self.sim.enable_f64_outputs()
        self.sim.reset(mask=mask, state=st, target=tg, noise=noise)
        obs = self.sim.obs64.cpu().numpy()
        return [obs[i].reshape(self.observation_space.shape).copy() for i in idx]

    # ------------------------------------------------------------------ info dicts
    _done_info = staticmethod(_done_info)

    def state_dict(self, i):
        y = self.sim.get_field(bt.FIELD_Y)[i].cpu().numpy()
        e = self.sim.get_field(bt.FIELD_EULER)[i].cpu().numpy()
        v = self.sim.get_field(bt.FIELD_VAB)[i].cpu().numpy()
        names = ("omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d", "velocity_u",
                 "velocity_v", "velocity_w", "elevon_right", "elevon_left", "throttle")
        d = {k: float(y[4 + j]) for j, k in enumerate(names)}
        d.update(roll=float(e[0]), pitch=float(e[1]), Va=float(v[0]), alpha=float(v[1]), beta=float(v[2]))
        return d

    def _build_infos(self, done, packed_rows=True):
        """packed_rows: the step was an auto-reset step whose outputs (and episode-end rows) are in the pinned buffer."""
        n = self.num_envs
        compat = self.info_mode == "compat"
        # lazy mode hands out ONE persistent list whose entries are a shared empty dict except for the envs that
        # finished at this step (treat as read-only; valid until the next step): copying a 65 536-entry list per step
        # costs more than the simulator
        if not compat:
            infos = self._lazy_infos
            for i in self._lazy_dirty:
                infos[i] = _EMPTY_INFO
            self._lazy_dirty = []
        # the episode-end rows came with the step outputs (fw_set_info_rows); only a step in which more than INFO_CAP
        # episodes end needs a second fetch
        rows = done_idx = None
        if packed_rows and self._info_np is not None:
            head = self._info_np[self._flip]
            cnt = int(head[:1].view(np.int32)[0])
            if cnt == 0 and not compat:
                return infos
            if cnt <= self.sim.info_cap:
                packed = head[1:1 + cnt * self.sim.info_width].reshape(cnt, self.sim.info_width)
                done_idx = packed[:, 0].astype(np.int64)
                # -> the layout of episode_info_rows: metrics | return | length | term code | terminal observation
                rows = np.concatenate([packed[:, 4:32], packed[:, 3:4], packed[:, 2:3], packed[:, 1:2], packed[:, 32:]],
                                      axis=1)
        if rows is None:
            done_idx = np.flatnonzero(done)
            if done_idx.size == 0 and not compat:
                return infos
            rows = (self.sim.episode_info_rows(torch.as_tensor(done_idx, device=self.device)) if done_idx.size
                    else np.zeros((0, 31 + self.sim.obs_dim)))
        waypoint = self.cfg.env_kind != 0
        if compat:
            tgt = self._targets_host()
            names = ("position_n", "position_e", "position_d") if waypoint else self._target_names()
            infos = [{"target": dict(zip(names, map(float, tgt[i])))} for i in range(n)]
        now = round(time.time() - self._t_start, 6)
        term_obs_all = rows[:, 31:].astype(np.float32).reshape((-1,) + tuple(self.observation_space.shape))
        generic = bool(self.cfg.obs_generic)
        for j, (row, i) in enumerate(zip(rows[:, :31].tolist(), done_idx.tolist())):
            term_obs = term_obs_all[j]
            info = self._done_info(row, term_obs, now)
            if compat:
                info.update(infos[i])                 # keeps the per-step "target" entry of compat mode
            elif not generic and not waypoint:
                # fixed_wing.py:626 reports the finished episode's target; it sits in terminal_observation[6:9]
                info["target"] = dict(zip(TARGET_STATES, term_obs[6:9].tolist()))
            infos[i] = info
        if getattr(self.cfg, "ang_on", 0) and len(done_idx):
            self._merge_angular_metrics(infos, done_idx)
        if not compat:
            self._lazy_dirty = done_idx.tolist()
        return infos

    # ---- target class attitude_angular: omega_p / omega_q / omega_r are target states too (fixed_wing.py:671-746) ----
    def _target_names(self):
        return TARGET_STATES + ANGULAR_TARGET_STATES if getattr(self.cfg, "ang_on", 0) else TARGET_STATES

    def _targets_host(self):
        t = self.sim.get_field(bt.FIELD_TARGET).cpu().numpy()
        if getattr(self.cfg, "ang_on", 0):
            t = np.concatenate([t, self.sim.get_field(bt.FIELD_ATARGET).cpu().numpy()], axis=1)
        return t

    def _merge_angular_metrics(self, infos, done_idx):
        """The per-state metric dicts of a finished episode get the entries of the rate targets (get_metric iterates every
        target state, fixed_wing.py:1644-1736; goal metrics only for states with a bound), in the reference's key order."""
        am = self.sim.episode_info_angular().index_select(
            0, torch.as_tensor(np.asarray(done_idx), dtype=torch.long, device=self.device)).cpu().numpy()
        bounded = [a for a in range(3) if np.isfinite(self.cfg.ang_bound[a])]
        err_names = ("avg_error", "total_error", "end_error", "rise_time", "overshoot")
        goal_names = ("success", "settling_time", "success_time_frac")
        for row, i in zip(am, np.asarray(done_idx).tolist()):
            info = infos[i]
            for q, name in enumerate(err_names):
                if name in info:
                    info[name].update({ANGULAR_TARGET_STATES[a]: float(row[q * 3 + a]) for a in range(3)})
            for q, name in enumerate(goal_names):
                if name in info:
                    d = info[name]
                    allv = d.pop("all", None)
                    for a in bounded:
                        v = float(row[15 + q * 3 + a])
                        d[ANGULAR_TARGET_STATES[a]] = bool(v) if name == "success" else v
                    if allv is not None:
                        d["all"] = allv


def state_dict_to_row(state):
    """Reference-style state dict (pyfly.py:1262-1294 keys) -> FW_NSTATE_INJECT row; missing entries = NaN (sampled)."""
    row = np.full(21, np.nan)
    keys = ("roll", "pitch", "yaw", "omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d",
            "velocity_u", "velocity_v", "velocity_w")
    for j, k in enumerate(keys):
        if k in state:
            row[j] = float(state[k])
    for j, k in enumerate(("elevon_right", "elevon_left", "throttle")):
        if k in state:
            v = state[k]
            try:
                row[12 + j], row[15 + j] = float(v[0]), float(v[1])
            except (TypeError, IndexError):
                row[12 + j], row[15 + j] = float(v), 0.0
    if "wind" in state and not np.isscalar(state["wind"]):
        row[18:21] = np.asarray(state["wind"], dtype=np.float64)
    elif all(k in state for k in ("wind_n", "wind_e", "wind_d")):
        row[18:21] = [float(state["wind_n"]), float(state["wind_e"]), float(state["wind_d"])]
    return row


class FixedWingAircraft:
    """Single-env gym-style API of the reference env (fixed_wing.py:13-628) on a one-env batch: reset(state, target,
    turbulence_noise) -> obs (float64 like the reference), step(action) -> (obs, reward, done, info),
    seed, set_curriculum_level, target, observation_space / action_space."""

    def __init__(self, config_path=None, sim_config_path=None, sim_parameter_path=None, config_kw=None,
                 sim_config_kw=None, device=0, **kw):
        self._v = FixedWingVecEnv(1, config_path=config_path, config_kw=config_kw, sim_config_path=sim_config_path,
                                  sim_config_kw=sim_config_kw, device=device, info_mode="compat", **kw)
        self._v.sim.enable_f64_outputs()
        self.cfg = self._v.cfg
        self.observation_space, self.action_space = self._v.observation_space, self._v.action_space
        self.steps_max = self.cfg.steps_max
        self.simulator = _SimulatorView(self._v, 0)

    @property
    def target(self):
        return self._v.get_attr("target")[0]

    @property
    def steps_count(self):
        return self._v.get_attr("steps_count")[0]

    def seed(self, seed=None):
        return self._v.seed(seed)[:1]

    def set_curriculum_level(self, level):
        self._v.env_method("set_curriculum_level", level)

    def reset(self, state=None, target=None, **sim_reset_kw):
        return self._v.env_method("reset", indices=[0], state=state, target=target,
                                  turbulence_noise=sim_reset_kw.get("turbulence_noise"))[0]

    def step(self, action):
        a = np.asarray(action)
        f64 = a.dtype != np.float32
        t = torch.as_tensor(a.reshape(1, 3), dtype=torch.float64 if f64 else torch.float32).to(self._v.device)
        self._v.sim.step(t.contiguous(), auto_reset=False)
        obs = self._v.sim.obs64.cpu().numpy()[0].reshape(self.observation_space.shape).copy()
        rew = float(self._v.sim.rew64.cpu().numpy()[0])
        done = bool(self._v.sim.done.cpu().numpy()[0])
        if done:
            self._v.sim.term_obs.copy_(self._v.sim.obs)
        info = self._v._build_infos(np.array([done]), packed_rows=False)[0]
        info.pop("terminal_observation", None)
        info.pop("episode", None)
        return obs, rew, done, info

    def close(self):
        self._v.close()


WAYPOINT_KEYS = ("position_n", "position_e", "position_d", "roll", "pitch", "yaw", "velocity_u", "velocity_v",
                 "velocity_w", "wind_n", "wind_e", "wind_d", "omega_p", "omega_q", "omega_r")


def waypoint_tasks_to_array(tasks):
    """List of tasks, each a sequence of waypoint dicts as in magpie/magpy/tasks/*/*.npy (keys position_*, roll, pitch,
    yaw, velocity_*, wind_*; optional omega_*), -> [n_tasks, wp_len, 15] float64 with NaN for missing omega."""
    wp_len = min(len(t) for t in tasks)
    out = np.full((len(tasks), wp_len, 15), np.nan)
    for i, task in enumerate(tasks):
        for j in range(wp_len):
            for k, key in enumerate(WAYPOINT_KEYS):
                if key in task[j]:
                    out[i, j, k] = float(task[j][key])
    assert not np.isnan(out[:, :, :12]).any(), "waypoints need position, attitude, velocity and wind"
    return out


class WaypointVecEnv(FixedWingVecEnv):
    """`FixedWingAircraft_simple` (magpie/magpy/simple_train.py:197-702) behind the same VecEnv surface as
    FixedWingVecEnv: every env flies the waypoint chain of its task; reaching a waypoint (0.5 m box) teleports to the next
    leg's start, reward exp(-sum |position error| / 6), 12 raw states observed, 500 steps per episode, commands passed
    straight through.  Episode-end infos carry termination / terminal_observation / episode (this env computes no
    metrics, simple_train.py:501-503)."""

    def __init__(self, num_envs, tasks, task_of_env=None, device=0, seed=0, env_id_offset=0, sim_config_kw=None,
                 config_kw=None, precision="f64", integrator="rk45", rk4_substeps=4, info_mode="lazy",
                 copy_outputs=False):
        self._build_kw = dict(env_kind="waypoint", config_kw=config_kw, sim_config_kw=sim_config_kw, seed=seed,
                              env_id_offset=env_id_offset, precision=precision, integrator=integrator,
                              rk4_substeps=rk4_substeps)
        self.cfg = build_config(**self._build_kw)
        self.env_config, self.sim_config = resolve_configs(None, None, config_kw, sim_config_kw, env_kind="waypoint")
        self.sim = bt.BatchedFixedWing(int(num_envs), cfg=self.cfg, device=device)
        self.device = self.sim.device
        self.tasks = tasks if isinstance(tasks, np.ndarray) else waypoint_tasks_to_array(tasks)
        if task_of_env is None:
            task_of_env = np.arange(int(num_envs)) % self.tasks.shape[0]
        self.task_of_env = np.asarray(task_of_env, dtype=np.int32)
        self.sim.set_waypoint_tasks(self.tasks, self.task_of_env)
        self._init_vecenv(num_envs, make_box(np.full(12, -np.inf), np.full(12, np.inf)),     # simple_train.py:267-273
                          make_box(np.array([-1, -1, 0]), np.array([1, 1, 1])))          # simple_train.py:275-279
        self.curriculum_level = 1.0
        self._init_host_edge(info_mode, copy_outputs)

    def reset_task(self, task_of_env):
        """reset_task(idx) for every env (simple_train.py:368-375): takes effect at the next reset."""
        self.task_of_env = np.asarray(task_of_env, dtype=np.int32)
        self.sim.set_waypoint_tasks(self.tasks, self.task_of_env)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        if method_name == "reset_task":
            idx = self._indices(indices)
            toe = self.task_of_env.copy()
            toe[idx] = args[0] if args else kwargs["idx"]
            self.reset_task(toe)
            return [None for _ in idx]
        if method_name in ("set_curriculum_level", "reset"):
            raise NotImplementedError("env_method(%r) on the waypoint env (simple_train.py has no such hook)" % method_name)
        return super().env_method(method_name, *args, indices=indices, **kwargs)

    def get_attr(self, attr_name, indices=None):
        if attr_name == "target":
            g = self.sim.get_field(bt.FIELD_TARGET).cpu().numpy()
            return [dict(zip(("position_n", "position_e", "position_d"), map(float, g[i]))) for i in self._indices(indices)]
        return super().get_attr(attr_name, indices)

    def _done_info(self, row, term_obs, now):
        term = int(row[30])
        return {"termination": TERM_NAMES.get(term, term), "terminal_observation": term_obs,
                "episode": {"r": row[28], "l": int(row[29]), "t": now}}

    def close(self):
        self.sim.close()
