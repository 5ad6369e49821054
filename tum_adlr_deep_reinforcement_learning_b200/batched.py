"""BatchedFixedWing — thin tensor-level front-end over the C ABI (one handle = n envs on one GPU).

This is the fast path: every argument and result is a torch CUDA tensor, nothing is copied to the host and no
per-env Python object is created.  `FixedWingVecEnv` (vec_env.py) layers the SB3 VecEnv contract on top.
"""
import ctypes

import torch

from . import _lib
from .config import (FW_NMETRIC, FW_NOBS, FW_NSTATE_INJECT, FW_NY, build_config)

FIELD_Y, FIELD_EULER, FIELD_VAB, FIELD_WIND, FIELD_TARGET, FIELD_CMD, FIELD_TURB, FIELD_COUNTERS, FIELD_NFEV, FIELD_PARAMS, FIELD_ATARGET = range(11)
_FIELD_SHAPE = {FIELD_Y: (FW_NY, torch.float64), FIELD_EULER: (3, torch.float64), FIELD_VAB: (3, torch.float64),
                FIELD_WIND: (3, torch.float64), FIELD_TARGET: (3, torch.float64), FIELD_CMD: (3, torch.float64),
                FIELD_TURB: (6, torch.float64), FIELD_COUNTERS: (4, torch.int32), FIELD_NFEV: (2, torch.int32),
                FIELD_PARAMS: (48, torch.float64), FIELD_ATARGET: (3, torch.float64)}


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


INFO_CAP = 64          # episode-end rows that travel with the step outputs; more than that in one step -> second fetch
FW_INFO_HEAD = 4 + FW_NMETRIC


def unpack_outputs(buf, n, obs_dim):
    """Views (obs [n, obs_dim] f32, rew [n] f32, done [n] u8) of a packed output buffer (torch uint8 tensor)."""
    o = n * obs_dim * 4
    return (buf[:o].view(torch.float32).view(n, obs_dim), buf[o:o + n * 4].view(torch.float32),
            buf[o + n * 4:o + n * 5])


class BatchedFixedWing:
    def __init__(self, n_envs, cfg=None, device=0, **cfg_kw):
        if not torch.cuda.is_available():
            raise _lib.FwError("BatchedFixedWing needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg if cfg is not None else build_config(**cfg_kw)
        self.n = int(n_envs)
        self.device = torch.device("cuda", device)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fw_create(ctypes.byref(self.cfg), self.n, device, ctypes.byref(self._h)), "fw_create")
        n, dev = self.n, self.device
        self.obs_dim = _lib.lib().fw_obs_dim(self._h)
        # obs | rew | done are views of ONE byte buffer, so that a host-facing caller fetches a step's outputs with a
        # single device-to-host copy (vec_env.FixedWingVecEnv)
        # ... followed (8-byte aligned) by the episode-end rows of fw_set_info_rows: count + INFO_CAP rows
        self.info_width = FW_INFO_HEAD + self.obs_dim
        self.info_offset = -(-(n * self.obs_dim * 4 + n * 4 + n) // 8) * 8
        self.info_cap = INFO_CAP
        self.out_nbytes = self.info_offset + (8 + self.info_cap * self.info_width * 8 if self.info_cap else 0)
        self.out_packed = torch.zeros(self.out_nbytes, dtype=torch.uint8, device=dev)
        self.obs, self.rew, self.done = unpack_outputs(self.out_packed, n, self.obs_dim)
        if self.info_cap:
            self.info_rows = self.out_packed[self.info_offset:].view(torch.float64)
            _lib.check(_lib.lib().fw_set_info_rows(self._h, _ptr(self.info_rows), self.info_cap), "fw_set_info_rows")
        self.term_obs = torch.zeros(n, self.obs_dim, dtype=torch.float32, device=dev)
        self.obs64 = None
        self.rew64 = None
        self._noise = None   # keeps an injected noise buffer alive

    def close(self):
        if self._h:
            _lib.lib().fw_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_config(self, cfg):
        """Live update of the reset-time configuration (fw_set_config): init / target ranges and the seed.  Running
        episodes, device pointers and captured CUDA graphs stay as they are; resets from now on use the new values."""
        _lib.check(_lib.lib().fw_set_config(self._h, ctypes.byref(cfg), self._stream()), "fw_set_config")
        self.cfg = cfg

    def get_state(self):
        """The whole env state of the handle as one uint8 CUDA tensor (fw_get_state_blob): SoA state, episode
        bookkeeping, precomputed next-episode rows, reset-time configuration.  `.cpu()` it and torch.save it."""
        nbytes = int(_lib.lib().fw_state_blob_size(self._h))
        blob = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib().fw_get_state_blob(self._h, _ptr(blob), self._stream()), "fw_get_state_blob")
        return blob

    def set_state(self, blob):
        """Restore a blob written by get_state() of a handle with the same configuration and n_envs."""
        blob = torch.as_tensor(blob, dtype=torch.uint8).to(self.device).contiguous()
        assert blob.numel() == int(_lib.lib().fw_state_blob_size(self._h)), "state blob of another handle shape"
        _lib.check(_lib.lib().fw_set_state_blob(self._h, _ptr(blob), self._stream()), "fw_set_state_blob")
        torch.cuda.current_stream(self.device).synchronize()

    def set_waypoint_tasks(self, tasks, task_of_env):
        """Waypoint head: tasks [n_tasks, wp_len, 15] float64 rows (position n e d, roll pitch yaw, velocity u v w, wind
        n e d, omega p q r with NaN = sample), task_of_env [n] int.  The library keeps its own device copy."""
        tasks = torch.as_tensor(tasks, dtype=torch.float64, device=self.device).contiguous()
        toe = torch.as_tensor(task_of_env, dtype=torch.int32, device=self.device).contiguous()
        assert tasks.dim() == 3 and tasks.shape[2] == 15 and toe.shape == (self.n,)
        _lib.check(_lib.lib().fw_set_waypoint_tasks(self._h, _ptr(tasks), tasks.shape[0], tasks.shape[1], _ptr(toe),
                                                    self._stream()), "fw_set_waypoint_tasks")
        torch.cuda.current_stream(self.device).synchronize()

    def enable_f64_outputs(self):
        self.obs64 = torch.zeros(self.n, self.obs_dim, dtype=torch.float64, device=self.device)
        self.rew64 = torch.zeros(self.n, dtype=torch.float64, device=self.device)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, mask=None, state=None, target=None, noise=None):
        """mask [n] uint8/bool, state [n,21] f64 (NaN = sample), target [n,3] f64, noise [n,4,L] f64 — all optional."""
        dev = self.device

        def prep(x, dtype, shape):
            if x is None:
                return None
            x = torch.as_tensor(x, dtype=dtype, device=dev).contiguous()
            assert tuple(x.shape) == shape, (tuple(x.shape), shape)
            return x

        mask = prep(mask, torch.uint8, (self.n,))
        state = prep(state, torch.float64, (self.n, FW_NSTATE_INJECT))
        target = prep(target, torch.float64, (self.n, 3))
        nlen = 0
        if noise is not None:
            noise = torch.as_tensor(noise, dtype=torch.float64, device=dev).contiguous()
            assert noise.shape[:2] == (self.n, 4)
            nlen = noise.shape[2]
        self._noise = noise                      # the kernels keep reading it until the next reset: keep it alive
        self._reset_args = (mask, state, target)   # the reset kernel is asynchronous: keep its inputs alive too
        _lib.check(_lib.lib().fw_reset(self._h, _ptr(mask), _ptr(state), _ptr(target), _ptr(noise), nlen,
                                       _ptr(self.obs), _ptr(self.obs64), self._stream()), "fw_reset")
        return self.obs

    def step(self, actions, auto_reset=True):
        """actions [n,3] float32 or float64 CUDA tensor (raw agent actions).  Returns (obs, rew, done) tensors that
        are overwritten by the next call."""
        assert actions.is_cuda and actions.is_contiguous() and tuple(actions.shape) == (self.n, 3)
        f64 = actions.dtype == torch.float64
        assert f64 or actions.dtype == torch.float32
        _lib.check(_lib.lib().fw_step(self._h, _ptr(actions), int(f64), _ptr(self.obs), _ptr(self.rew),
                                      _ptr(self.done), _ptr(self.term_obs), _ptr(self.obs64), _ptr(self.rew64),
                                      int(auto_reset), self._stream()), "fw_step")
        return self.obs, self.rew, self.done

    def step_random(self, k_steps, action_seed=1):
        _lib.check(_lib.lib().fw_step_random(self._h, int(k_steps), int(action_seed), _ptr(self.obs), _ptr(self.rew),
                                             _ptr(self.done), self._stream()), "fw_step_random")
        return self.obs, self.rew, self.done

    def join(self):
        """Make the current stream wait for the side-stream refill of precomputed reset rows (fw_join): needed before
        a CUDA graph capture that contains step() calls ends."""
        _lib.check(_lib.lib().fw_join(self._h, self._stream()), "fw_join")

    def set_profiling(self, on=True):
        """Per-kernel CUDA-event timing of every following step (fw_set_profiling); clears the sums."""
        _lib.check(_lib.lib().fw_set_profiling(self._h, int(bool(on))), "fw_set_profiling")

    def profile(self):
        """{"init_ms", "integrate_ms", "head_ms"}: mean duration per step of each kernel since set_profiling."""
        import ctypes
        ms = (ctypes.c_double * 3)()
        steps = ctypes.c_int64()
        _lib.check(_lib.lib().fw_get_profile(self._h, ms, ctypes.byref(steps)), "fw_get_profile")
        k = max(int(steps.value), 1)
        return {"init_ms": ms[0] / k, "integrate_ms": ms[1] / k, "head_ms": ms[2] / k, "steps": int(steps.value)}

    def episode_info(self):
        """(term_code [n] i32, metrics [n,28] f64, episode return [n] f64, episode length [n] i32) device tensors,
        valid for the envs whose done flag was set by the last step.  The tensors are reused by the next call."""
        if getattr(self, "_ep_bufs", None) is None:
            dev = self.device
            self._ep_bufs = (torch.zeros(self.n, dtype=torch.int32, device=dev),
                             torch.zeros(self.n, FW_NMETRIC, dtype=torch.float64, device=dev),
                             torch.zeros(self.n, dtype=torch.float64, device=dev),
                             torch.zeros(self.n, dtype=torch.int32, device=dev))
        term, metrics, ret, length = self._ep_bufs
        _lib.check(_lib.lib().fw_get_episode_info(self._h, _ptr(term), _ptr(metrics), _ptr(ret), _ptr(length),
                                                  self._stream()), "fw_get_episode_info")
        return term, metrics, ret, length

    def episode_info_angular(self):
        """attitude_angular configs: [n, 24] float64 device tensor of the omega_p/q/r metrics (avg_error, total_error,
        end_error, rise_time, overshoot, success, settling_time, success_time_frac x 3, metric-major) of the envs whose
        done flag was set by the last step."""
        if getattr(self, "_ang_buf", None) is None:
            self._ang_buf = torch.zeros(self.n, 24, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().fw_get_episode_info_angular(self._h, _ptr(self._ang_buf), self._stream()),
                   "fw_get_episode_info_angular")
        return self._ang_buf

    def episode_info_rows(self, idx):
        """Rows `idx` (LongTensor on the device) of (metrics | return | length | term_code | terminal observation) as
        ONE packed host array [k, 28 + 3 + obs_dim] — a single small D2H copy instead of four full-size ones."""
        term, metrics, ret, length = self.episode_info()
        packed = torch.cat([metrics.index_select(0, idx), ret.index_select(0, idx)[:, None],
                            length.index_select(0, idx)[:, None].to(torch.float64),
                            term.index_select(0, idx)[:, None].to(torch.float64),
                            self.term_obs.index_select(0, idx).to(torch.float64)], dim=1)
        return packed.cpu().numpy()

    def get_field(self, field):
        width, dtype = _FIELD_SHAPE[field]
        out = torch.zeros(self.n, width, dtype=dtype, device=self.device)
        _lib.check(_lib.lib().fw_get_field(self._h, field, _ptr(out), self._stream()), "fw_get_field")
        return out

    def set_field(self, field, value):
        width, dtype = _FIELD_SHAPE[field]
        value = torch.as_tensor(value, dtype=dtype, device=self.device).contiguous()
        assert tuple(value.shape) == (self.n, width)
        _lib.check(_lib.lib().fw_set_field(self._h, field, _ptr(value), self._stream()), "fw_set_field")
        torch.cuda.current_stream(self.device).synchronize()


def gae(rew, val, done, last_val, last_done, gamma=0.99, gae_lambda=0.95):
    """Time-major [T, N] float32 CUDA tensors -> (advantages, returns); RolloutBuffer semantics (buffers.py:304-333)."""
    T, N = rew.shape
    for t in (rew, val, done):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (T, N)
    last_val = last_val.to(torch.float32).contiguous().view(-1)
    last_done = last_done.to(torch.uint8).contiguous().view(-1)
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    st = ctypes.c_void_p(torch.cuda.current_stream(rew.device).cuda_stream)
    _lib.check(_lib.lib().fw_gae(_ptr(rew), _ptr(val), _ptr(done), _ptr(last_val), _ptr(last_done), _ptr(adv),
                                 _ptr(ret), T, N, float(gamma), float(gae_lambda), st), "fw_gae")
    return adv, ret


def measure_fma_peak(device=0, precision="f64"):
    out = ctypes.c_double()
    _lib.check(_lib.lib().fw_measure_fma_peak(device, 0 if precision == "f64" else 1, ctypes.byref(out)),
               "fw_measure_fma_peak")
    return out.value


def debug_math(op, x, y=None):
    """Elementwise evaluation of the RHS hot-loop math (0 exp, 1 asin, 2 atan2(y, x), 3 rsqrt, 4..6 immediate-coefficient builds of 0..2, 7 log, 8 / 9 x^y) on float64 CUDA tensors."""
    x = x.contiguous()
    out = torch.empty_like(x)
    st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(_lib.lib().fw_debug_math(int(op), _ptr(x), _ptr(None if y is None else y.contiguous()), _ptr(out),
                                        x.numel(), st), "fw_debug_math")
    return out
