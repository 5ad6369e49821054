"""ctypes front-end of the C oracle (oracle/fw_oracle.c).  TEST INFRASTRUCTURE — see the header of that file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

from tum_adlr_deep_reinforcement_learning_b200.config import FW_NMETRIC, FW_NOBS, FW_NY, FwConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfworacle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    src = os.path.join(_HERE, "fw_oracle.c")
    hdr = os.path.join(os.path.dirname(_HERE), "include", "fwb200.h")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src),
                                                                                   os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfworacle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.fwo_create.restype = ctypes.c_void_p
        L.fwo_create.argtypes = [ctypes.POINTER(FwConfig), ctypes.c_int64]
        L.fwo_destroy.argtypes = [ctypes.c_void_p]
        L.fwo_reset.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, ctypes.c_int, _dp]
        L.fwo_step.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, _dp, _dp, ctypes.POINTER(ctypes.c_int),
                               ctypes.POINTER(ctypes.c_int)]
        L.fwo_get.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _i32p]
        L.fwo_get_metrics.argtypes = [ctypes.c_void_p, _dp, _dp, _i32p, _i32p]
        L.fwo_get_angular.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.fwo_turbulence.restype = _dp
        L.fwo_turbulence.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        L.fwo_rhs.argtypes = [ctypes.POINTER(FwConfig), _dp, _dp, _dp, _dp, _dp]
        L.fwo_dryden.argtypes = [ctypes.POINTER(FwConfig), _dp, ctypes.c_int, _dp]
        L.fwo_gae.argtypes = [_fp, _fp, _fp, _fp, _u8p, _fp, _fp, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                              ctypes.c_double]
        L.fwo_philox4x32.argtypes = [ctypes.POINTER(ctypes.c_uint32)] * 3
        L.fwo_noise4.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32, _dp]
        L.fwo_batch_create.restype = ctypes.c_void_p
        L.fwo_batch_create.argtypes = [ctypes.POINTER(FwConfig), ctypes.c_int]
        L.fwo_batch_destroy.argtypes = [ctypes.c_void_p]
        L.fwo_batch_set_config.argtypes = [ctypes.c_void_p, ctypes.POINTER(FwConfig)]
        L.fwo_set_config.argtypes = [ctypes.c_void_p, ctypes.POINTER(FwConfig)]
        L.fwo_set_params.argtypes = [ctypes.c_void_p, _dp]
        L.fwo_get_params.argtypes = [ctypes.c_void_p, _dp]
        L.fwo_batch_env.restype = ctypes.c_void_p
        L.fwo_batch_env.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.fwo_batch_reset.argtypes = [ctypes.c_void_p, _dp]
        L.fwo_batch_step.argtypes = [ctypes.c_void_p, _fp, _dp, _dp, _u8p]
        L.fwo_batch_step_random.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, _dp,
                                            _dp, _u8p]
        L.fwo_set_waypoint_tasks.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.fwo_batch_set_waypoint_tasks.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, ctypes.c_int, _i32p]
        L.fwo_batch_counters.argtypes = [ctypes.c_void_p, _i32p, _i32p]
        L.fwo_random_action.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint64, _fp]
        L.fwo_config_size.restype = ctypes.c_int
        L.fwo_obs_dim.argtypes = [ctypes.POINTER(FwConfig)]
        _lib = L
    return _lib


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


class OracleEnv:
    """One reference-semantics env (FixedWingAircraft over PyFly) evaluated by the C restatement."""

    def __init__(self, cfg, env_id=0):
        self.cfg = cfg
        self.obs_dim = lib().fwo_obs_dim(ctypes.byref(cfg))
        self._h = lib().fwo_create(ctypes.byref(cfg), env_id)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fwo_destroy(self._h)
            self._h = None

    def set_waypoint_tasks(self, tasks, task=0):
        """tasks: [n_tasks, wp_len, 15] float64 (FW_WP_ROW layout); the array is kept alive by this object."""
        self._tasks = np.ascontiguousarray(tasks, dtype=np.float64)
        lib().fwo_set_waypoint_tasks(self._h, _p(self._tasks), self._tasks.shape[0], self._tasks.shape[1], int(task))

    def reset(self, state=None, target=None, noise=None):
        obs = np.zeros(self.obs_dim)
        st = None if state is None else np.ascontiguousarray(state, dtype=np.float64)
        tg = None if target is None else np.ascontiguousarray(target, dtype=np.float64)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        lib().fwo_reset(self._h, _p(st), _p(tg), _p(nz), 0 if nz is None else nz.shape[1], _p(obs))
        return obs

    def step(self, action, f32=False):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs = np.zeros(self.obs_dim)
        rew = ctypes.c_double()
        done, term = ctypes.c_int(), ctypes.c_int()
        lib().fwo_step(self._h, _p(a), int(f32), _p(obs), ctypes.byref(rew), ctypes.byref(done), ctypes.byref(term))
        return obs, rew.value, bool(done.value), term.value

    def angular(self):
        """(rate targets omega_p/q/r, the 24 attitude_angular metrics of the last finished episode)."""
        at, am = np.zeros(3), np.zeros(24)
        lib().fwo_get_angular(self._h, _p(at), _p(am))
        return at, am

    def get(self):
        y, eu, vab, cmd, tgt, wind = (np.zeros(FW_NY), np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3),
                                      np.zeros(3))
        cnt = np.zeros(6, dtype=np.int32)
        lib().fwo_get(self._h, _p(y), _p(eu), _p(vab), _p(cmd), _p(tgt), _p(wind), _p(cnt, _i32p))
        return dict(y=y, euler=eu, vab=vab, cmd=cmd, target=tgt, wind=wind, steps_count=int(cnt[0]),
                    steps_for_target=int(cnt[1]), sim_step=int(cnt[2]), episode=int(cnt[3]), nfev=int(cnt[4]),
                    natt=int(cnt[5]))

    def set_params(self, par48):
        """Aircraft parameters (mass .. C_n_delta_r, FwConfig order) of the running episode."""
        lib().fwo_set_params(self._h, _p(np.ascontiguousarray(par48, dtype=np.float64)))

    def params(self):
        out = np.zeros(48)
        lib().fwo_get_params(self._h, _p(out))
        return out

    def metrics(self):
        m = np.zeros(FW_NMETRIC)
        ret = ctypes.c_double()
        ln, term = ctypes.c_int32(), ctypes.c_int32()
        lib().fwo_get_metrics(self._h, _p(m), ctypes.byref(ret), ctypes.byref(ln), ctypes.byref(term))
        return m, ret.value, ln.value, term.value

    def turbulence(self):
        n = ctypes.c_int()
        ptr = lib().fwo_turbulence(self._h, ctypes.byref(n))
        return np.ctypeslib.as_array(ptr, shape=(6, n.value)).copy()


def rhs(cfg, y, cmd_dyn, wind, turb6):
    dy = np.zeros(FW_NY)
    rc = lib().fwo_rhs(ctypes.byref(cfg), _p(np.ascontiguousarray(y, dtype=np.float64)),
                       _p(np.ascontiguousarray(cmd_dyn, dtype=np.float64)),
                       _p(np.ascontiguousarray(wind, dtype=np.float64)),
                       _p(np.ascontiguousarray(turb6, dtype=np.float64)), _p(dy))
    return dy, rc


def dryden(cfg, noise):
    noise = np.ascontiguousarray(noise, dtype=np.float64)
    out = np.zeros((6, noise.shape[1]))
    lib().fwo_dryden(ctypes.byref(cfg), _p(noise), noise.shape[1], _p(out))
    return out


def gae(rew, val, done, last_val, last_done, gamma=0.99, lam=0.95):
    T, N = rew.shape
    rew, val, done = (np.ascontiguousarray(x, dtype=np.float32) for x in (rew, val, done))
    last_val = np.ascontiguousarray(last_val, dtype=np.float32)
    last_done = np.ascontiguousarray(last_done, dtype=np.uint8)
    adv, ret = np.zeros((T, N), np.float32), np.zeros((T, N), np.float32)
    lib().fwo_gae(_p(rew, _fp), _p(val, _fp), _p(done, _fp), _p(last_val, _fp), _p(last_done, _u8p), _p(adv, _fp),
                  _p(ret, _fp), T, N, gamma, lam)
    return adv, ret


def philox4x32(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().fwo_philox4x32(c, k, o)
    return list(o)


def noise4(seed, env_id, episode, k):
    out = np.zeros(4)
    lib().fwo_noise4(seed, env_id, episode, k, _p(out))
    return out


def random_action(seed, env_id, step):
    a = np.zeros(3, np.float32)
    lib().fwo_random_action(seed, env_id, step, _p(a, _fp))
    return a


class OracleBatch:
    """n oracle envs with the VecEnv auto-reset contract; used as the CPU baseline (threads over sub-batches)."""

    def __init__(self, cfg, n):
        self.n = n
        self._h = lib().fwo_batch_create(ctypes.byref(cfg), n)
        self.obs = np.zeros((n, lib().fwo_obs_dim(ctypes.byref(cfg))))
        self.rew = np.zeros(n)
        self.done = np.zeros(n, np.uint8)

    def __del__(self):
        try:                                   # module globals may be gone at interpreter shutdown
            if getattr(self, "_h", None):
                lib().fwo_batch_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_waypoint_tasks(self, tasks, task_of_env):
        self._tasks = np.ascontiguousarray(tasks, dtype=np.float64)
        toe = np.ascontiguousarray(task_of_env, dtype=np.int32)
        lib().fwo_batch_set_waypoint_tasks(self._h, _p(self._tasks), self._tasks.shape[0], self._tasks.shape[1],
                                           _p(toe, _i32p))

    def set_config(self, cfg):
        """Live change of the reset-time configuration (the oracle's fw_set_config)."""
        lib().fwo_batch_set_config(self._h, ctypes.byref(cfg))

    def reset(self):
        lib().fwo_batch_reset(self._h, _p(self.obs))
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        lib().fwo_batch_step(self._h, _p(a, _fp), _p(self.obs), _p(self.rew), _p(self.done, _u8p))
        return self.obs, self.rew, self.done

    def env_angular(self, i):
        """(rate targets, the 24 attitude_angular metrics of the last finished episode) of env i."""
        at, am = np.zeros(3), np.zeros(24)
        lib().fwo_get_angular(lib().fwo_batch_env(self._h, i), _p(at), _p(am))
        return at, am

    def params(self):
        out = np.zeros((self.n, 48))
        for i in range(self.n):
            lib().fwo_get_params(lib().fwo_batch_env(self._h, i), _p(out[i]))
        return out

    def counters(self):
        nfev, natt = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32)
        lib().fwo_batch_counters(self._h, _p(nfev, _i32p), _p(natt, _i32p))
        return nfev, natt

    def step_random(self, k, seed, step0):
        lib().fwo_batch_step_random(self._h, k, seed, step0, _p(self.obs), _p(self.rew), _p(self.done, _u8p))
        return self.obs, self.rew, self.done
