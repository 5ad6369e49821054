"""ctypes binding of libfwb200.so — the only way host code reaches the kernels.  Fails loudly when the library
is missing or cannot be loaded: there is no CPU / PyTorch fallback for the env step."""
import ctypes
import os

from .build import LIB_PATH as _DEFAULT_LIB_PATH
from .config import FwConfig

_lib = None
_vp = ctypes.c_void_p


class FwRolloutPost(ctypes.Structure):
    """Mirror of FwRolloutPost in include/fwb200.h (field order and types must match)."""
    _fields_ = ([(k, ctypes.c_void_p) for k in (
        "obs_raw", "rew_raw", "done", "actions", "values", "log_probs", "last_obs", "last_dones", "ret", "obs_mean",
        "obs_var", "obs_count", "ret_mean", "ret_var", "ret_count", "run_ret", "run_len", "ep_stats", "buf_obs",
        "buf_actions", "buf_rewards", "buf_dones", "buf_values", "buf_log_probs", "scratch")] +
        [("n", ctypes.c_int32), ("obs_dim", ctypes.c_int32), ("act_dim", ctypes.c_int32), ("gamma", ctypes.c_float),
         ("clip_obs", ctypes.c_float), ("clip_reward", ctypes.c_float), ("epsilon", ctypes.c_float),
         ("norm_obs", ctypes.c_int32), ("norm_reward", ctypes.c_int32), ("training", ctypes.c_int32)])


class FwReplay(ctypes.Structure):
    """Mirror of FwReplay in include/fwb200.h."""
    _fields_ = [("rows", ctypes.c_void_p), ("head_dev", ctypes.c_void_p), ("size_dev", ctypes.c_void_p),
                ("sample_calls_dev", ctypes.c_void_p), ("capacity", ctypes.c_int64), ("obs_dim", ctypes.c_int32),
                ("act_dim", ctypes.c_int32), ("row_floats", ctypes.c_int32), ("_pad", ctypes.c_int32)]


class FwReplayNorm(ctypes.Structure):
    """Mirror of FwReplayNorm in include/fwb200.h."""
    _fields_ = [("obs_mean", ctypes.c_void_p), ("obs_var", ctypes.c_void_p), ("ret_var", ctypes.c_void_p),
                ("clip_obs", ctypes.c_float), ("clip_reward", ctypes.c_float), ("epsilon", ctypes.c_float),
                ("norm_obs", ctypes.c_int32), ("norm_reward", ctypes.c_int32), ("_pad", ctypes.c_int32)]


class FwError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    # FWB200_LIB: load another build of the same library (kernel A/B runs of tools/ab_bench.py)
    LIB_PATH = os.environ.get("FWB200_LIB", _DEFAULT_LIB_PATH)
    if not os.path.exists(LIB_PATH):
        raise FwError("%s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.fw_last_error.restype = ctypes.c_char_p
    L.fw_create.argtypes = [ctypes.POINTER(FwConfig), ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_vp)]
    L.fw_destroy.argtypes = [_vp]
    L.fw_set_config.argtypes = [_vp, ctypes.POINTER(FwConfig), _vp]
    L.fw_state_blob_size.argtypes = [_vp]
    L.fw_state_blob_size.restype = ctypes.c_int64
    L.fw_get_state_blob.argtypes = [_vp, _vp, _vp]
    L.fw_set_state_blob.argtypes = [_vp, _vp, _vp]
    L.fw_reset.argtypes = [_vp, _vp, _vp, _vp, _vp, ctypes.c_int32, _vp, _vp, _vp]
    L.fw_step.argtypes = [_vp, _vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int32, _vp]
    L.fw_step_random.argtypes = [_vp, ctypes.c_int32, ctypes.c_uint64, _vp, _vp, _vp, _vp]
    L.fw_get_episode_info.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
    L.fw_get_field.argtypes = [_vp, ctypes.c_int32, _vp, _vp]
    L.fw_set_field.argtypes = [_vp, ctypes.c_int32, _vp, _vp]
    L.fw_gae.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_float,
                         ctypes.c_float, _vp]
    L.fw_measure_fma_peak.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
    L.fw_debug_math.argtypes = [ctypes.c_int32, _vp, _vp, _vp, ctypes.c_int32, _vp]
    L.fw_obs_dim.argtypes = [_vp]
    L.fw_ppo_loss.argtypes = [_vp] * 7 + [ctypes.c_int32, ctypes.c_float, ctypes.c_float, ctypes.c_float] + [_vp] * 6
    L.fw_rollout_post_step.argtypes = [ctypes.POINTER(FwRolloutPost), _vp]
    L.fw_adam_clip_step.argtypes = [_vp] * 5 + [ctypes.c_int32] + [ctypes.c_float] * 5 + [_vp]
    L.fw_replay_insert.argtypes = [ctypes.POINTER(FwReplay), _vp, _vp, _vp, _vp, _vp, ctypes.c_int32, _vp]
    L.fw_replay_sample.argtypes = [ctypes.POINTER(FwReplay), ctypes.POINTER(FwReplayNorm), ctypes.c_int32, ctypes.c_uint64,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.fw_replay_size.restype = ctypes.c_int
    L.fw_replay_norm_size.restype = ctypes.c_int
    if L.fw_replay_size() != ctypes.sizeof(FwReplay) or L.fw_replay_norm_size() != ctypes.sizeof(FwReplayNorm):
        raise FwError("FwReplay / FwReplayNorm layout mismatch between _lib.py and include/fwb200.h")
    L.fw_comm_create.argtypes = [ctypes.c_int32] * 4 + [ctypes.POINTER(_vp)]
    L.fw_comm_export.argtypes = [_vp, _vp]
    L.fw_comm_connect.argtypes = [_vp, _vp]
    L.fw_comm_allreduce_adam.argtypes = [_vp] * 6 + [ctypes.c_int32] + [ctypes.c_float] * 5 + [_vp]
    L.fw_comm_error.argtypes = [_vp]
    L.fw_comm_destroy.argtypes = [_vp]
    L.fw_comm_last_error.restype = ctypes.c_char_p
    L.fw_join.argtypes = [_vp, _vp]
    L.fw_set_info_rows.argtypes = [_vp, _vp, ctypes.c_int32]
    L.fw_set_profiling.argtypes = [_vp, ctypes.c_int32]
    L.fw_get_profile.argtypes = [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]
    L.fw_set_waypoint_tasks.argtypes = [_vp, _vp, ctypes.c_int32, ctypes.c_int32, _vp, _vp]
    L.fw_config_size.restype = ctypes.c_int
    L.fw_abi_version.restype = ctypes.c_int
    if L.fw_config_size() != ctypes.sizeof(FwConfig):
        raise FwError("FwConfig layout mismatch between config.py and include/fwb200.h")
    L.fw_rollout_post_size.restype = ctypes.c_int
    if L.fw_rollout_post_size() != ctypes.sizeof(FwRolloutPost):
        raise FwError("FwRolloutPost layout mismatch between _lib.py and include/fwb200.h")
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise FwError("%s failed (%d): %s" % (what, rc, lib().fw_last_error().decode()))


EXPORTS = ("fw_create", "fw_destroy", "fw_set_config", "fw_state_blob_size", "fw_get_state_blob", "fw_set_state_blob", "fw_last_error", "fw_abi_version", "fw_reset", "fw_step", "fw_step_random", "fw_get_episode_info_angular",
           "fw_get_episode_info", "fw_get_field", "fw_set_field", "fw_gae", "fw_measure_fma_peak", "fw_debug_math",
           "fw_obs_dim", "fw_set_waypoint_tasks", "fw_set_profiling", "fw_get_profile", "fw_join", "fw_set_info_rows", "fw_ppo_loss", "fw_rollout_post_step", "fw_adam_clip_step", "fw_replay_insert", "fw_replay_sample", "fw_comm_create", "fw_comm_export", "fw_comm_connect",
           "fw_comm_allreduce_adam", "fw_comm_error", "fw_comm_destroy", "fw_comm_last_error",
           "fw_config_size", "fw_rollout_post_size", "fw_replay_size", "fw_replay_norm_size")
