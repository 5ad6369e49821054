"""Import shim for the *live* reference (fixture generation only; never shipped, never on the product path).

TEST INFRASTRUCTURE.  The reference (`/root/reference`, pure Python) imports `gym` and `matplotlib`,
neither of which is installed in this image.  This module fabricates just enough of both for
`pyfly`, `gym_fixed_wing` and the stable-baselines3 fork to import and run unmodified from where they
lie, so that `tests/golden/make_golden.py` can record golden input/output vectors.

`/root/reference` does not exist on the GPU box.  There the same unmodified libraries are found in the mirror
`baseline/_ref/` (git-ignored; written by `baseline/install_ref.py`, travels with the gpurun snapshot), which lets
`bench.py`'s reference legs time the Python reference on the box's host cores and the drop-in tests run the fork's own
PPO on `FixedWingVecEnv`.  The product package never imports this file.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

_MIRROR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("FW_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "magpie", "libs")) and os.path.isdir(os.path.join(_MIRROR, "magpie", "libs")):
    REFERENCE_ROOT = _MIRROR
_LIBS = os.path.join(REFERENCE_ROOT, "magpie", "libs")


def available():
    """True when the reference libraries can be imported (mounted tree or the baseline/_ref mirror)."""
    return os.path.isdir(_LIBS)


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)

    def seed(self, seed=None):                      # gym.Space.seed (BaseAlgorithm.set_random_seed calls it)
        self._rng = np.random.RandomState(seed)
        return [seed]


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low, high = np.asarray(low), np.asarray(high)
        super().__init__(low.shape if shape is None else shape, dtype)
        self.low = np.broadcast_to(low, self.shape).astype(dtype)
        self.high = np.broadcast_to(high, self.shape).astype(dtype)

    def sample(self):
        return np.random.uniform(-1, 1, self.shape).astype(self.dtype)

    def contains(self, x):
        return True


class Env:
    metadata = {}

    def close(self):
        pass


class GoalEnv(Env):
    pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env


class _Fabricated(types.ModuleType):
    """Module whose missing attributes are fabricated: lower-case -> sub-module, Capitalised -> empty class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[0].isupper():
            return type(name, (), {})
        sub = _Fabricated(self.__name__ + "." + name)
        sub.__path__ = []
        sys.modules[sub.__name__] = sub
        setattr(self, name, sub)
        return sub


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    roots = ("gym", "matplotlib", "mpl_toolkits")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _Fabricated(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, mod):
        if mod.__name__ == "gym":
            mod.Env, mod.GoalEnv, mod.Wrapper = Env, GoalEnv, Wrapper
        elif mod.__name__ == "gym.spaces":
            mod.Space, mod.Box = Space, Box
        elif mod.__name__ == "gym.spaces.utils":
            mod.flatdim = lambda space: int(np.prod(space.shape))
        elif mod.__name__ == "gym.utils.seeding":
            mod.np_random = lambda seed=None: (np.random.RandomState(seed), seed)


def install():
    """Make `import pyfly`, `import gym_fixed_wing`, `import stable_baselines3` resolve to the reference."""
    if not os.path.isdir(_LIBS):
        raise RuntimeError("reference tree not found at %s (fixtures can only be generated where it is mounted)"
                           % REFERENCE_ROOT)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
        for sub in ("pyfly", "fixed-wing-gym", "stable-baselines3"):
            sys.path.insert(0, os.path.join(_LIBS, sub))
        import gym  # noqa: F401
        import gym.spaces  # noqa: F401
        import gym.spaces.utils  # noqa: F401
        import gym.utils.seeding  # noqa: F401


GYM_CONFIG = os.path.join(_LIBS, "fixed-wing-gym", "gym_fixed_wing", "fixed_wing_config.json")
PID_TEST_SET = os.path.join(_LIBS, "fixed-wing-gym", "gym_fixed_wing", "examples", "test_sets",
                            "test_set_wind_none_step20-20-3.npy")
PID_GOLDEN = os.path.join(_LIBS, "fixed-wing-gym", "gym_fixed_wing", "examples", "evaluations",
                          "eval_res_PID_none.npy")
X8_PARAMS = os.path.join(_LIBS, "pyfly", "pyfly", "x8_param.mat")
