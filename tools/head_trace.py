"""Per-step kernel durations over the first 400 steps (profiling API) next to the number of finished episodes."""
import sys, ctypes; sys.path.insert(0, ".")
import torch
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt, _lib
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
n = 65536
env = bt.BatchedFixedWing(n, cfg=build_config(sim_config_kw={"turbulence": True})); env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
pool = [(torch.rand(n, 3, device="cuda", generator=g) * 2 - 1).contiguous() for _ in range(8)]
for i in range(3): env.step(pool[i % 8])
env.set_profiling(True)
prev = (0.0, 0.0, 0.0)
for i in range(3, 400):
    env.step(pool[i % 8])
    ms = (ctypes.c_double * 3)(); st = ctypes.c_int64()
    _lib.lib().fw_get_profile(env._h, ms, ctypes.byref(st))
    cur = (ms[0], ms[1], ms[2])
    if i < 12 or i % 10 == 0:
        nf = env.get_field(bt.FIELD_NFEV).float().mean(0).tolist()
        print("step %3d init %5.1f integ %6.1f head %6.1f us  dones %4d  nfev %.1f" % (
            i, (cur[0] - prev[0]) * 1e3, (cur[1] - prev[1]) * 1e3, (cur[2] - prev[2]) * 1e3, int(env.done.sum()), nf[0]))
    prev = cur
