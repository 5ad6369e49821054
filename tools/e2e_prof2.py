import sys, time, cProfile, pstats; sys.path.insert(0, '.')
import numpy as np, torch
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
n = 65536
v = FixedWingVecEnv(n, sim_config_kw={"turbulence": True}, seed=0)
v.reset()
a = np.random.uniform(-1, 1, (n, 3)).astype(np.float32)
nd = []
import sys
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10): v.step(a)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for _ in range(50):
    o, r, d, i = v.step(a); nd.append(int(d.sum()))
dt = (time.perf_counter() - t0) / 50
pr.disable()
print("step %.0f us, dones per step: %s" % (dt * 1e6, nd[:20]))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
