"""e2e VecEnv.step in steady state: wall time per step, kernel times (profiling API), host-side split."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
n = 65536
v = FixedWingVecEnv(n, sim_config_kw={"turbulence": True}, seed=0)
v.reset()
rs = np.random.RandomState(0)
pool = [rs.uniform(-1, 1, (n, 3)).astype(np.float32) for _ in range(8)]
for i in range(250): v.step(pool[i % 8])
for label, prof in (("plain", False), ("profiled", True)):
    v.sim.set_profiling(prof)
    t_async = t_wait = 0.0; nd = 0
    t0 = time.perf_counter()
    for i in range(100):
        a = time.perf_counter(); v.step_async(pool[i % 8]); b = time.perf_counter()
        o, r, d, info = v.step_wait(); c = time.perf_counter()
        t_async += b - a; t_wait += c - b; nd += int(d.sum())
    dt = (time.perf_counter() - t0) / 100
    print(label, "step %.0f us  async %.0f  wait %.0f  dones/step %.1f" % (dt * 1e6, t_async * 1e4, t_wait * 1e4, nd / 100),
          v.sim.profile() if prof else "")
# the same without info building
v.sim.set_profiling(False)
v.info_mode = "none"
