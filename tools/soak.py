"""Soak: 65 536 envs through the steps_max boundary (tens of thousands of episodes end in ONE step: precomputed-row
list, refill kernel and the VecEnv's second-fetch path at their maximum), then 3000 more random-action steps."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
n = 65536
v = FixedWingVecEnv(n, config_kw={"steps_max": 300}, sim_config_kw={"turbulence": True}, seed=1)
obs = v.reset()
a0 = np.zeros((n, 3), np.float32); a0[:, 2] = 0.2
ends = 0
for t in range(305):
    obs, rew, done, infos = v.step(a0)
    k = int(done.sum()); ends += k
    if k > 1000:
        assert all(("episode" in infos[i]) for i in np.flatnonzero(done)[:200])
        print("step", t + 1, "episodes ended:", k, "mean length", np.mean([infos[i]["episode"]["l"] for i in np.flatnonzero(done)[:2000]]))
assert np.isfinite(obs).all()
g = torch.Generator(device="cuda"); g.manual_seed(0)
t0 = time.time()
for t in range(3000):
    a = (torch.rand(n, 3, device="cuda", generator=g) * 2 - 1).contiguous()
    o, r, d = v.step_tensor(a)
    ends += int(d.sum()) if t % 100 == 0 else 0
torch.cuda.synchronize()
assert bool(torch.isfinite(o).all()) and bool(torch.isfinite(r).all())
print("3000 device steps ok in %.1f s (%.3e env-steps/s incl. host RNG)" % (time.time() - t0, 3000 * n / (time.time() - t0)))
