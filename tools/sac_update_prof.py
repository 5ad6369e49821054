"""Kernel list of one SAC gradient step (eager) via torch.profiler."""
import sys; sys.path.insert(0, ".")
import collections, torch
from torch.profiler import profile, ProfilerActivity
from tum_adlr_deep_reinforcement_learning_b200.sac import SAC
from tum_adlr_deep_reinforcement_learning_b200.vec_env import FixedWingVecEnv
env = FixedWingVecEnv(1024, sim_config_kw={"turbulence": True}, seed=0)
algo = SAC(env, buffer_size=200_000, batch_size=4096, gradient_steps=2, learning_starts=8192, use_cuda_graph=False)
algo.learn(total_timesteps=20 * 1024)
for _ in range(3): algo.train_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    algo.train_step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
print("kernels", len(ev), "total device us", sum(e.device_time for e in ev))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e.name[:84]][0] += 1; agg[e.name[:84]][1] += e.device_time
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%4d %8.1f us  %s" % (c, t, k))
