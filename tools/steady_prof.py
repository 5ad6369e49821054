"""260 steps of the C3 workload with random actions (steady state: episodes end every step) for an ncu capture."""
import sys; sys.path.insert(0, ".")
import torch
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
n = 65536
env = bt.BatchedFixedWing(n, cfg=build_config(sim_config_kw={"turbulence": True})); env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
pool = [(torch.rand(n, 3, device="cuda", generator=g) * 2 - 1).contiguous() for _ in range(8)]
nd = 0
for i in range(260):
    env.step(pool[i % 8])
    if i >= 250: nd += int(env.done.sum())
torch.cuda.synchronize()
print("dones in the last 10 steps:", nd)
