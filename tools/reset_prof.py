"""Times env.reset() (main + spare rows) and a burst of steps, to size the reset critical path."""
import sys, json; sys.path.insert(0, ".")
import torch
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
for n in (65536, 32):
    env = bt.BatchedFixedWing(n, cfg=build_config(sim_config_kw={"turbulence": True})); env.reset()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): env.reset()
    e1.record(); torch.cuda.synchronize()
    print("reset (main + spare) n=%d: %.1f us" % (n, e0.elapsed_time(e1) * 100))
    env.close()
