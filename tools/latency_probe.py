"""Per-attempt-iteration latency of the RK45 attempt kernel: one warp alone vs a full machine."""
import sys; sys.path.insert(0, ".")
import torch
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
for n in (32, 1024, 4736, 37888, 65536):
    env = bt.BatchedFixedWing(n, cfg=build_config(sim_config_kw={"turbulence": True})); env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    pool = [(torch.rand(n, 3, device="cuda", generator=g) * 2 - 1).contiguous() for _ in range(8)]
    for i in range(30): env.step(pool[i % 8])
    env.set_profiling(True)
    mx = 0; tot = 0
    for i in range(50):
        env.step(pool[i % 8])
        natt = env.get_field(bt.FIELD_NFEV)[:, 1]
        mx += int(natt.max()); tot += float(natt.float().mean())
    p = env.profile()
    print("n=%6d integrate %.1f us  init %.1f  head %.1f | mean max-attempts %.2f mean attempts %.2f -> %.1f us per max-attempt" % (
        n, p["integrate_ms"] * 1e3, p["init_ms"] * 1e3, p["head_ms"] * 1e3, mx / 50, tot / 50, p["integrate_ms"] * 1e3 / (mx / 50)))
    env.close()
