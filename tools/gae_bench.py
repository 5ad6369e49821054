"""GAE kernel bandwidth at the SB3-default rollout size (T = 2048, N = 8192): python tools/gae_bench.py [T N]
FWB200_GAE_SW=8|16|32 forces the strip width, 1 the plain one-thread-per-column kernel."""
import sys
import torch
sys.path.insert(0, ".")
from tum_adlr_deep_reinforcement_learning_b200.batched import gae

T, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 8192)
g = torch.Generator(device="cuda").manual_seed(0)
rew, val = (torch.randn(T, N, device="cuda", generator=g) for _ in range(2))
done = (torch.rand(T, N, device="cuda", generator=g) < 0.001).float()
lv, ld = torch.randn(N, device="cuda", generator=g), (torch.rand(N, device="cuda", generator=g) < 0.1).to(torch.uint8)
for _ in range(3):
    gae(rew, val, done, lv, ld)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    a.record(); gae(rew, val, done, lv, ld); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
print("T %d N %d: median %.1f us, best %.1f us -> %.0f GB/s (20 B per transition)" % (T, N, ms[10] * 1e3, ms[0] * 1e3, 20.0 * T * N / (ms[10] * 1e-3) / 1e9))
