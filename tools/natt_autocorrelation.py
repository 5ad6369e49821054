"""CPU experiment (oracle only, no GPU): is the RK45 attempt count of an env at step t predictable from its count at
step t-1?  If it were, head_kernel could order the attempt kernel's queue so that likely stragglers start first.
Result on the C3 workload (4096 envs x 120 steps): correlation 0.087; of the envs that need >= 7 attempts, 12 % had
>= 5 attempts the step before against a base rate of 15 % -- no signal, no ordering built."""
import sys, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
from oracle import fw_oracle as O
cfg = build_config(sim_config_kw={"turbulence": True}, precision="f64", integrator="rk45", seed=0)
n, T = 4096, 120
b = O.OracleBatch(cfg, n); b.reset()
A = np.zeros((T, n), np.int32); D = np.zeros((T, n), np.uint8)
t0 = time.time()
for t in range(T):
    b.step_random(1, 1, t)
    A[t] = b.counters()[1]; D[t] = b.done
print("time", time.time() - t0, "mean natt", A.mean(), "max per step mean", A.max(1).mean())
cur, prev = A[1:].ravel(), A[:-1].ravel()
print("hist", np.bincount(cur))
for thr in (6, 7, 8):
    pos = cur >= thr
    print("thr", thr, "P=", pos.mean())
    for pthr in (4, 5, 6):
        sel = prev >= pthr
        print("   prev>=%d: frac of envs %.3f, recall %.3f, precision %.4f" % (pthr, sel.mean(), (pos & sel).sum() / max(pos.sum(), 1), (pos & sel).sum() / max(sel.sum(), 1)))
print("corr", np.corrcoef(cur, prev)[0, 1])
