import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from conftest import angular_env_config, integrator_env_config
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
for name, ecfg in (("angular", angular_env_config()), ("integrator", integrator_env_config(4, 3, 2))):
    outs = {}
    for prec in ("f64", "f32"):
        cfg = build_config(env_cfg=ecfg, sim_config_kw={"turbulence": True}, seed=3, precision=prec)
        env = bt.BatchedFixedWing(256, cfg=cfg); env.reset()
        rs = np.random.RandomState(0); O = []; R = []
        for t in range(40):
            a = torch.as_tensor(rs.uniform(-1, 1, (256, 3)).astype(np.float32)).cuda()
            o, r, d = env.step(a)
            O.append(o.cpu().numpy().copy()); R.append(r.cpu().numpy().copy())
        outs[prec] = (np.array(O), np.array(R)); env.close()
    do = np.abs(outs["f64"][0] - outs["f32"][0]); dr = np.abs(outs["f64"][1] - outs["f32"][1])
    print(name, "finite", np.isfinite(outs["f32"][0]).all(), "median obs diff %.2e  p99 %.2e | reward median %.2e p99 %.2e" % (
        np.median(do), np.percentile(do, 99), np.median(dr), np.percentile(dr, 99)))
