"""CPU only, build container only (needs /root/reference): a randomised sweep of the oracle against the LIVE reference
beyond the committed fixtures.  Fresh seeds, every turbulence intensity, calm / windy / hard initial states, large
action amplitudes; every step compares what tests/test_oracle_golden.py compares (done / termination / RK45 RHS count
exact, state / observation / reward <= 1e-9 relative).  Prints the worst deviation per category and every mismatch.

    python tools/oracle_sweep.py [episodes_per_category] [steps] [seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as G  # noqa: E402  (installs oracle/refshim and imports the reference)
from oracle import fw_oracle as O  # noqa: E402
from tum_adlr_deep_reinforcement_learning_b200.config import build_config  # noqa: E402

N_EP = int(sys.argv[1]) if len(sys.argv) > 1 else 12
N_STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
SEED = int(sys.argv[3]) if len(sys.argv) > 3 else 20260101

CASES = [  # name, turbulence, intensity, wind_mag, action_amp, hard, f32 actions
    ("calm", False, "light", 0.0, 1.0, False, False),
    ("calm_clip", False, "light", 0.0, 2.5, False, False),
    ("wind", False, "light", 8.0, 1.3, False, False),
    ("hard", False, "light", 4.0, 1.5, True, False),
    ("f32act", False, "light", 3.0, 1.2, False, True),
    ("turb_light", True, "light", 8.0, 1.0, False, False),
    ("turb_moderate", True, "moderate", 6.0, 1.3, False, False),
    ("turb_severe", True, "severe", 5.0, 1.5, False, False),
    ("turb_hard", True, "severe", 8.0, 2.0, True, False),
]


def rel(a, b):
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())


bad = 0
for ci, (name, turb, inten, wind, amp, hard, f32) in enumerate(CASES):
    t0 = time.time()
    rs = np.random.RandomState(SEED + ci)
    env = G.make_env(turb, intensity=inten)
    g = G.run_episodes(env, N_EP, N_STEPS, rs, turb, wind_mag=wind, action_amp=amp, hard=hard, f32_actions=f32)
    cfg = build_config(sim_config_kw={"turbulence": turb, "turbulence_intensity": inten})
    worst = {"y": 0.0, "obs": 0.0, "rew": 0.0}
    steps = ends = 0
    for ep in range(N_EP):
        o = O.OracleEnv(cfg)
        obs = o.reset(g["init_state"][ep], g["init_target"][ep], g["noise"][ep] if turb else None)
        if np.abs(obs - g["obs0"][ep]).max() > 1e-12:
            print("  MISMATCH", name, ep, "reset observation"); bad += 1
        for t in range(int(g["n_valid"][ep])):
            obs, rew, done, term = o.step(g["actions"][ep, t], f32)
            s = o.get()
            steps += 1
            if done != bool(g["done"][ep, t]) or term != int(g["term"][ep, t]):
                print("  MISMATCH", name, ep, t, "done/term", done, term, g["done"][ep, t], g["term"][ep, t]); bad += 1
                break
            if s["nfev"] != int(g["nfev"][ep, t]):
                print("  MISMATCH", name, ep, t, "nfev", s["nfev"], g["nfev"][ep, t]); bad += 1
            if term < 10:
                worst["y"] = max(worst["y"], rel(s["y"], g["y"][ep, t]))
            worst["obs"] = max(worst["obs"], rel(obs, g["obs"][ep, t]))
            worst["rew"] = max(worst["rew"], abs(rew - g["reward"][ep, t]) / max(1.0, abs(g["reward"][ep, t])))
            ends += int(done)
    flag = "" if max(worst.values()) < 1e-9 else "  <-- above 1e-9"
    bad += int(bool(flag))
    print("%-14s %5d steps, %2d episode ends, worst rel: y %.1e obs %.1e reward %.1e  (%.0f s)%s"
          % (name, steps, ends, worst["y"], worst["obs"], worst["rew"], time.time() - t0, flag), flush=True)
print("MISMATCHES:", bad)
sys.exit(1 if bad else 0)
