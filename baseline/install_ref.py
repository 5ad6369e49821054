#!/usr/bin/env python
"""Puts the UNMODIFIED reference libraries where the GPU box can see them: baseline/_ref/ (git-ignored, not
gpurun-ignored, so it travels with the snapshot; /root/reference itself does not exist on the box).

    python baseline/install_ref.py          (also run by __graft_entry__.build() when /root/reference is mounted)

The reference has no setup.py / pyproject.toml at its root — its three libraries are vendored under magpie/libs/ and
imported by path — so "installing" it is mirroring those trees: pyfly, fixed-wing-gym (gym_fixed_wing) and the
stable-baselines3 fork, source / config / parameter files only (no tensorboard logs, saved models or plots).
Consumers: bench.py's reference arms (the Python reference stepped in a SubprocVecEnv on the box's host cores) and
the drop-in tests that run the fork's own PPO on FixedWingVecEnv.  `gym` and `matplotlib`, which the reference imports
and this image lacks, are fabricated by oracle/refshim.py.  Nothing in the product package reads baseline/_ref.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
LIBS = ("pyfly", "fixed-wing-gym", "stable-baselines3")
KEEP = (".py", ".json", ".mat", ".txt", ".md")
EXTRA = (os.path.join("fixed-wing-gym", "gym_fixed_wing", "examples", "test_sets", "test_set_wind_none_step20-20-3.npy"),)


def install(reference_root="/root/reference", force=False):
    src_libs = os.path.join(reference_root, "magpie", "libs")
    if not os.path.isdir(src_libs):
        return None
    stamp = os.path.join(DEST, ".installed")
    if os.path.exists(stamp) and not force:
        return DEST
    dst_libs = os.path.join(DEST, "magpie", "libs")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    n = 0
    for lib in LIBS:
        for dirpath, dirnames, filenames in os.walk(os.path.join(src_libs, lib)):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", ".git", "models", "tensorboard", "docs")]
            for f in filenames:
                if f.endswith(KEEP):
                    rel = os.path.relpath(os.path.join(dirpath, f), src_libs)
                    out = os.path.join(dst_libs, rel)
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    shutil.copyfile(os.path.join(dirpath, f), out)
                    n += 1
    for rel in EXTRA:
        out = os.path.join(dst_libs, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src_libs, rel), out)
        n += 1
    with open(stamp, "w") as f:
        f.write("%d files mirrored from %s\n" % (n, src_libs))
    return DEST


if __name__ == "__main__":
    d = install(force="--force" in sys.argv)
    print("baseline/_ref:", d if d else "reference tree not mounted; nothing installed")
