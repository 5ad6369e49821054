"""Builds libfwb200.so (hand-written sm_100a CUDA + the C ABI of include/fwb200.h) in-tree with nvcc."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfwb200.so")
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("fw_step.cu", "fw_gae.cu", "fw_ppo.cu", "fw_replay.cu", "fw_comm.cu")]
HEADERS = [os.path.join(_HERE, "csrc", "fw_device.cuh"), os.path.join(_HERE, "csrc", "fw_math.cuh"), os.path.join(os.path.dirname(_HERE), "include", "fwb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libfwb200.so cannot be built (there is no CPU fallback)")
    return p


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into one shared library next to this file."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
