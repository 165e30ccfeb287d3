"""Compile the sm_100a CUDA library in-tree (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_NAME = "libfootsies_b200.so"
LIB_PATH = os.environ.get("FOOTSIES_B200_LIB") or os.path.join(PKG_DIR, LIB_NAME)   # override: developer experiments only

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # fp32 ops round one by one, like the scalar C# expressions they restate
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libfootsies_b200.so")


def sources():
    return [os.path.join(CSRC, "footsies_kernels.cu")]


def deps():
    inc = os.path.join(os.path.dirname(PKG_DIR), "include", "footsies_b200.h")
    return sources() + [os.path.join(CSRC, f) for f in ("state_codec.h", "frame_tables.h", "frame_logic.cuh", "tables_host.h")] + [inc]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in deps() if os.path.exists(d))


def build(force=False, verbose=False):
    """Build libfootsies_b200.so next to this file; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
