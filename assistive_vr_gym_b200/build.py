"""Builds the CUDA extension in-tree: assistive_vr_gym_b200/libavg_b200.so (sm_100a only, explicit nvcc)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavg_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("avg_kernels.cu", "avg_capi.cu")]
HEADERS = [os.path.join(_HERE, "csrc", f) for f in ("avg_kernels.h", "avg_math.cuh")] + [
    os.path.join(_HERE, "..", "include", f) for f in ("avg_model.h", "avg_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-Xcompiler", "-fPIC", "-shared", "-std=c++17"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 extension cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
