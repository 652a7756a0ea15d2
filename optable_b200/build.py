"""In-tree build of the CUDA extension (optable_b200/liboptb.so) with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "liboptb.so")
SOURCES = ["optb.cu"]
DEPS = ["optb.cu", "optb_device.cuh", "optb_flags.cuh", os.path.join("..", "..", "include", "optb.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-ldl",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", SO, *SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        subprocess.run(cmd, cwd=CSRC, check=True)
    return SO


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))
