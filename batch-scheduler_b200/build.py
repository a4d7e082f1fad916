"""Builds the C-ABI shared library in-tree: nvcc, sm_100a only.

    python -m importlib ...  # not importable by dotted name (hyphen); use __graft_entry__.build()
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbsched.so")
SOURCES = ["engine.cu"]
HEADERS = ["kernels.cuh", "sort.cuh", "replay.cuh", os.path.join("..", "..", "include", "bsched.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine is CUDA-only, there is no CPU build")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(CSRC, "plugin.cpp"),
                                                                  os.path.join(CSRC, "plugin.hpp")]
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in files)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    plugin = os.path.join(CSRC, "plugin.cpp")
    if os.path.exists(plugin):
        srcs.append(plugin)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-shared", "-Xcompiler", "-fPIC,-Wall,-fopenmp", "-fmad=false", "-diag-suppress=186,550", "-o", LIB] + srcs + ["-lgomp"]
    extra = os.environ.get("BS_NVCC_EXTRA", "").split()
    if extra:
        cmd[1:1] = extra
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
