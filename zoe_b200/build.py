"""Builds zoe_b200/libzoe_cuda.so in-tree with nvcc for sm_100a (no torch involved)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzoe_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
    "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-cudart", "shared",
]
# Development only: ZOE_CUDA_SPLIT_COMPILE=N optimises the kernels of the one translation unit in parallel (build 165 s ->
# 75 s) -- NOT for a build that is measured: with -split-compile ptxas emitted 17 % more instructions over the 236
# kernels (outside the hot loops) and every configuration ran 1-5 % slower (cfg 2 295.5 vs 292.1 ms, cfg 4 341 vs 324 ms,
# same box, same sources); which kernels are hit depends on how the functions fall into the partitions.
if os.environ.get("ZOE_CUDA_SPLIT_COMPILE"):
    FLAGS += ["-split-compile", os.environ["ZOE_CUDA_SPLIT_COMPILE"]]


def sources():
    return [os.path.join(CSRC, "zoe_cuda.cu")]


def deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(os.path.dirname(HERE), "include", "zoe_cuda.h"))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps()):
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
