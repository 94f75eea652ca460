"""Build recipe for the native library.

    python -m plonky2_bn254_b200.build            # libpb254.so  (nvcc, sm_100a)  -- the product
    python -m plonky2_bn254_b200.build --hostsim  # tests/hostsim/libpb254_hostsim.so (g++) -- test-only

nvcc cross-compiles for sm_100a without a GPU. The .so files are git-ignored but travel to the GPU
box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "pb254.cu")
LIB = os.path.join(HERE, "libpb254.so")
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_LIB = os.path.join(HOSTSIM_DIR, "libpb254_hostsim.so")
HOST_CXX = "/usr/bin/g++"


def _sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(ROOT, "include", "pb254.h")]


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_cuda(force=False, verbose=False):
    if not force and not _stale(LIB):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-ccbin", HOST_CXX, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


def build_hostsim(force=False):
    os.makedirs(HOSTSIM_DIR, exist_ok=True)
    if not force and not _stale(HOSTSIM_LIB):
        return HOSTSIM_LIB
    cmd = [HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-DPB254_HOSTSIM", "-x", "c++", SRC,
           "-o", HOSTSIM_LIB]
    subprocess.check_call(cmd)
    return HOSTSIM_LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    if "--hostsim" in sys.argv:
        print(build_hostsim(force))
    else:
        print(build_cuda(force, verbose="-v" in sys.argv))
