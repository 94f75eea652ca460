"""Build recipe for the native library.

    python -m plonky2_bn254_b200.build            # libpb254.so  (nvcc, sm_100a)  -- the product
    python -m plonky2_bn254_b200.build --hostsim  # tests/hostsim/libpb254_hostsim.so (g++) -- test-only

nvcc cross-compiles for sm_100a without a GPU. Translation units are compiled in parallel and only
when stale. The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpb254.so")
OBJ_DIR = os.path.join(HERE, "build")
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_LIB = os.path.join(HOSTSIM_DIR, "libpb254_hostsim.so")
HOST_CXX = "/usr/bin/g++"


def _units():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + \
           [os.path.join(ROOT, "include", "pb254.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + os.path.basename(cmd[-1]))
    return r.stdout + r.stderr


def _build(objdir, lib, compile_cmd, link_cmd, force, verbose):
    os.makedirs(objdir, exist_ok=True)
    hdrs = _headers()
    jobs, objs = [], []
    for u in _units():
        src = os.path.join(CSRC, u)
        obj = os.path.join(objdir, u[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append(compile_cmd + ["-c", src, "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(_run, jobs):
                if verbose and out:
                    sys.stderr.write(out)
    if jobs or not os.path.exists(lib):
        _run(link_cmd + objs + ["-o", lib])
    return lib


def build_cuda(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cc = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin", HOST_CXX,
          "-Xcompiler", "-fPIC"]
    if verbose:
        cc += ["-Xptxas", "-v"]
    ld = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOST_CXX, "-shared", "-Xcompiler", "-fPIC"]
    return _build(OBJ_DIR, LIB, cc, ld, force, verbose)


def build_hostsim(force=False):
    cc = [HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-DPB254_HOSTSIM", "-x", "c++"]
    ld = [HOST_CXX, "-shared", "-fopenmp"]
    return _build(os.path.join(HOSTSIM_DIR, "obj"), HOSTSIM_LIB, cc, ld, force, False)


if __name__ == "__main__":
    force = "--force" in sys.argv
    if "--hostsim" in sys.argv:
        print(build_hostsim(force))
    else:
        print(build_cuda(force, verbose="-v" in sys.argv))
