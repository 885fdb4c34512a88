"""In-tree build of the CUDA library (sm_100a only): incompact3d_b200/libx3d_b200.so.

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libx3d_b200.so")

NVCC = os.environ.get("X3D_NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = os.environ.get("X3D_CXX", "/usr/bin/g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", HOSTCXX,
          "-Xptxas", "-v", "--expt-relaxed-constexpr", "--extended-lambda"]
# per-file additions.  x3d_ibm.cu: the reconstruction polynomials / splines are ill-conditioned when a fluid point sits
# next to a wall (izap = 0); without fused multiply-adds the kernels do the reference's operations one for one
# (a handful of points per line: speed is irrelevant there).
EXTRA = {"x3d_ibm.cu": ["--fmad=false"]}


def _newer(src, obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "x3d_b200.h")]
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _newer(s, o, hdrs):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        cmd = [NVCC] + ARCH + CFLAGS + EXTRA.get(os.path.basename(s), []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(o + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        return s

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s in ex.map(cc, jobs):
                if verbose:
                    print("compiled", os.path.basename(s))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", HOSTCXX, "-o", LIB] + objs + ["-lcufft", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
