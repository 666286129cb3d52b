"""Builds the CUDA libraries in-tree with nvcc for sm_100a (no JIT cache: the .so files travel with the repo snapshot).

    libetude_b200.so      the product: the C ABI of include/etude_b200.h + the kernel-level entry points of
                          include/etude_b200_kernels.h, nothing else
    libetude_b200_dev.so  the same source compiled with -DETUDE_DEV_BUILD: adds the micro-benchmarks, the kernel timelines
                          and the generic GEMM epilogues declared in include/etude_b200_dev.h.  Loaded by tests/ only.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libetude_b200.so")
LIB_DEV = os.path.join(HERE, "libetude_b200_dev.so")
SOURCES = ["api.cu"]


def _deps():
    root = os.path.dirname(HERE)
    return glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(root, "include", "*.h"))


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(f) > t for f in _deps())


def _cmd(lib, dev, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC,-O2,-pthread", "-o", lib] + [os.path.join(CSRC, s) for s in SOURCES]
    if dev:
        cmd.insert(1, "-DETUDE_DEV_BUILD")
    for flag in os.environ.get("ETUDE_NVCC_FLAGS_DEV" if dev else "ETUDE_NVCC_FLAGS", "").split():
        cmd.insert(1, flag)   # experiments only (e.g. -DAQ_EXCHANGE_PULL=1 in the dev build)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    return cmd


def build(force=False, verbose=False, dev=True):
    """Compiles whichever of the two libraries is stale (both compile in parallel).  Returns the product library path."""
    jobs = []
    for lib, is_dev in ((LIB, False), (LIB_DEV, True)):
        if is_dev and not dev:
            continue
        if force or needs_build(lib):
            jobs.append((lib, subprocess.Popen(_cmd(lib, is_dev, verbose))))
    for lib, p in jobs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed for {lib}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
