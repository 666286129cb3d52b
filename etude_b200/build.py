"""Builds libetude_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libetude_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "attention.cuh", "attention2.cuh", "attention3.cuh", "attention4.cuh", "chain.cuh", "chain2.cuh", "mmabench.cuh", "embed.cuh", "embed2.cuh", "logmel.cuh", "logmel2.cuh", "ingest.cuh", "notes.cuh",
           os.path.join("..", "..", "include", "etude_b200.h"), os.path.join("..", "..", "include", "etude_b200_kernels.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC,-O2,-pthread", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
