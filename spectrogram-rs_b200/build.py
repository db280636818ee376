"""Builds libsbx_b200.so (the C-ABI library of include/sbx.h) in-tree with nvcc for sm_100a.

    python spectrogram-rs_b200/build.py [--force] [--verbose]

The shared library lands next to this file so that it travels to the GPU box with the repo
snapshot; it is git-ignored (*.so).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsbx_b200.so")
SOURCES = ["sbx_plan.cu", "sbx_live.cu"]
DEPS = SOURCES + ["stft_kernels.cuh", "stft_large.cuh", "stft_generic.cuh", "sbx_internal.h", "sbx_tables.h", "luts.inc", os.path.join("..", "..", "include", "sbx.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall",
    "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS) and os.path.getmtime(__file__) <= t


def build(force=False, verbose=False, extra_flags=()):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags)
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libsbx_b200.so:\n" + proc.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
