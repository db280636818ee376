"""Builds libsbx_b200.so (the C-ABI library of include/sbx.h) in-tree with nvcc for sm_100a.

    python spectrogram-rs_b200/build.py [--force] [--verbose]

The kernels' template instances are spread over several translation units (csrc/sbx_dispatch_*.cu) that are
compiled in parallel and linked into one shared library.  The library lands next to this file so that it travels
to the GPU box with the repo snapshot; it and the object directory are git-ignored.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsbx_b200.so")
# (object name, source, extra flags)
UNITS = [
    ("plan", "sbx_plan.cu", []),
    ("live", "sbx_live.cu", []),
    ("multi", "sbx_multi.cu", []),
    ("fused0", "sbx_dispatch_fused.cu", ["-DSBX_FUSED_PART=0"]),
    ("fused1", "sbx_dispatch_fused.cu", ["-DSBX_FUSED_PART=1"]),
    ("real", "sbx_dispatch_real.cu", []),
    ("q2", "sbx_dispatch_q2.cu", []),
    ("wide", "sbx_dispatch_wide.cu", []),
    ("large", "sbx_dispatch_large.cu", []),
]
DEPS = sorted({u[1] for u in UNITS}) + ["stft_kernels.cuh", "stft_q2.cuh", "sbx_pk.cuh", "stft_large.cuh", "stft_generic.cuh", "sbx_internal.h", "sbx_launch.h",
                                         "sbx_tables.h", "luts.inc", os.path.join("..", "..", "include", "sbx.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall",
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


def build(force=False, verbose=False, extra_flags=(), lib=LIB, obj_dir=OBJ):
    """extra_flags / lib / obj_dir: variant libraries for A/B measurements (tools/build_variant.sh, SBX_LIB_PATH)."""
    if lib == LIB and not force and up_to_date():
        return LIB
    os.makedirs(obj_dir, exist_ok=True)
    common = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])

    def compile_unit(unit):
        name, src, flags = unit
        obj = os.path.join(obj_dir, name + ".o")
        proc = subprocess.run(common + flags + ["-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        return name, obj, proc

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_unit, UNITS))
    for name, _, proc in results:
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed compiling {name}:\n" + proc.stderr[-4000:])
    link = subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + [o for _, o, _ in results],
                          capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("nvcc failed linking libsbx_b200.so:\n" + link.stderr[-4000:])
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
