"""Build libbmcts.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree.

Usage: python planner-mcts_b200/build.py [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libbmcts.so")

# --fmad=false / -ffp-contract=off: FP32 contraction only where bm_fma() is
# written, so the kernels round exactly like the CPU oracle (include/bmcts_math.h).
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall", "-shared",
    "-I", INCLUDE, "-I", CSRC, "-I", HOST,
]


def sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp"))]
    if os.path.isdir(HOST):
        srcs += [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp")]
    return srcs


def deps():
    d = sources()
    for folder in (CSRC, HOST, INCLUDE):
        if os.path.isdir(folder):
            d += [os.path.join(folder, f) for f in os.listdir(folder)
                  if f.endswith((".h", ".hpp", ".cuh"))]
    return d


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in deps())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libbmcts.so cannot be built (there is no CPU fallback)")
    extra = os.environ.get("BMCTS_NVCC_EXTRA", "").split()  # tuning experiments only
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libbmcts.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
