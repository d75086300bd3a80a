"""Builds libsadmc_gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "engine.cu")]
OUT = os.path.join(HERE, "libsadmc_gpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # every f64 operation rounds once, as in the reference (rustc never fuses); the
    # kernels ask for FMA explicitly where the tolerance tier allows it.
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall",
    "-shared", "-cudart", "shared",
]


def deps():
    d = list(SRC)
    for sub in ("csrc",):
        for f in os.listdir(os.path.join(HERE, sub)):
            if f.endswith((".cuh", ".hpp", ".h")):
                d.append(os.path.join(HERE, sub, f))
    for f in os.listdir(os.path.join(ROOT, "include")):
        d.append(os.path.join(ROOT, "include", f))
    return d


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in deps())


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return OUT
    extra = list(extra) + os.environ.get("SADMC_NVCC_EXTRA", "").split()
    cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-o", OUT] + SRC
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=HERE)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "-v" in sys.argv else [])
