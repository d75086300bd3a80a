"""Builds libsadmc_gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per system family (csrc/kernels_*.cu) plus the host side (csrc/engine.cu), compiled in
parallel and linked into one shared library; objects are cached under csrc/_obj/ and rebuilt when any header or
the unit's source is newer.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "libsadmc_gpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # every f64 operation rounds once, as in the reference (rustc never fuses); the
    # kernels ask for FMA explicitly where the tolerance tier allows it.
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall",
]


def units():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h"))]
    d += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return d


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in sources)


def needs_build():
    return _stale(OUT, units() + headers())


HOST_SRC = os.path.join(ROOT, "host")
HOST_BIN = os.path.join(HERE, "bin", "histogram")


def build_host(force=False, verbose=False):
    """The compiled host: host/histogram.cpp -> sad_monte_carlo_b200/bin/histogram (dlopens libsadmc_gpu.so)."""
    srcs = [os.path.join(HOST_SRC, f) for f in os.listdir(HOST_SRC)] + [os.path.join(ROOT, "include", "sadmc_gpu.h")]
    if not force and not _stale(HOST_BIN, srcs):
        return HOST_BIN
    os.makedirs(os.path.dirname(HOST_BIN), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", HOST_BIN, os.path.join(HOST_SRC, "histogram.cpp"), "-ldl"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=ROOT)
    return HOST_BIN


def build(force=False, verbose=False, extra=()):
    build_host(force=force, verbose=verbose)
    if not force and not needs_build():
        return OUT
    extra = list(extra) + os.environ.get("SADMC_NVCC_EXTRA", "").split()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = headers()
    todo = [u for u in units() if force or extra or _stale(_obj(u), [u] + hdrs)]

    def compile_one(src):
        cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-c", "-o", _obj(src), src]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1) or 1) as ex:
        for src, rc, log in ex.map(compile_one, todo):
            if verbose and log.strip():
                print(log, flush=True)
            if rc != 0:
                sys.stderr.write(log)
                raise subprocess.CalledProcessError(rc, "nvcc " + src)
    link = ["nvcc", "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + [_obj(u) for u in units()]
    if verbose:
        print(" ".join(link), flush=True)
    subprocess.check_call(link, cwd=HERE)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "-v" in sys.argv else [])
