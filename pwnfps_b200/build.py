"""pwnfps_b200/build.py -- builds libpwntrace.so in-tree for sm_100a.

    python -m pwnfps_b200.build [--force]

nvcc cross-compiles without a GPU.  The image exports CC/CXX=/opt/gcc/bin/*,
wrappers that lack libgomp; the PATH gcc/g++ are pinned instead.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libpwntrace.so")
HOSTDIR = os.path.join(HERE, "host")
HOSTLIB = os.path.join(HERE, "libpwnhost.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # the reference build has no FMA (Makefile:3-5 has no -mfma)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
GCC_FLAGS = ["-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra", "-Wno-unused-function"]


def _newer(target, sources):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources():
    out = [os.path.join(INC, "pwntrace.h"), os.path.abspath(__file__)]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".c", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def host_sources():
    out = [os.path.join(INC, "pwntrace.h"), os.path.abspath(__file__)]
    for f in sorted(os.listdir(HOSTDIR)):
        if f.endswith((".c", ".h")):
            out.append(os.path.join(HOSTDIR, f))
    return out


def build_host(force=False):
    """libpwnhost.so: level loader, camera script (plain C, no CUDA)."""
    if not force and not _newer(HOSTLIB, host_sources()):
        return HOSTLIB
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    cfiles = [os.path.join(HOSTDIR, f) for f in sorted(os.listdir(HOSTDIR))
              if f.endswith(".c") and f not in ("pwnbench.c",)]
    subprocess.run([gcc] + GCC_FLAGS + ["-shared", "-I", INC, "-I", HOSTDIR, "-o", HOSTLIB] + cfiles + ["-lm"], check=True)
    return HOSTLIB


def build(force=False, verbose=False):
    build_host(force)
    if not force and not _newer(LIB, sources()):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    gxx = shutil.which("g++") or "/usr/bin/g++"
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    o_host = os.path.join(obj_dir, "pwn_hostmath.o")
    o_cu = os.path.join(obj_dir, "pwntrace.o")
    subprocess.run([gcc] + GCC_FLAGS + ["-I", INC, "-I", CSRC, "-c", os.path.join(CSRC, "pwn_hostmath.c"), "-o", o_host], check=True)
    r = subprocess.run([nvcc, "-ccbin", gxx] + NVCC_FLAGS + ["-I", INC, "-I", CSRC, "-c",
                        os.path.join(CSRC, "pwntrace.cu"), "-o", o_cu],
                       check=False, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(obj_dir, "ptxas.log"), "w") as f:
        f.write(log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(log)
    subprocess.run([nvcc, "-ccbin", gxx, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-o", LIB, o_cu, o_host, "-lcudart"], check=True)
    build_pwnbench()
    return LIB


BENCH_BIN = os.path.join(HERE, "pwnbench")


def build_pwnbench():
    """The headless C host (host/pwnbench.c): plain gcc against the two shared libraries."""
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    build_host()
    subprocess.run([gcc, "-O2", "-Wall", "-I", INC, "-I", HOSTDIR, "-o", BENCH_BIN, os.path.join(HOSTDIR, "pwnbench.c"),
                    "-L", HERE, "-lpwnhost", "-lpwntrace", "-Wl,-rpath,$ORIGIN", "-lm"], check=True)
    return BENCH_BIN


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
