"""
Builds libvc2b200.so (the C-ABI CUDA library, include/vc2_b200.h) in-tree with nvcc for sm_100a.

    python -m vc2_conformance_b200._build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the tree.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libvc2b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
if os.environ.get("VC2B_DEV_BUILD"):  # developer-only: a single filter pair, seconds to compile
    FLAGS.append("-DVC2B_DEV_BUILD")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    inc = os.path.join(os.path.dirname(HERE), "include", "vc2_b200.h")
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [inc]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    hdrs = headers()
    jobs = []
    objs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([NVCC] + ARCH + FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stdout + r.stderr))
        return r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
