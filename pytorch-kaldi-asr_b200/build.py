"""In-tree nvcc build of libpka_b200.so for sm_100a (no torch linkage: plain C ABI, static cudart).

    python pytorch-kaldi-asr_b200/build.py [--force] [--verbose]

The shared object is written next to the sources (pytorch-kaldi-asr_b200/csrc/libpka_b200.so); it is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libpka_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-DPKA_BUILD"] + os.environ.get("PKA_NVCC_EXTRA", "").split()


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libpka_b200.so for sm_100a)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "pka_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = [s[:-3] + ".o" for s in srcs]
    todo = [(s, o) for s, o in zip(srcs, objs) if force or _stale(o, [s] + headers)]

    def compile_one(so):
        s, o = so
        cmd = [nvcc()] + ARCH + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (os.path.basename(s), r.stdout, r.stderr))
        return os.path.basename(s), r.stderr

    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for name, log in ex.map(compile_one, todo):
                if verbose and log:
                    print("---- %s\n%s" % (name, log))
    if todo or force or _stale(LIB, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcuda"]
        # libcuda is only needed for cuTensorMapEncodeTiled; resolve it lazily through the runtime instead of linking
        cmd = [c for c in cmd if c != "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
