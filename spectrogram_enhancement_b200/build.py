"""Build libspecgpu.so (sm_100a) in-tree with nvcc.

    python -m spectrogram_enhancement_b200.build [--force] [--verbose]

The library is CUDA-only (`-gencode arch=compute_100a,code=sm_100a -lineinfo`).  (The CPU test-suite has its own
builder, tests/emu/build_emu.py, which compiles the same kernel sources against a CUDA-execution-model emulator;
nothing in this package builds or loads it.)
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(ROOT, "build")
LIB = os.path.join(PKG, "libspecgpu.so")

SOURCES = ["specgpu.cu", "stft.cu", "stft_gram.cu", "elementwise.cu", "quantile.cu", "svd.cu", "eig_tridiag.cu", "gram_tc.cu", "csd.cu", "imgchain.cu"]

NVCC_FLAGS = [
    "-std=c++17", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _deps_digest(extra=()) -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(ROOT, "include", "specgpu.h")] + list(extra)
    for f in files:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(os.path.relpath(p, ROOT).encode())      # relative: the stamp survives a move of the checkout
            h.update(open(p, "rb").read())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def _up_to_date(lib, digest):
    stamp = lib + ".digest"
    return os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == digest


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    digest = _deps_digest() + "|" + " ".join(NVCC_FLAGS)
    if not force and _up_to_date(LIB, digest):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    objs = []

    def one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        flags = list(NVCC_FLAGS) + (["-Xptxas", "-v"] if verbose else [])
        out = _run([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj])
        if verbose:
            open(obj + ".ptxas.log", "w").write(out)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, _sources()))
    _run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"])
    open(LIB + ".digest", "w").write(digest)
    return LIB


if __name__ == "__main__":
    print(build_cuda("--force" in sys.argv, verbose="--verbose" in sys.argv))
