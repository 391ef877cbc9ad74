"""Build the CPU-emulation library of the kernel sources (TEST INFRASTRUCTURE ONLY).

    python tests/emu/build_emu.py [--force]

Compiles spectrogram_enhancement_b200/csrc/*.cu as plain C++ with -DSPECGPU_EMULATE against cuda_emu.h (CUDA threads ->
OS threads) into tests/emu/libspecgpu_emu.so.  Only the CPU test-suite loads it; the product package neither builds
nor loads it and has no CPU path.
"""
from __future__ import annotations

import os
import sys
from concurrent.futures import ThreadPoolExecutor

EMU_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(EMU_DIR))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from spectrogram_enhancement_b200.build import BUILD, CSRC, _deps_digest, _run, _sources, _up_to_date  # noqa: E402

EMU_LIB = os.path.join(EMU_DIR, "libspecgpu_emu.so")
GXX_FLAGS = ["-std=c++20", "-O2", "-fPIC", "-pthread", "-DSPECGPU_EMULATE", "-x", "c++", "-I", EMU_DIR,
             "-Wno-unknown-pragmas", "-Wno-attributes"]


def build_emu(force: bool = False) -> str:
    emu_files = [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(EMU_DIR, "cuda_emu.cpp")]
    digest = _deps_digest(emu_files) + "|" + " ".join(GXX_FLAGS)
    if not force and _up_to_date(EMU_LIB, digest):
        return EMU_LIB
    os.makedirs(os.path.join(BUILD, "emu"), exist_ok=True)

    def one(path):
        obj = os.path.join(BUILD, "emu", os.path.basename(path).rsplit(".", 1)[0] + ".o")
        _run(["g++"] + GXX_FLAGS + ["-c", path, "-o", obj])
        return obj

    srcs = [os.path.join(CSRC, s) for s in _sources()] + [emu_files[1]]
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, srcs))
    _run(["g++", "-shared", "-pthread", "-o", EMU_LIB] + objs)
    open(EMU_LIB + ".digest", "w").write(digest)
    return EMU_LIB



if __name__ == "__main__":
    print(build_emu("--force" in sys.argv))
