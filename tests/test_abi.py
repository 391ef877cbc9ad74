"""The C-ABI shared library loads and exports every symbol include/specgpu.h declares (no compute
calls: this runs without a GPU), the ctypes table matches the header, and the product loader fails
loudly rather than falling back when the CUDA library is absent."""
import ctypes
import os
import re

import pytest

from spectrogram_enhancement_b200 import _ffi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "specgpu.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(specgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_ctypes_binds():
    assert header_symbols() == sorted(_ffi.PROTOTYPES)


def test_cuda_library_builds_and_exports_every_symbol():
    path = build.build_cuda()                       # nvcc cross-compiles sm_100a without a GPU
    dll = ctypes.CDLL(path)
    for name in header_symbols():
        assert hasattr(dll, name), name
    dll.specgpu_version.restype = ctypes.c_int
    assert dll.specgpu_version() == 1               # major*1000 + minor


def test_cuda_library_is_sm100a_with_tcgen05():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    path = build.build_cuda()
    elf = subprocess.run([cuobjdump, "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass      # tcgen05.mma / tcgen05.ld (B200_PROFILING.md)
    # TMA: bulk copies (STFT input span, Gram partial), tensor loads (Gram operands, CSD slabs), tensor stores (STFT tile)
    assert "UBLKCP" in sass and "UTMALDG" in sass and "UTMASTG" in sass


def test_missing_library_raises(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ffi.Library(str(tmp_path / "libspecgpu.so"))


def test_runtime_requires_cuda():
    import torch
    from spectrogram_enhancement_b200 import api
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        api.Runtime()
