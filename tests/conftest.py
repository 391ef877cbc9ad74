import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when no device is visible and they were not deselected.
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def emu_rt():
    """Runtime over the CPU-emulation build of the kernel sources (test infrastructure only)."""
    from emu.build_emu import build_emu
    from spectrogram_enhancement_b200 import _ffi, api
    return api.Runtime(_ffi.Library(build_emu()), "cpu")


@pytest.fixture(scope="session")
def cuda_rt():
    """The product runtime: libspecgpu.so on cuda:0.  Fails (does not skip) if the library is missing."""
    from spectrogram_enhancement_b200 import api
    return api.Runtime()
