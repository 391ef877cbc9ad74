"""B200-native (sm_100a) spectrogram / denoise / cross-spectrum hot path of
PlasmaControl/spectrogram-enhancement behind the reference's own Python function signatures.

    from spectrogram_enhancement_b200 import specgr, quantfilt, denoiseSignal, patch, csd_allpairs

Everything computes in libspecgpu.so (hand-written CUDA, C ABI in include/specgpu.h, loaded through
ctypes); importing the package does not need a GPU, calling any function does.
"""
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all

__version__ = "0.1.0"
__all__ = list(_api_all)
