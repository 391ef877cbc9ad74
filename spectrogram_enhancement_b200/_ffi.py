"""ctypes binding of libspecgpu.so (include/specgpu.h).

The product path is CUDA-only: `load()` raises if the in-tree library has not been built or cannot
be loaded -- there is no CPU fallback.  (The CPU test-suite loads an emulation build of the same
sources through `Library(path)` explicitly; nothing in the package does.)
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPECGPU_LIB") or os.path.join(_PKG, "libspecgpu.so")   # SPECGPU_LIB: another build of the CUDA library

OK = 0
ERR_INVALID_ARG = -1
ERR_UNSUPPORTED_NPERSEG = -2
ERR_CUDA = -3
ERR_NCCL = -4
ERR_WORKSPACE = -5
ERR_UNSUPPORTED_SHAPE = -6
PIPE_CLIP = 1          # specgpu_pipeline flags
PIPE_FALLBACK = 2
PIPE_STATIC_TILES = 4

DETREND = {False: 0, None: 0, "constant": 1, "linear": 2}
SCALING = {"density": 0, "spectrum": 1}
WINDOW = {"hann": 1, "hanning": 1, "han": 1, "hamming": 2, "hamm": 2, "ham": 2, "boxcar": 3, "box": 3,
          "ones": 3, "rect": 3, "rectangular": 3}


class StftParams(C.Structure):
    _fields_ = [("nperseg", C.c_int32), ("noverlap", C.c_int32), ("detrend", C.c_int32), ("scaling", C.c_int32),
                ("window", C.c_int32), ("reserved", C.c_int32), ("fs", C.c_double), ("eps", C.c_double)]


_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float

# name -> (restype, argtypes); every symbol include/specgpu.h declares
PROTOTYPES = {
    "specgpu_version": (C.c_int, []),
    "specgpu_init": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "specgpu_destroy": (C.c_int, [_vp]),
    "specgpu_last_error": (C.c_char_p, [_vp]),
    "specgpu_workspace_reserve": (C.c_int, [_vp, _i64]),
    "specgpu_plan_create": (C.c_int, [_vp, C.POINTER(StftParams), C.POINTER(C.c_double), C.POINTER(_vp)]),
    "specgpu_plan_destroy": (C.c_int, [_vp]),
    "specgpu_plan_num_segments": (_i64, [_vp, _i64]),
    "specgpu_plan_num_freqs": (_i32, [_vp]),
    "specgpu_plan_axes": (C.c_int, [_vp, _i64, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "specgpu_spectrogram": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "specgpu_specgr": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp]),
    "specgpu_stft_num_segments": (_i64, [_vp, _i64, C.c_int, C.c_int]),
    "specgpu_stft": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, C.c_int, C.c_int, _vp, _i64, _vp]),
    "specgpu_rescale": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "specgpu_norm": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "specgpu_clip": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "specgpu_quantfilt": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp, _vp, _vp, _vp]),
    "specgpu_gaussblr": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _i64, _vp, _vp]),
    "specgpu_meansub": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp]),
    "specgpu_filter_chain": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, C.c_float, _i32, _i32, _vp, _i64, _vp]),
    "specgpu_morph": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp]),
    "specgpu_svd_denoise": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp,
                                      _vp, _vp]),
    "specgpu_compute_signal": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp]),
    "specgpu_patch": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _i32, _vp]),
    "specgpu_unpatch": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i32, _i32, _vp, _i32, _i64, _vp]),
    "specgpu_csd_spectra": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "specgpu_csd_pairs": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "specgpu_csd_pairs_block": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _vp, _vp]),
    "specgpu_csd_spectra_blocked": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i32, _i32, _vp]),
    "specgpu_csd_pairs_bins": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _vp, _vp]),
    "specgpu_csd_frames": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i64, _vp, _vp]),
    "specgpu_csd_allpairs": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "specgpu_pipeline": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp]),
    "specgpu_copy_rows": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "specgpu_set_pipeline_group": (C.c_int, [_vp, _i32]),
    "specgpu_set_pipeline_interlock": (C.c_int, [_vp, _vp, _vp]),
    "specgpu_set_power_iterations": (C.c_int, [_vp, _i32]),
    "specgpu_launch_count": (_i64, [_vp]),
    "specgpu_profile_enable": (C.c_int, [_vp, C.c_int]),
    "specgpu_profile_count": (C.c_int, [_vp]),
    "specgpu_profile_name": (C.c_char_p, [_vp, C.c_int]),
    "specgpu_profile_ms": (C.c_double, [_vp, C.c_int]),
    "specgpu_profile_calls": (_i64, [_vp, C.c_int]),
}


class SpecGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libspecgpu error {code}: {msg}")
        self.code = code
        self.msg = msg


class Library:
    """A loaded libspecgpu with typed entry points."""

    def __init__(self, path: str = LIB_PATH, allow_missing: bool = False):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: build it with `python -m spectrogram_enhancement_b200.build` "
                "(libspecgpu is CUDA-only; there is no CPU fallback)")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            try:
                fn = getattr(self.dll, name)
            except AttributeError:
                if allow_missing:      # development aid for partially built test libraries only
                    continue
                raise
            fn.restype = res
            fn.argtypes = args
            setattr(self, name[len("specgpu_"):], fn)


_lib = None


def load() -> Library:
    global _lib
    if _lib is None:
        if "SPECGPU_LIB" not in os.environ:
            try:        # a library older than its sources is a silent way to measure the wrong code: say so
                from . import build as _build
                digest = _build._deps_digest() + "|" + " ".join(_build.NVCC_FLAGS)
                if not _build._up_to_date(LIB_PATH, digest):
                    import sys
                    print("libspecgpu: WARNING: libspecgpu.so is older than csrc/ -- run "
                          "`python -m spectrogram_enhancement_b200.build`", file=sys.stderr)
            except Exception:
                pass
        _lib = Library(LIB_PATH)
    return _lib
