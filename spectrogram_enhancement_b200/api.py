"""Host-side mirror of the reference's Python functions on top of libspecgpu (ctypes, C ABI).

Every function keeps the name, argument meaning and error behaviour of the reference function it
replaces (file:line into PlasmaControl/spectrogram-enhancement):

    specgr / norm / rescale / quantfilt     spec_denoising/pipeline_data.py:28-49
    omega / computeSignal / denoiseSignal   spec_denoising/denoising_by_svd.ipynb:155-229, clip :280-281
    patch / unpatch / reshape               VAE/manual_scan.py:28-54
    ae_co2                                  interferometer/crosspowerspec.py:39 (body absent upstream)
    spectrogram / stft / csd                scipy.signal call sites of the above (pipeline_data.py:32)

plus the batched array-level entry points the reference's `for shot: for channel:` loops collapse
into (`spectrogram_batch`, `csd_allpairs`, `pipeline`).

Buffers: numpy arrays are staged to the device and results come back as numpy; torch CUDA tensors
are used in place (zero copy) and results come back as torch tensors on the same device.  All
arithmetic runs in float32 on the GPU (the reference's own dtype for float32 ECE data).  There is
no CPU implementation behind these functions: without libspecgpu.so and a B200 they raise.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import threading

import numpy as np
import torch

from . import _ffi

__all__ = [
    "Runtime", "default_runtime", "DEFAULT_SPEC_PARAMS", "spectrogram", "stft", "csd", "csd_allpairs",
    "spectrogram_batch", "specgr_array", "specgr", "load_shot", "save_shot_hdf5", "load_hdf5_dataset", "norm", "rescale", "quantfilt", "quantfilt_mask", "gaussblr", "meansub", "morph",
    "filter_chain", "process_shot", "omega",
    "computeSignal", "denoiseSignal", "clip", "patch", "unpatch", "reshape", "pipeline", "HostPipeline", "ShotStreams", "ae_co2",
]

# spec_denoising/pipeline_data.py:77-84
DEFAULT_SPEC_PARAMS = {"nperseg": 512, "noverlap": 256, "fs": 500000, "window": "hamm", "scaling": "density",
                       "detrend": "linear", "eps": 1e-11}


def _is_torch(x):
    return isinstance(x, torch.Tensor)


def _ld(t):
    """Leading dimension (elements between consecutive rows) of a row-major tensor view."""
    return max(int(t.stride(-2)), int(t.shape[-1]), 1) if t.shape[-2] > 1 else max(int(t.shape[-1]), 1)


class Runtime:
    """One libspecgpu context on one device, with a plan cache.

    `lib`/`device` exist so the CPU test-suite can drive the same host logic against the emulation
    build of the kernels (tests/emu); the package itself only ever creates the CUDA runtime."""

    def __init__(self, lib: _ffi.Library | None = None, device=None):
        self.lib = lib if lib is not None else _ffi.load()
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("libspecgpu needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self._ctx = C.c_void_p()
        index = self.device.index if self.device.type == "cuda" and self.device.index is not None else 0
        rc = self.lib.init(index, C.byref(self._ctx))
        if rc != 0:
            raise _ffi.SpecGpuError(rc, "specgpu_init failed (is this an sm_100 device?)")
        self._plans = {}
        self._lock = threading.Lock()

    # ---- plumbing -------------------------------------------------------------------------------
    def close(self):
        if self._ctx:
            for h in self._plans.values():
                self.lib.plan_destroy(h)
            self._plans.clear()
            self.lib.destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            msg = self.lib.last_error(self._ctx)
            msg = msg.decode() if msg else ""
            if rc in (_ffi.ERR_INVALID_ARG, _ffi.ERR_UNSUPPORTED_NPERSEG, _ffi.ERR_UNSUPPORTED_SHAPE):
                raise ValueError(msg)
            raise _ffi.SpecGpuError(rc, msg)

    def stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return None

    def launch_count(self) -> int:
        return int(self.lib.launch_count(self._ctx))

    def set_pipeline_group(self, channels: int):
        """Channels per group of `pipeline` (0 = automatic; >= batch = one group on the caller's stream)."""
        self.check(self.lib.set_pipeline_group(self._ctx, int(channels)))

    def set_power_iterations(self, max_iter: int):
        """Cap of the leading-pair power iteration (0 = default).  Channels that do not converge within it take the
        float64 fallback route; a cap of 1 forces that route (tests)."""
        self.check(self.lib.set_power_iterations(self._ctx, int(max_iter)))

    def profile(self, enable: bool):
        """Bracket every kernel launch group with CUDA events (bench.py's roofline leg)."""
        self.check(self.lib.profile_enable(self._ctx, 1 if enable else 0))

    def profile_read(self):
        """{kernel name: (total ms, timed launches)} since profile(True); synchronises."""
        n = self.lib.profile_count(self._ctx)
        return {self.lib.profile_name(self._ctx, i).decode(): (float(self.lib.profile_ms(self._ctx, i)),
                                                               int(self.lib.profile_calls(self._ctx, i)))
                for i in range(n)}

    def reserve(self, nbytes: int):
        self.check(self.lib.workspace_reserve(self._ctx, int(nbytes)))

    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def empty_image(self, B, rows, cols, dtype=torch.float32):
        """[B, rows, cols] view of a buffer whose rows are pitched to a multiple of 32 elements (128 bytes for float32):
        every row then starts on a cache line, 16-byte vector / bulk (TMA) accesses are legal, and the library takes
        its fast paths.  The view indexes like a dense array; .contiguous() / .cpu() give the dense copy."""
        pitch = (int(cols) + 31) // 32 * 32
        return torch.empty((B, rows, pitch), dtype=dtype, device=self.device)[:, :, :cols]

    def to_device(self, x, dtype=torch.float32):
        """-> (contiguous device tensor, came_from_torch)."""
        if _is_torch(x):
            return x.to(device=self.device, dtype=dtype).contiguous(), True
        a = np.ascontiguousarray(np.asarray(x))
        if a.dtype == object:
            raise TypeError("object arrays are not supported")
        t = torch.from_numpy(a)
        return t.to(device=self.device, dtype=dtype).contiguous(), False

    def to_device_image(self, x, dtype=torch.float32):
        """Like to_device for [..., rows, cols] images, but a device tensor whose rows are pitched (unit column stride,
        uniform row and plane pitch -- what empty_image hands out) is used in place instead of being densified."""
        if _is_torch(x) and x.device == self.device and x.dtype == dtype and x.dim() >= 2 and x.stride(-1) == 1:
            ok = x.stride(-2) >= x.shape[-1]
            for d in range(x.dim() - 2):
                ok = ok and x.stride(d) == x.stride(d + 1) * x.shape[d + 1]
            if ok:
                return x, True
        return self.to_device(x, dtype)

    @staticmethod
    def ret(t, as_torch):
        return t if as_torch else t.cpu().numpy()

    def plan(self, nperseg, noverlap, fs, window, scaling, detrend, eps=0.0):
        nperseg = int(nperseg)
        noverlap = int(noverlap)
        if nperseg < 1:
            raise ValueError("nperseg must be a positive integer")
        if noverlap >= nperseg:
            raise ValueError("noverlap must be less than nperseg.")      # scipy's message
        if isinstance(window, str):
            if window.lower() not in _ffi.WINDOW:
                raise ValueError(f"Unknown window type: {window!r}")
            wcode, warr, wkey = _ffi.WINDOW[window.lower()], None, window.lower()
        else:
            w = np.ascontiguousarray(np.asarray(window, dtype=np.float64))
            if w.ndim != 1:
                raise ValueError("window must be 1-D")
            if w.shape[0] != nperseg:
                raise ValueError("window must have length of nperseg")
            wcode, warr, wkey = 0, w, ("array", w.tobytes())
        if scaling not in _ffi.SCALING:
            raise ValueError(f"Unknown scaling: {scaling!r}")
        if isinstance(detrend, str) and detrend not in ("constant", "linear"):
            raise ValueError("Trend type must be 'linear' or 'constant'.")
        if callable(detrend):
            raise ValueError("callable detrend is not supported by libspecgpu")
        dcode = _ffi.DETREND[detrend if detrend else False]
        key = (nperseg, noverlap, float(fs), wkey, scaling, dcode, float(eps))
        with self._lock:
            h = self._plans.get(key)
            if h is None:
                p = _ffi.StftParams(nperseg, noverlap, dcode, _ffi.SCALING[scaling], wcode, 0, float(fs), float(eps))
                h = C.c_void_p()
                wp = warr.ctypes.data_as(C.POINTER(C.c_double)) if warr is not None else None
                self.check(self.lib.plan_create(self._ctx, C.byref(p), wp, C.byref(h)))
                self._plans[key] = h
        return h

    def plan_from_params(self, sp):
        return self.plan(sp["nperseg"], sp["noverlap"], sp["fs"], sp["window"], sp["scaling"], sp["detrend"],
                         sp.get("eps", 0.0))

    def axes(self, plan, n):
        F = self.lib.plan_num_freqs(plan)
        T = self.lib.plan_num_segments(plan, n)
        f = np.empty(F, np.float64)
        t = np.empty(max(T, 0), np.float64)
        self.check(self.lib.plan_axes(plan, n, f.ctypes.data_as(C.POINTER(C.c_double)),
                                      t.ctypes.data_as(C.POINTER(C.c_double))))
        return f, t

    # ---- raw entry points on device tensors (used by bench.py and the wrappers below) -------------
    def specgr_dev(self, plan, x2d, S=None, minmax=None):
        B, n = x2d.shape
        T = self.lib.plan_num_segments(plan, n)
        F = self.lib.plan_num_freqs(plan)
        if S is None:
            S = self.empty_image(B, F - 1, T)
        mmp = minmax.data_ptr() if minmax is not None else None
        self.check(self.lib.specgr(self._ctx, plan, x2d.data_ptr(), B, n, _ld(x2d), S.data_ptr(), _ld(S), mmp,
                                   self.stream()))
        return S

    def spectrogram_dev(self, plan, x2d, Sxx=None):
        B, n = x2d.shape
        T = self.lib.plan_num_segments(plan, n)
        F = self.lib.plan_num_freqs(plan)
        if Sxx is None:
            Sxx = self.empty_image(B, F, T)
        self.check(self.lib.spectrogram(self._ctx, plan, x2d.data_ptr(), B, n, _ld(x2d), Sxx.data_ptr(),
                                        _ld(Sxx), self.stream()))
        return Sxx

    def pipeline_dev(self, plan, x2d, S=None, D=None, clip=True, tiles=None, tile_w=128, ntiles=0, info=None,
                     fallback=True, static_tiles=False):
        """specgpu_pipeline on device tensors.  `fallback` (default on): channels whose leading singular pair did not
        converge in the power iteration are redone in the stream by the float64 eigensolver; with fallback=False the
        caller must look at info[:, 3] itself.  `static_tiles`: scheduling hint for calls that overlap with others on
        other streams (SPECGPU_PIPE_STATIC_TILES); results do not depend on it."""
        B, n = x2d.shape
        T = self.lib.plan_num_segments(plan, n)
        F = self.lib.plan_num_freqs(plan)
        if S is None:
            S = self.empty_image(B, F - 1, T)
        if D is None:
            D = self.empty_image(B, F - 1, T)
        flags = ((_ffi.PIPE_CLIP if clip else 0) | (_ffi.PIPE_FALLBACK if fallback else 0)
                 | (_ffi.PIPE_STATIC_TILES if static_tiles else 0))
        self.check(self.lib.pipeline(self._ctx, plan, x2d.data_ptr(), B, n, _ld(x2d), S.data_ptr(), D.data_ptr(),
                                     _ld(S), flags, tiles.data_ptr() if tiles is not None else None,
                                     tile_w, ntiles, info.data_ptr() if info is not None else None, self.stream()))
        return S, D


_defaults = {}
_default_lock = threading.Lock()


def default_runtime() -> Runtime:
    """The Runtime behind the module-level functions.  A libspecgpu context owns one scratch workspace and its calls are
    only ordered by the stream they are enqueued on, so there is one default Runtime per (host thread, device,
    current CUDA stream): work issued under different torch streams or from different threads never shares scratch."""
    if not torch.cuda.is_available():
        raise RuntimeError("libspecgpu needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device()
    key = (threading.get_ident(), dev, int(torch.cuda.current_stream(dev).cuda_stream))
    with _default_lock:
        rt = _defaults.get(key)
        if rt is None:
            rt = _defaults[key] = Runtime(device=torch.device("cuda", dev))
        return rt


def _rt(rt):
    return rt if rt is not None else default_runtime()


def _as_2d(t):
    """[..., N] -> ([B, N], leading shape)."""
    lead = tuple(t.shape[:-1])
    return t.reshape(-1, t.shape[-1]) if t.dim() != 2 else t, lead


# ================================================================================================
# scipy.signal call sites
# ================================================================================================
def spectrogram(x, fs=1.0, window="hann", nperseg=256, noverlap=None, detrend="constant", scaling="density",
                runtime=None):
    """scipy.signal.spectrogram(x, fs, window, nperseg, noverlap, detrend=, scaling=, mode='psd'),
    real input, one-sided, over the last axis.  Returns (f, t, Sxx[..., F, T]).
    (scipy's own defaults differ in two places: window=('tukey', .25) and nperseg=None.)"""
    rt = _rt(runtime)
    if noverlap is None:
        noverlap = int(nperseg) // 8
    plan = rt.plan(nperseg, noverlap, fs, window, scaling, detrend)
    xd, as_torch = rt.to_device(x)
    x2, lead = _as_2d(xd)
    f, t = rt.axes(plan, x2.shape[-1])
    Sxx = rt.spectrogram_dev(plan, x2)
    Sxx = Sxx.reshape(lead + tuple(Sxx.shape[1:]))
    return f, t, rt.ret(Sxx, as_torch)


def stft(x, fs=1.0, window="hann", nperseg=256, noverlap=None, detrend=False, boundary="zeros", padded=True,
         scaling="spectrum", runtime=None):
    """scipy.signal.stft (real input, one-sided).  Returns (f, t, Zxx[..., F, T]) complex64."""
    rt = _rt(runtime)
    if noverlap is None:
        noverlap = int(nperseg) // 2
    if boundary not in (None, "zeros"):
        raise ValueError(f"Unknown boundary option '{boundary}', must be one of: ['zeros', None]")
    if scaling == "psd":
        scaling = "density"
    plan = rt.plan(nperseg, noverlap, fs, window, scaling, detrend)
    xd, as_torch = rt.to_device(x)
    x2, lead = _as_2d(xd)
    B, n = x2.shape
    bz, pd = (1 if boundary == "zeros" else 0), (1 if padded else 0)
    T = rt.lib.stft_num_segments(plan, n, bz, pd)
    F = rt.lib.plan_num_freqs(plan)
    Z = rt.empty((B, F, T, 2))
    rt.check(rt.lib.stft(rt._ctx, plan, x2.data_ptr(), B, n, _ld(x2), bz, pd, Z.data_ptr(), T if T else 1, rt.stream()))
    Zc = torch.view_as_complex(Z).reshape(lead + (F, T))
    hop = int(nperseg) - int(noverlap)
    f = np.fft.rfftfreq(int(nperseg), 1.0 / fs)
    # scipy: time = arange(nperseg/2, padded_len - nperseg/2 + 1, hop)/fs, shifted by -nperseg/2/fs with a boundary
    t = (np.arange(T) * hop + int(nperseg) / 2) / float(fs)
    if boundary is not None:
        t = t - (int(nperseg) / 2) / float(fs)
    return f, t, rt.ret(Zc, as_torch)


def csd_allpairs(x, fs=1.0, window="hann", nperseg=256, noverlap=None, detrend="constant", scaling="density",
                 runtime=None):
    """All-pairs Welch cross-power spectrum of x[C, N]:
    P[i, j] = scipy.signal.csd(x[i], x[j], fs, window, nperseg, noverlap, detrend=, scaling=, average='mean').
    Returns (f, P[C, C, F]) complex64."""
    rt = _rt(runtime)
    if noverlap is None:
        noverlap = int(nperseg) // 2
    plan = rt.plan(nperseg, noverlap, fs, window, scaling, detrend)
    xd, as_torch = rt.to_device(x)
    if xd.dim() != 2:
        raise ValueError("csd_allpairs expects x[C, N]")
    Cn, n = xd.shape
    F = rt.lib.plan_num_freqs(plan)
    P = rt.empty((Cn, Cn, F, 2))
    rt.check(rt.lib.csd_allpairs(rt._ctx, plan, xd.data_ptr(), Cn, n, _ld(xd), P.data_ptr(), rt.stream()))
    f = np.fft.rfftfreq(int(nperseg), 1.0 / fs)
    return f, rt.ret(torch.view_as_complex(P), as_torch)


def csd(x, y, fs=1.0, window="hann", nperseg=256, noverlap=None, detrend="constant", scaling="density", runtime=None):
    """scipy.signal.csd(x, y, ...) for two real 1-D records (average='mean', one-sided):
    Pxy = mean over segments of conj(X) Y.  Returns (f, Pxy[F]) complex64."""
    rt = _rt(runtime)
    xd, as_torch = rt.to_device(x)
    yd, _ = rt.to_device(y)
    if xd.dim() != 1 or yd.dim() != 1:
        raise ValueError("csd expects 1-D x and y (use csd_allpairs for channel stacks)")
    if xd.shape != yd.shape:
        raise ValueError("x and y must have the same length")   # scipy zero-pads the shorter; not supported
    f, P = csd_allpairs(torch.stack([xd, yd]), fs, window, nperseg, noverlap, detrend, scaling, runtime=rt)
    return f, rt.ret(P[0, 1], as_torch)


# ================================================================================================
# spec_denoising/pipeline_data.py
# ================================================================================================
def spectrogram_batch(x, spec_params=DEFAULT_SPEC_PARAMS, runtime=None):
    """Body of `specgr` after the pickle slice (pipeline_data.py:32-35) for a stack of signals
    x[..., N]: spectrogram -> log(Sxx+eps) -> global min-max over all nperseg/2+1 rows (per signal)
    -> Nyquist row dropped.  Returns (Sxx[..., nperseg/2, T], f, t) in the reference's order."""
    rt = _rt(runtime)
    plan = rt.plan_from_params(spec_params)
    xd, as_torch = rt.to_device(x)
    x2, lead = _as_2d(xd)
    f, t = rt.axes(plan, x2.shape[-1])
    S = rt.specgr_dev(plan, x2)
    S = S.reshape(lead + tuple(S.shape[1:]))
    return rt.ret(S, as_torch), f[:-1], t


def specgr_array(sig_in, spec_params=DEFAULT_SPEC_PARAMS, runtime=None):
    return spectrogram_batch(sig_in, spec_params, runtime)


_SHOT_CACHE = {}


def _load_shot_pickle(fname):
    """The reference unpickles the whole shot once per channel (pipeline_data.py:29, 40 times per shot); keep the
    last file, keyed by path, mtime and size."""
    st = os.stat(fname)
    key = (os.path.abspath(fname), st.st_mtime_ns, st.st_size)
    if _SHOT_CACHE.get("key") != key:
        with open(fname, "rb") as fh:
            _SHOT_CACHE.update(key=key, data=pickle.load(fh))
    return _SHOT_CACHE["data"]


def _channel_signal(data, ecen, key="ece"):
    """One channel's raw record out of a shot pickle.  key='ece': data['\\tecefNN'] (pipeline_data.py:30);
    key='bes': data['besfuNN']['data.BES'] (the BES twin of specgr, denoising_by_svd.ipynb:50-52); or a callable
    key(data, ecen) -> 1-D array for other diagnostics."""
    if callable(key):
        return key(data, ecen)
    if key == "ece":
        return data["\\tecef%.2i" % (ecen)]
    if key == "bes":
        return data["besfu{:02d}".format(ecen)]["data.BES"]
    raise ValueError(f"unknown channel key {key!r}: use 'ece', 'bes' or a callable")


def load_shot(fname, channels=range(1, 41), cut_shot=2, fs=500000.0, key="ece"):
    """One read of a shot pickle -> x[C, N] float32, the batched input of `pipeline` / `spectrogram_batch`
    (channel keys and slice of pipeline_data.py:29-31; key='bes' for the BES pickles of denoising_by_svd.ipynb:50-52).
    The reference keeps the pickle's dtype; libspecgpu computes in float32, so a float64 record is rounded here."""
    data = _load_shot_pickle(fname)
    n = int(np.int_(cut_shot * fs))
    return np.stack([np.asarray(_channel_signal(data, c, key)[:n], dtype=np.float32) for c in channels])


def specgr(fname, ecen, spec_params, cut_shot=2, runtime=None, key="ece"):
    """pipeline_data.py:28-36, same signature: load the shot pickle (cached between calls), take channel `ecen`, the
    first cut_shot*fs samples, and return (Sxx[nperseg/2, T], f, t).  key='bes' is the notebook's BES variant
    (denoising_by_svd.ipynb:49-63: 'besfuNN' -> ['data.BES'])."""
    ece_data = _load_shot_pickle(fname)
    sig_in = _channel_signal(ece_data, ecen, key)[:np.int_(cut_shot * spec_params["fs"])]
    return spectrogram_batch(np.asarray(sig_in), spec_params, runtime)


def _matrix_batch(rt, data):
    """[rows, cols] or [B, rows, cols] -> (device [B, rows, cols], as_torch, squeeze)."""
    d, as_torch = rt.to_device(data)
    if d.dim() == 2:
        return d.unsqueeze(0), as_torch, True
    if d.dim() == 3:
        return d, as_torch, False
    raise ValueError("expected a 2-D matrix or a [B, rows, cols] stack")


def rescale(data, runtime=None):
    """pipeline_data.py:43-44: (data-min)/(max-min) over the whole array."""
    rt = _rt(runtime)
    d, as_torch = rt.to_device(data)
    flat = d.reshape(1, 1, -1)
    out = torch.empty_like(flat)
    rt.check(rt.lib.rescale(rt._ctx, flat.data_ptr(), 1, 1, flat.shape[-1], flat.shape[-1], out.data_ptr(), rt.stream()))
    return rt.ret(out.reshape(d.shape), as_torch)


def norm(data, runtime=None):
    """pipeline_data.py:38-41: (data-mean)/std (population std) over the whole array."""
    rt = _rt(runtime)
    d, as_torch = rt.to_device(data)
    flat = d.reshape(1, 1, -1)
    out = torch.empty_like(flat)
    rt.check(rt.lib.norm(rt._ctx, flat.data_ptr(), 1, 1, flat.shape[-1], flat.shape[-1], out.data_ptr(), rt.stream()))
    return rt.ret(out.reshape(d.shape), as_torch)


def _quantfilt_impl(rt, src, thr, want_mask):
    d, as_torch = rt.to_device(src)
    if d.dim() == 2:
        batch = d.unsqueeze(0)
    elif d.dim() == 3:
        # denoising_spectrogram.ipynb:115 passes [F, T, C]; np.quantile(axis=0) is per (t, c) column
        batch = d.permute(2, 0, 1).contiguous()
    else:
        raise ValueError("quantfilt expects [F, T] or [F, T, C]")
    B, rows, cols = batch.shape
    out = torch.empty_like(batch)
    mask = torch.empty(batch.shape, dtype=torch.uint8, device=rt.device) if want_mask else None
    thr_out = rt.empty((B, cols))
    rt.check(rt.lib.quantfilt(rt._ctx, batch.data_ptr(), B, rows, cols, cols, float(thr), out.data_ptr(),
                              thr_out.data_ptr(), mask.data_ptr() if want_mask else None, rt.stream()))
    if d.dim() == 2:
        out, thr_out = out[0], thr_out[0]
        mask = mask[0] if want_mask else None
    else:
        out, thr_out = out.permute(1, 2, 0), thr_out.permute(1, 0)
        mask = mask.permute(1, 2, 0) if want_mask else None
    return out, thr_out, mask, as_torch


def quantfilt(src, thr=0.9, runtime=None):
    """pipeline_data.py:46-49: zero everything strictly below the per-column `thr` quantile taken
    over axis 0 (np.quantile, linear interpolation; bit-exact for float32 input)."""
    rt = _rt(runtime)
    if not (0.0 <= float(thr) <= 1.0):
        raise ValueError("Quantiles must be in the range [0, 1]")      # numpy's message
    out, _, _, as_torch = _quantfilt_impl(rt, src, thr, False)
    return rt.ret(out, as_torch)


def quantfilt_mask(src, thr=0.9, runtime=None):
    """quantfilt plus its integer outputs: (out, threshold[T], mask uint8 (1 = kept))."""
    rt = _rt(runtime)
    if not (0.0 <= float(thr) <= 1.0):
        raise ValueError("Quantiles must be in the range [0, 1]")
    out, thr_out, mask, as_torch = _quantfilt_impl(rt, src, thr, True)
    return rt.ret(out, as_torch), rt.ret(thr_out, as_torch), rt.ret(mask, as_torch)


def _img_batch(rt, src):
    """[rows, cols] or [B, rows, cols], float32 or float64 -> (device tensor [B, rows, cols], as_torch, squeeze)."""
    as_torch = _is_torch(src)
    a = src if as_torch else np.asarray(src)
    f64 = (a.dtype == torch.float64) if as_torch else (a.dtype == np.float64)
    d, _ = rt.to_device(a, torch.float64 if f64 else torch.float32)
    if d.dim() == 2:
        return d.unsqueeze(0), as_torch, True
    if d.dim() == 3:
        return d, as_torch, False
    raise ValueError("expected a 2-D image or a [B, rows, cols] stack")


def gaussblr(src, filt=(31, 3), return_uint8=False, runtime=None):
    """pipeline_data.py:52-55: uint8-quantise, cv2.GaussianBlur(src, filt, 0), rescale -> float64.
    `filt` is cv2's ksize = (width along the last axis, height along rows).  The uint8 blur is bit-exact."""
    rt = _rt(runtime)
    d, as_torch, squeeze = _img_batch(rt, src)
    B, rows, cols = d.shape
    out = torch.empty((B, rows, cols), dtype=torch.float64, device=rt.device)
    u8 = torch.empty((B, rows, cols), dtype=torch.uint8, device=rt.device) if return_uint8 else None
    rt.check(rt.lib.gaussblr(rt._ctx, d.data_ptr(), 1 if d.dtype == torch.float64 else 0, B, rows, cols, _ld(d), int(filt[0]),
                             int(filt[1]), out.data_ptr(), cols, u8.data_ptr() if return_uint8 else None, rt.stream()))
    if squeeze:
        out, u8 = out[0], (u8[0] if return_uint8 else None)
    return (rt.ret(out, as_torch), rt.ret(u8, as_torch)) if return_uint8 else rt.ret(out, as_torch)


def meansub(src, runtime=None):
    """pipeline_data.py:58-61: rescale(|src - mean over time of each frequency row|) -> float64."""
    rt = _rt(runtime)
    as_torch = _is_torch(src)
    d, _ = rt.to_device(src, torch.float64)
    squeeze = d.dim() == 2
    if squeeze:
        d = d.unsqueeze(0)
    B, rows, cols = d.shape
    out = torch.empty_like(d)
    rt.check(rt.lib.meansub(rt._ctx, d.data_ptr(), B, rows, cols, _ld(d), out.data_ptr(), cols, rt.stream()))
    return rt.ret(out[0] if squeeze else out, as_torch)


def morph(src, return_uint8=False, runtime=None):
    """pipeline_data.py:64-72: uint8-quantise, MORPH_CLOSE rect(4,4), MORPH_OPEN rect(3,1), rescale -> float64."""
    rt = _rt(runtime)
    d, as_torch, squeeze = _img_batch(rt, src)
    B, rows, cols = d.shape
    out = torch.empty((B, rows, cols), dtype=torch.float64, device=rt.device)
    u8 = torch.empty((B, rows, cols), dtype=torch.uint8, device=rt.device) if return_uint8 else None
    rt.check(rt.lib.morph(rt._ctx, d.data_ptr(), 1 if d.dtype == torch.float64 else 0, B, rows, cols, _ld(d), out.data_ptr(),
                          cols, u8.data_ptr() if return_uint8 else None, rt.stream()))
    if squeeze:
        out, u8 = out[0], (u8[0] if return_uint8 else None)
    return (rt.ret(out, as_torch), rt.ret(u8, as_torch)) if return_uint8 else rt.ret(out, as_torch)


def filter_chain(Sxx, thr=0.9, filt=(31, 3), runtime=None, fused=True):
    """The denoising pipeline of pipeline_data.py:101-110 on a spectrogram (or a [B, F, T] stack), kept on the device
    between stages: quantfilt -> gaussblr -> meansub -> morph -> meansub.  Returns `pipeline_out` (float64).
    `fused=True` is one library call (specgpu_filter_chain: uint8 planes between the stages); `fused=False` chains the
    five public functions - the results are identical bit for bit."""
    rt = _rt(runtime)
    as_torch = _is_torch(Sxx)
    d, _ = rt.to_device_image(Sxx)
    squeeze = d.dim() == 2
    if squeeze:
        d = d.unsqueeze(0)
    if not 0.0 <= float(thr) <= 1.0:
        raise ValueError("Quantiles must be in the range [0, 1]")
    B, rows, cols = d.shape
    if fused:
        out = torch.empty((B, rows, cols), dtype=torch.float64, device=rt.device)
        rt.check(rt.lib.filter_chain(rt._ctx, d.data_ptr(), B, rows, cols, _ld(d), float(thr), int(filt[0]), int(filt[1]),
                                     out.data_ptr(), cols, rt.stream()))
    else:
        q = torch.empty_like(d)
        rt.check(rt.lib.quantfilt(rt._ctx, d.data_ptr(), B, rows, cols, _ld(d), float(thr), q.data_ptr(), None, None, rt.stream()))
        out = meansub(morph(meansub(gaussblr(q, filt, runtime=rt), runtime=rt), runtime=rt), runtime=rt)
    if squeeze:
        out = out[0]
    return rt.ret(out, as_torch)


# ================================================================================================
# spec_denoising/denoising_by_svd.ipynb
# ================================================================================================
def omega(beta):
    """denoising_by_svd.ipynb:155-159 (host scalar)."""
    return 0.56 * beta ** 3 - 0.95 * beta ** 2 + 1.82 * beta + 1.43


def _svd_call(rt, matrix, kind, start, stop, use_optimal, clip_neg, method, want_s):
    m, as_torch, squeeze = _matrix_batch(rt, matrix)
    transposed = m.shape[1] > m.shape[2]
    if transposed:                       # A^T = V S U^T: the same singular-index range, transposed back below
        m = m.transpose(1, 2).contiguous()
    B, rows, cols = m.shape
    info = torch.empty((B, 4), dtype=torch.int32, device=rt.device)
    s_out = rt.empty((B, rows)) if want_s else None
    sp = s_out.data_ptr() if want_s else None
    if kind == 2:
        out = torch.empty((B, rows, cols), dtype=torch.float64, device=rt.device)
        rt.check(rt.lib.compute_signal(rt._ctx, m.data_ptr(), B, rows, cols, cols, out.data_ptr(), cols, sp,
                                       info.data_ptr(), rt.stream()))
    else:
        out = rt.empty((B, rows, cols))
        rt.check(rt.lib.svd_denoise(rt._ctx, m.data_ptr(), B, rows, cols, cols, int(start), int(stop),
                                    1 if use_optimal else 0, 1 if clip_neg else 0, 1 if method == "jacobi" else 0,
                                    out.data_ptr(), cols, sp, info.data_ptr(), rt.stream()))
    if transposed:
        out = out.transpose(1, 2).contiguous()
    if squeeze:
        out = out[0]
        info = info[0]
        s_out = s_out[0] if want_s else None
    return out, s_out, info, as_torch


def denoiseSignal(matrix, start=None, stop=None, use_optimal=False, clip=False, method="auto", return_info=False,
                  runtime=None):
    """denoising_by_svd.ipynb:188-229: U[:, start:stop] diag(s[start:stop]) Vh[start:stop, :] of the thin
    SVD (defaults start=1, stop=len(s): drop the leading component); `use_optimal` keeps
    start=0, stop=num_sing-1 with num_sing = #(s > omega(beta)*median(s)).
    Extras (keyword-only in spirit): clip=True fuses the notebook's `hacked[hacked<0]=0` (:280-281);
    method='jacobi' forces the full eigen-decomposition; return_info=True also returns
    (s[len], (start, stop, num_sing)) with the reference's integer bookkeeping."""
    rt = _rt(runtime)
    shape = tuple(matrix.shape)
    k = min(shape[-2], shape[-1])
    a = 1 if start is None else int(start)
    b = k if stop is None else int(stop)
    out, s_out, info, as_torch = _svd_call(rt, matrix, 1 if use_optimal else 0, a, b, use_optimal, clip, method,
                                           return_info)
    if not return_info:
        return rt.ret(out, as_torch)
    return rt.ret(out, as_torch), rt.ret(s_out, as_torch), info.cpu().numpy()


def computeSignal(matrix, return_info=False, runtime=None):
    """denoising_by_svd.ipynb:161-186: sum over idx in range(1, 2*num_sing) of s[idx] u_idx v_idx^T,
    float64 output.  Like the reference it raises IndexError when 2*num_sing-1 >= len(s)."""
    rt = _rt(runtime)
    out, s_out, info, as_torch = _svd_call(rt, matrix, 2, 0, 0, False, False, "jacobi", True)
    inf = info.cpu().numpy().reshape(-1, 4)
    k = min(matrix.shape[-2], matrix.shape[-1])
    for row in inf:
        if 2 * int(row[2]) - 1 >= k and int(row[2]) > 0:
            raise IndexError(f"index {k} is out of bounds for axis 0 with size {k}")
    if not return_info:
        return rt.ret(out, as_torch)
    return rt.ret(out, as_torch), rt.ret(s_out, as_torch), info.cpu().numpy()


def clip(x, runtime=None):
    """denoising_by_svd.ipynb:280-281: copy with negatives set to 0 (fused into denoiseSignal(clip=True)
    and `pipeline` on the hot path)."""
    rt = _rt(runtime)
    d, as_torch = rt.to_device(x)
    out = torch.empty_like(d)
    rt.check(rt.lib.clip(rt._ctx, d.data_ptr(), d.numel(), out.data_ptr(), rt.stream()))
    return rt.ret(out, as_torch)


# ================================================================================================
# VAE/manual_scan.py tile cut
# ================================================================================================
def patch(arr, tile=128, ntiles=30, dtype=np.float64, runtime=None):
    """VAE/manual_scan.py:28-36: list (or stack) of [rows, >= tile*ntiles] spectrograms ->
    [len(arr)*ntiles, rows, tile]; float64 like the reference's np.empty (dtype=np.float32 halves the
    bytes for the VAE input)."""
    rt = _rt(runtime)
    as_torch = _is_torch(arr) or (len(arr) > 0 and _is_torch(arr[0]))
    if isinstance(arr, (list, tuple)):
        if len(arr) == 0:
            out = torch.empty((0, 0, tile), dtype=torch.float64)
            return out if as_torch else out.numpy()
        width = tile * ntiles
        parts = [rt.to_device(a)[0][:, :width] for a in arr]
        d = torch.stack(parts)
    else:
        d, _ = rt.to_device(arr)
    n, rows, ld = d.shape
    if ld < tile * ntiles:
        raise ValueError(f"could not broadcast input array: need at least {tile * ntiles} columns, got {ld}")
    f64 = np.dtype(dtype) == np.float64
    out = torch.empty((n * ntiles, rows, tile), dtype=torch.float64 if f64 else torch.float32, device=rt.device)
    rt.check(rt.lib.patch(rt._ctx, d.data_ptr(), n, rows, _ld(d), tile, ntiles, out.data_ptr(), 1 if f64 else 0,
                          rt.stream()))
    return rt.ret(out, as_torch)


def unpatch(arr, ntiles=30, runtime=None):
    """VAE/manual_scan.py:39-49: [n*ntiles, rows, tile] -> [n, rows, tile*ntiles] (dtype kept)."""
    rt = _rt(runtime)
    as_torch = _is_torch(arr)
    a = arr if as_torch else np.asarray(arr)
    f64 = (a.dtype == torch.float64) if as_torch else (a.dtype == np.float64)
    d, _ = rt.to_device(a, torch.float64 if f64 else torch.float32)
    total, rows, tile = d.shape
    n = int(total / ntiles)
    out = torch.empty((n, rows, tile * ntiles), dtype=d.dtype, device=rt.device)
    rt.check(rt.lib.unpatch(rt._ctx, d.data_ptr(), 1 if f64 else 0, n, rows, tile, ntiles, out.data_ptr(), 1 if f64 else 0,
                            tile * ntiles, rt.stream()))
    return rt.ret(out, as_torch)


def reshape(arr):
    """VAE/manual_scan.py:52-54."""
    if _is_torch(arr):
        return arr.reshape(len(arr), 256, 128, 1)
    return np.reshape(arr, (len(arr), 256, 128, 1))


# ================================================================================================
# the whole path, batched
# ================================================================================================
def pipeline(x, spec_params=DEFAULT_SPEC_PARAMS, clip=True, tiles=False, tile=128, return_info=False, runtime=None):
    """x[B, N] -> (S, D[, tiles]): S = specgr body, D = denoiseSignal(S) (default range: leading
    component removed) with the clip fused, tiles = patch(D) as float32 [B*(T//tile), rows, tile]."""
    rt = _rt(runtime)
    plan = rt.plan_from_params(spec_params)
    xd, as_torch = rt.to_device(x)
    x2, lead = _as_2d(xd)
    B, n = x2.shape
    T = rt.lib.plan_num_segments(plan, n)
    rows = rt.lib.plan_num_freqs(plan) - 1
    ntiles = T // tile if tiles else 0
    tl = rt.empty((B * ntiles, rows, tile)) if tiles else None
    info = torch.empty((B, 4), dtype=torch.int32, device=rt.device)
    S, D = rt.pipeline_dev(plan, x2, clip=clip, tiles=tl, tile_w=tile, ntiles=ntiles, info=info)
    S = S.reshape(lead + tuple(S.shape[1:]))
    D = D.reshape(lead + tuple(D.shape[1:]))
    res = [rt.ret(S, as_torch), rt.ret(D, as_torch)]
    if tiles:
        res.append(rt.ret(tl, as_torch))
    if return_info:
        res.append(info.cpu().numpy())
    return tuple(res)


def process_shot(x, spec_params=DEFAULT_SPEC_PARAMS, thr=0.9, filt=(31, 3), channels=range(1, 41), cut_shot=2, runtime=None,
                 key="ece"):
    """One shot of the reference's main loop (pipeline_data.py:92-116) in two library calls: `x` is the shot pickle's
    path (read once with `load_shot`) or the signals x[C, N].  Returns the datasets the loop writes per channel, stacked
    over channels: {'spec': Sxx[C, F, T], 'f': f, 't': t, 'pipeline_out': [C, F, T] float64}."""
    rt = _rt(runtime)
    if isinstance(x, (str, os.PathLike)):
        x = load_shot(x, channels=channels, cut_shot=cut_shot, fs=spec_params["fs"], key=key)
    plan = rt.plan_from_params(spec_params)
    xd, as_torch = rt.to_device(x)
    if xd.dim() != 2:
        raise ValueError("process_shot expects x[C, N] or a pickle path")
    f, t = rt.axes(plan, xd.shape[-1])
    S = rt.specgr_dev(plan, xd)                                  # stays on the device between the two calls
    out = filter_chain(S, thr, filt, runtime=rt)
    return {"spec": rt.ret(S, as_torch), "f": f[:-1], "t": t, "pipeline_out": rt.ret(out, as_torch)}


def _h5py():
    try:
        import h5py
    except ImportError as e:      # the interchange file is optional: everything else works without h5py
        raise ImportError("the HDF5 interchange (save_shot_hdf5 / load_hdf5_dataset) needs h5py") from e
    return h5py


def save_shot_hdf5(out_file, shotn, result, channels=None):
    """Write one shot in the reference's interchange layout (pipeline_data.py:112-116): group
    'ece_<shotn>/chn_<n>' with datasets 'spec', 'f', 't', 'pipeline_out' per channel.  `result` is what `process_shot`
    returns ({'spec': [C, F, T], 'f', 't', 'pipeline_out': [C, F, T]}); `channels` are the 1-based channel numbers
    (default 1..C); `out_file` is an open h5py.File (or Group) or a path, opened in 'a' mode like the reference.
    Unlike the reference (whose create_group raises on a second run) an existing channel group is replaced."""
    def cpu(a):
        return a.cpu().numpy() if _is_torch(a) else np.asarray(a)
    spec, out = cpu(result["spec"]), cpu(result["pipeline_out"])
    f, t = np.asarray(result["f"]), np.asarray(result["t"])
    if spec.ndim != 3 or out.shape != spec.shape:
        raise ValueError("expected result['spec'] and result['pipeline_out'] as [C, F, T]")
    chans = list(range(1, spec.shape[0] + 1)) if channels is None else list(channels)
    if len(chans) != spec.shape[0]:
        raise ValueError("one channel number per row of result['spec']")
    own = isinstance(out_file, (str, os.PathLike))
    fh = _h5py().File(out_file, "a") if own else out_file
    try:
        for i, chn in enumerate(chans):
            name = "ece_" + str(shotn) + "/chn_" + str(chn)
            if name in fh:
                del fh[name]
            grp = fh.create_group(name)
            grp.create_dataset("spec", data=spec[i])
            grp.create_dataset("f", data=f)
            grp.create_dataset("t", data=t)
            grp.create_dataset("pipeline_out", data=out[i])
    finally:
        if own:
            fh.close()


def load_hdf5_dataset(in_file, shots=None, n_channels=20):
    """Read (spectrograms, final) lists the way the VAE scripts do (VAE/manual_scan.py:137-148): for every shot group
    (all of them, or the names in `shots`) and channel 1..n_channels, file[shot/'chn_<n>']['spec'] and ['pipeline_out'].
    Returns two lists of [F, T] arrays, ready for `patch`."""
    own = isinstance(in_file, (str, os.PathLike))
    fh = _h5py().File(in_file, "r") if own else in_file
    try:
        names = list(fh.keys()) if shots is None else list(shots)
        spectrograms, final = [], []
        for fname in names:
            for chn in range(n_channels):
                name = fname + "/chn_" + str(chn + 1)
                if name not in fh:
                    continue
                spectrograms.append(np.array(fh[name]["spec"]))
                final.append(np.array(fh[name]["pipeline_out"]))
        return spectrograms, final
    finally:
        if own:
            fh.close()


class HostPipeline:
    """The whole path for shots that live in HOST memory: x[C, N] (pinned) -> D[C, rows, T] (pinned), optionally S.

    The channels of a shot are cut into `groups`; group g goes through staging set g % streams as
    H2D(x_g) on ONE upload stream -> specgpu_pipeline on the set's compute stream -> D2H(D_g) on ONE download stream,
    chained by events.  The two copy engines therefore run back to back (PCIe is full duplex) and never wait for each
    other's stream: with upload, kernels and download of a group on the same stream (the first version of this class) the
    next upload of a set queued behind its previous download although the upload engine was idle (3.6 ms per 40-channel
    shot; the duplex copies alone take 3.24).  Every staging set owns a libspecgpu context (its own workspace) and its
    own device buffers; nothing is allocated after construction."""

    def __init__(self, spec_params=DEFAULT_SPEC_PARAMS, channels=40, samples=1_000_000, groups=8, streams=3, clip=True,
                 want_S=False, device=None, lib=None):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("libspecgpu needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.clip, self.want_S = clip, want_S
        groups = max(1, min(groups, channels))
        streams = max(1, min(streams, groups))
        bounds = np.linspace(0, channels, groups + 1).astype(int)
        self.ranges = [(int(bounds[i]), int(bounds[i + 1])) for i in range(groups) if bounds[i + 1] > bounds[i]]
        gmax = max(b - a for a, b in self.ranges)
        self.rts = [Runtime(lib=lib, device=self.device) for _ in range(streams)]
        self.plans = [rt.plan_from_params(spec_params) for rt in self.rts]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(streams)]      # compute, one per staging set
        self.up = torch.cuda.Stream(device=self.device)
        self.down = torch.cuda.Stream(device=self.device)
        self.ev_up = [torch.cuda.Event() for _ in range(streams)]        # the set's input has arrived
        self.ev_comp = [torch.cuda.Event() for _ in range(streams)]      # its kernels are done (input buffer free, outputs ready)
        self.ev_down = [torch.cuda.Event() for _ in range(streams)]      # its outputs have left (output buffers free)
        self._k = 0
        rt0 = self.rts[0]
        self.T = int(rt0.lib.plan_num_segments(self.plans[0], samples))
        self.rows = int(rt0.lib.plan_num_freqs(self.plans[0])) - 1
        self.xd = [rt0.empty((gmax, samples)) for _ in range(streams)]
        self.Sd = [rt0.empty_image(gmax, self.rows, self.T) for _ in range(streams)]     # row-pitched: the fast paths
        self.Dd = [rt0.empty_image(gmax, self.rows, self.T) for _ in range(streams)]
        self.channels, self.samples = channels, samples
        for i, rt in enumerate(self.rts):      # grow the workspaces now (growing synchronises the device)
            with torch.cuda.stream(self.streams[i]):
                rt.pipeline_dev(self.plans[i], self.xd[i].zero_(), self.Sd[i], self.Dd[i], clip=clip)
        torch.cuda.synchronize(self.device)

    def launch_count(self):
        return sum(rt.launch_count() for rt in self.rts)

    class _Done:
        """Completion handle of one submitted shot: a CUDA event behind its last download."""

        def __init__(self, events):
            self.events = events

        def synchronize(self):
            for e in self.events:
                e.synchronize()

    def _download(self, i, dst_host, src_dev):
        """Pitched device image [n, rows, T] -> dense host array, one strided DMA copy on the download stream."""
        rt = self.rts[i]
        n = src_dev.shape[0]
        if not dst_host.is_contiguous():
            raise ValueError("host result buffers must be contiguous")
        rt.check(rt.lib.copy_rows(rt._ctx, dst_host.data_ptr(), self.T * 4, src_dev.data_ptr(), int(src_dev.stride(1)) * 4,
                                  self.T * 4, n * self.rows, C.c_void_p(self.down.cuda_stream)))

    def submit(self, x_host, D_host, S_host=None, copy_only=False):
        """Enqueue one shot (uploads, kernels, downloads) and return a handle whose synchronize() returns when its last
        download has completed.  The shot is ordered only after earlier shots of this object, so shots submitted back to
        back overlap (the downloads of one run under the uploads of the next); x_host must already be final (host
        memory), and D_host must not be read before synchronize()."""
        if tuple(x_host.shape) != (self.channels, self.samples):
            raise ValueError(f"expected x[{self.channels}, {self.samples}], got {tuple(x_host.shape)}")
        for a, b in self.ranges:
            i = self._k % len(self.streams)
            self._k += 1
            n = b - a
            comp = self.streams[i]
            with torch.cuda.stream(self.up):
                self.up.wait_event(self.ev_comp[i])          # the set's previous kernels have read its input buffer
                self.xd[i][:n].copy_(x_host[a:b], non_blocking=True)
                self.ev_up[i].record(self.up)
            with torch.cuda.stream(comp):
                comp.wait_event(self.ev_up[i])
                comp.wait_event(self.ev_down[i])             # the set's previous outputs have been downloaded
                if not copy_only:       # copy_only: the same transfers without the kernels (bench.py's copy ceiling)
                    self.rts[i].pipeline_dev(self.plans[i], self.xd[i][:n], self.Sd[i][:n], self.Dd[i][:n], clip=self.clip,
                                             static_tiles=len(self.streams) > 1)
                self.ev_comp[i].record(comp)
            with torch.cuda.stream(self.down):
                self.down.wait_event(self.ev_comp[i])
                self._download(i, D_host[a:b], self.Dd[i][:n])
                if S_host is not None:
                    self._download(i, S_host[a:b], self.Sd[i][:n])
                self.ev_down[i].record(self.down)
        e = torch.cuda.Event()
        e.record(self.down)
        return HostPipeline._Done([e])

    def run(self, x_host, D_host, S_host=None):
        """x_host [C, N] float32 (pinned for real overlap), D_host [C, rows, T] float32 pinned; returns after the
        last download has completed."""
        self.submit(x_host, D_host, S_host).synchronize()
        return D_host


class ShotStreams:
    """The shot loop of the reference (spec_denoising/pipeline_data.py:92-97) for shots that already sit in DEVICE
    memory, `n` shots in flight: shot i runs on worker stream i % n with that stream's own libspecgpu context (its own
    workspace), so the latency-bound middle of one shot (Gram reduce + power iteration + the repair check, ~30 us during
    which most SMs idle) runs under the STFT / projection of its neighbour.  Measured on one B200, config 2:
    0.277 -> 0.258 ms per 40-channel shot with two shots in flight (three: 0.260).

    submit() orders the shot after everything already enqueued on the CALLER's current stream (so inputs produced there
    are visible) and after the previous shot on the same worker stream; join() makes the caller's stream wait for every
    shot submitted so far."""

    def __init__(self, spec_params=DEFAULT_SPEC_PARAMS, n=2, device=None, lib=None, interlock=False):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("libspecgpu needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        n = max(1, int(n))
        self.rts = [Runtime(lib=lib, device=self.device) for _ in range(n)]
        self.plans = [rt.plan_from_params(spec_params) for rt in self.rts]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n)]
        self._next = 0
        self._dirty = [False] * n
        # interlock (specgpu_set_pipeline_interlock): the STFT of shot i + 1 starts when shot i reaches its projection
        self._ilk = []
        if interlock and n > 1:
            self._ilk = [torch.cuda.Event() for _ in range(n)]
            for e in self._ilk:
                e.record(torch.cuda.current_stream(self.device))      # creates the CUDA event; complete at once
            for i, rt in enumerate(self.rts):
                rt.check(rt.lib.set_pipeline_interlock(rt._ctx, C.c_void_p(self._ilk[(i - 1) % n].cuda_event),
                                                       C.c_void_p(self._ilk[i].cuda_event)))

    def launch_count(self):
        return sum(rt.launch_count() for rt in self.rts)

    def empty_image(self, B, rows, cols):
        return self.rts[0].empty_image(B, rows, cols)

    def submit(self, x2d, S, D, **kw):
        """Runtime.pipeline_dev(plan, x2d, S, D, **kw) on the next worker stream; returns that stream's index."""
        i = self._next
        self._next = (i + 1) % len(self.streams)
        st = self.streams[i]
        st.wait_stream(torch.cuda.current_stream(self.device))
        kw.setdefault("static_tiles", len(self.streams) > 1)      # overlapping shots: the fixed tile walk overlaps better
        with torch.cuda.stream(st):
            self.rts[i].pipeline_dev(self.plans[i], x2d, S, D, **kw)
        self._dirty[i] = True
        return i

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for i, st in enumerate(self.streams):
            if self._dirty[i]:
                cur.wait_stream(st)
                self._dirty[i] = False


# ================================================================================================
# interferometer/crosspowerspec.py
# ================================================================================================
def ae_co2(signal1, signal2, t, nperseg=1024, noverlap=None, navg=8, window="hann", detrend="constant", runtime=None):
    """interferometer/crosspowerspec.py:39 call signature: (signal1, signal2, t[ms]) ->
    (ampsp[n_time, n_freq], freq[kHz], time[ms]).

    The body (`co2_deps.ae_co2`) is not in the reference tree; per the project brief its arithmetic is
    the segment-averaged cross-power spectrum conj(X1) X2 (scipy.signal.csd, mean).  Here the record is
    cut into consecutive frames of `navg` half-overlapped segments and ampsp[k] = |csd(frame k)|, so
    that ampsp is the time-resolved cross-power amplitude the script plots (PARITY-UNPINNED: frame
    length and overlap are this library's choice)."""
    rt = _rt(runtime)
    if noverlap is None:
        noverlap = nperseg // 2
    hop = nperseg - noverlap
    s1, as_torch = rt.to_device(signal1)
    s2, _ = rt.to_device(signal2)
    tt = np.asarray(t.cpu() if _is_torch(t) else t, dtype=np.float64)
    fs = 1000.0 / float(np.mean(np.diff(tt)))        # t in ms -> Hz
    frame = nperseg + (navg - 1) * hop
    if frame % hop != 0:
        raise ValueError("ae_co2: nperseg must be a multiple of nperseg - noverlap")
    seg_stride = frame // hop          # frame k = segments [k*seg_stride, k*seg_stride + navg) of the whole record
    n = int(s1.shape[-1])
    nframes = n // frame
    if nframes == 0:
        raise ValueError("record shorter than one averaging frame")
    if s2.shape != s1.shape or s1.dim() != 1:
        raise ValueError("ae_co2 expects two 1-D records of equal length")
    F = nperseg // 2 + 1
    plan = rt.plan(nperseg, noverlap, fs, window, "density", detrend)
    x = torch.stack([s1, s2])
    T = rt.lib.plan_num_segments(plan, n)
    ldf = (F + 1) & ~1
    X = rt.empty((2, T, ldf, 2))
    rt.check(rt.lib.csd_spectra(rt._ctx, plan, x.data_ptr(), 2, n, _ld(x), X.data_ptr(), ldf, rt.stream()))
    amps = rt.empty((nframes, F))
    rt.check(rt.lib.csd_frames(rt._ctx, plan, X.data_ptr(), 2, T, ldf, 0, 1, seg_stride, navg, nframes, amps.data_ptr(),
                               rt.stream()))
    freq = np.fft.rfftfreq(nperseg, 1.0 / fs) / 1e3
    time = tt[0] + (np.arange(nframes) * frame + frame / 2.0) / fs * 1e3
    return rt.ret(amps, as_torch), freq, time
