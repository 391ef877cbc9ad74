"""Pin the CPU oracle (oracle/spec_oracle.py): scipy's own known-answer vectors, scipy/numpy at
run time, the golden vectors produced by the reference's own functions, and (when the reference
tree is present, i.e. in the build container) the reference functions themselves."""
import numpy as np
import pytest
import scipy.signal

from oracle import ref_loader, spec_oracle as oc


# ---- scipy KATs (scipy/signal/tests/test_spectral.py, TestCSD / TestSpectrogram) ---------------
def _impulse16():
    x = np.zeros(16)
    x[0] = 1
    x[8] = 1
    return x


def test_kat_csd_real_onesided_even():
    f, p = oc.csd(_impulse16(), _impulse16(), nperseg=8)
    np.testing.assert_allclose(f, np.linspace(0, 0.5, 5))
    q = np.array([0.08333333, 0.15277778, 0.22222222, 0.22222222, 0.11111111])
    np.testing.assert_allclose(p, q, atol=1e-7, rtol=1e-7)


def test_kat_csd_real_onesided_odd():
    x = _impulse16()
    f, p = oc.csd(x, x, nperseg=9)
    np.testing.assert_allclose(f, np.arange(5.0) / 9.0)
    q = np.array([0.12477455, 0.23430933, 0.17072113, 0.17072113, 0.17072113])
    np.testing.assert_allclose(p, q, atol=1e-7, rtol=1e-7)


def test_kat_csd_real_spectrum():
    x = _impulse16()
    f, p = oc.csd(x, x, nperseg=8, scaling="spectrum")
    q = np.array([0.015625, 0.02864583, 0.04166667, 0.04166667, 0.02083333])
    np.testing.assert_allclose(p, q, atol=1e-7, rtol=1e-7)


def test_kat_csd_detrend_linear():
    x = np.arange(10, dtype=np.float64) + 0.04
    f, p = oc.csd(x, x, nperseg=10, detrend="linear")
    np.testing.assert_allclose(p, np.zeros_like(p), atol=1e-15)


def test_kat_spectrogram_average_all_segments():
    # TestSpectrogram.test_average_all_segments: mean over segments == welch
    x = np.random.default_rng(0).standard_normal(1024)
    f, t, P = oc.spectrogram(x, 1.0, ("hann"), 16, 2)
    fw, Pw = scipy.signal.welch(x, 1.0, "hann", 16, 2)
    np.testing.assert_allclose(f, fw)
    np.testing.assert_allclose(P.mean(axis=-1), Pw, rtol=1e-12)


def test_bad_noverlap_raises():
    with pytest.raises(ValueError):
        oc.spectrogram(np.zeros(64), nperseg=8, noverlap=8)
    with pytest.raises(ValueError):
        oc.csd(np.zeros(64), np.zeros(64), nperseg=8, scaling="foo")


# ---- restatement vs scipy.signal at run time -----------------------------------------------------
@pytest.mark.parametrize("nperseg,noverlap", [(512, 256), (1024, 512), (256, 32), (8, 4), (64, 63), (100, 37)])
@pytest.mark.parametrize("window", ["hann", "hamm", "boxcar"])
@pytest.mark.parametrize("detrend", [False, "constant", "linear"])
def test_spectrogram_matches_scipy(nperseg, noverlap, window, detrend):
    x = oc.synth_ece(1, 2, n=5000).astype(np.float64) + 3.0
    for scaling in ("density", "spectrum"):
        f0, t0, P0 = scipy.signal.spectrogram(x, fs=500000, window=window, nperseg=nperseg,
                                              noverlap=noverlap, detrend=detrend, scaling=scaling)
        f1, t1, P1 = oc.spectrogram(x, fs=500000, window=window, nperseg=nperseg, noverlap=noverlap,
                                    detrend=detrend, scaling=scaling)
        assert P0.shape == P1.shape
        assert np.array_equal(f0, f1) and np.array_equal(t0, t1)      # index-derived: exact
        np.testing.assert_allclose(P1, P0, rtol=1e-9, atol=1e-12 * P0.max())


@pytest.mark.parametrize("boundary,padded", [("zeros", True), (None, True), (None, False), ("zeros", False)])
def test_stft_matches_scipy(boundary, padded):
    x = oc.synth_ece(1, 5, n=10_000).astype(np.float64)
    f0, t0, Z0 = scipy.signal.stft(x, fs=500000, window="hann", nperseg=1024, noverlap=512,
                                   boundary=boundary, padded=padded)
    f1, t1, Z1 = oc.stft(x, fs=500000, window="hann", nperseg=1024, noverlap=512,
                         boundary=boundary, padded=padded)
    assert Z0.shape == Z1.shape
    np.testing.assert_allclose(t1, t0, rtol=0, atol=1e-15)
    assert np.array_equal(f0, f1)
    np.testing.assert_allclose(Z1, Z0, rtol=1e-9, atol=1e-12 * np.abs(Z0).max())


def test_stft_config1_shape():
    f, t, Z = oc.stft(np.zeros(1_000_000, np.float32), fs=500000, nperseg=1024, noverlap=512, dtype=np.float32)
    assert Z.shape == (513, 1955)
    f, t, P = oc.spectrogram(np.zeros(1_000_000, np.float32), fs=500000, nperseg=1024, noverlap=512, dtype=np.float32)
    assert P.shape == (513, 1952)


@pytest.mark.parametrize("detrend", [False, "constant", "linear"])
def test_csd_matches_scipy(detrend):
    x = oc.synth_ece(2, 0, n=40_000, fs=1.6e6).astype(np.float64)
    y = oc.synth_ece(2, 1, n=40_000, fs=1.6e6).astype(np.float64)
    f0, P0 = scipy.signal.csd(x, y, fs=1.6e6, window="hann", nperseg=4096, noverlap=2048, detrend=detrend)
    f1, P1 = oc.csd(x, y, fs=1.6e6, window="hann", nperseg=4096, noverlap=2048, detrend=detrend)
    assert np.array_equal(f0, f1)
    np.testing.assert_allclose(P1, P0, rtol=1e-9, atol=1e-12 * np.abs(P0).max())
    X = np.stack([x, y, x[::-1].copy()])
    f2, PP = oc.csd_allpairs(X, fs=1.6e6, window="hann", nperseg=4096, noverlap=2048, detrend=detrend)
    np.testing.assert_allclose(PP[0, 1], P0, rtol=1e-9, atol=1e-12 * np.abs(P0).max())
    np.testing.assert_allclose(PP[1, 0], np.conj(P0), rtol=1e-9, atol=1e-12 * np.abs(P0).max())
    _, P22 = scipy.signal.csd(X[2], X[2], fs=1.6e6, window="hann", nperseg=4096, noverlap=2048, detrend=detrend)
    np.testing.assert_allclose(PP[2, 2].real, P22, rtol=1e-9)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [256, 257, 10, 513])
@pytest.mark.parametrize("thr", [0.9, 0.5, 0.99, 0.123, 0.0, 1.0])
def test_quantile_lerp_bit_exact(dt, n, thr):
    a = np.random.default_rng(n).random((n, 333)).astype(dt)
    assert np.array_equal(np.quantile(a, thr, axis=0), oc.quantile_lerp(a, thr))


# ---- golden vectors from the reference's own functions ------------------------------------------
def test_golden_specgr_small(golden):
    g = golden("specgr_small.npz")
    sp = dict(oc.DEFAULT_SPEC_PARAMS)
    S64, f, t = oc.specgr_array(g["x"].astype(np.float64), sp)
    assert S64.shape == g["S_f64"].shape == (256, 77)
    assert np.array_equal(f, g["f_f64"]) and np.array_equal(t, g["t_f64"])
    np.testing.assert_allclose(S64, g["S_f64"], rtol=0, atol=1e-10)
    # the reference's native-f32 run differs from f64 only at the 1e-5 level after log+normalise
    assert np.abs(g["S_f32"] - S64).max() < 1e-4
    S32, _, _ = oc.specgr_array(g["x"], sp, dtype=np.float32)
    assert S32.dtype == np.float32 and g["S_f32"].dtype == np.float32
    assert np.abs(S32 - g["S_f32"]).max() < 1e-4


def test_golden_specgr_full_columns(golden):
    import zlib
    g = golden("specgr_full_cols.npz")
    x = oc.synth_ece(0, 0)
    assert np.uint32(zlib.crc32(x.tobytes())) == g["x_crc"]
    S, f, t = oc.specgr_array(x.astype(np.float64))
    assert tuple(g["shape"]) == S.shape == (256, 3905)
    assert np.array_equal(f, g["f"]) and np.array_equal(t, g["t"])
    np.testing.assert_allclose(S[:, g["cols"]], g["S_f64"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(S.sum(axis=1), g["rowsum_f64"], rtol=1e-10)


def test_golden_quantfilt_norm_rescale(golden):
    g = golden("specgr_small.npz")
    S = g["S_f32"]
    assert np.array_equal(oc.quantfilt(S, 0.9), g["quant_f32"])
    assert np.array_equal(oc.quantile_lerp(S, 0.9), g["quant_thr_f32"])
    assert np.array_equal(oc.quantfilt(g["S_f64"], 0.9), g["quant_f64"])
    # float32 mean/std depend on numpy's summation order (the reference summed a strided view)
    np.testing.assert_allclose(oc.norm(S), g["norm_f32"], rtol=0, atol=2e-5)
    assert np.array_equal(oc.rescale(S * 3 - 1), g["rescale_f32"])


def test_golden_svd(golden):
    g = golden("svd_small.npz")
    np.testing.assert_allclose(oc.omega(g["omega_beta"]), g["omega"], rtol=1e-15)
    small = golden("specgr_small.npz")
    mats = {"M32": g["M"], "M64": g["M"].astype(np.float64), "S32": small["S_f32"]}
    for tag, mat in mats.items():
        tol = 2e-4 if mat.dtype == np.float32 else 1e-9
        scale = np.abs(mat).max()
        for key, kw in (("default", {}), ("optimal", dict(use_optimal=True)), ("0_4", dict(start=0, stop=4)),
                        ("2_9", dict(start=2, stop=9)), ("m3_1000", dict(start=-3, stop=1000))):
            np.testing.assert_allclose(oc.denoiseSignal(mat, **kw), g[f"{tag}_{key}"], rtol=0, atol=tol * scale)
        out = oc.computeSignal(mat)
        assert out.dtype == np.float64
        np.testing.assert_allclose(out, g[f"{tag}_compute"], rtol=0, atol=tol * scale)


def test_svd_plan_quirks():
    s = np.array([10.0, 5.0, 1.0, 0.9, 0.8, 0.7])
    assert oc.svd_plan((6, 60), s) == (1, 6, -1)
    assert oc.svd_plan((6, 60), s, -3, 1000) == (0, 6, -1)
    a, b, ns = oc.svd_plan((6, 60), s, use_optimal=True)
    assert (a, b) == (0, ns - 1)
    # num_sing == 0 -> stop = -1 (python slice keeps all but the last component)
    s0 = np.ones(6)
    assert oc.svd_plan((6, 60), s0, use_optimal=True) == (0, -1, 0)


def test_patch_unpatch_roundtrip():
    rng = np.random.default_rng(3)
    arr = [rng.random((256, 3905)).astype(np.float32) for _ in range(3)]
    p = oc.patch(arr)
    assert p.shape == (90, 256, 128) and p.dtype == np.float64
    assert np.array_equal(p[31], arr[1][:, 128:256].astype(np.float64))
    u = oc.unpatch(p)
    assert u.shape == (3, 256, 3840)
    assert np.array_equal(u[2], arr[2][:, :3840].astype(np.float64))
    assert oc.reshape(p).shape == (90, 256, 128, 1)


# ---- against the reference itself (build container only) ----------------------------------------
@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_against_reference_functions():
    import os, pickle, tempfile
    ref = ref_loader.load_pipeline_data()
    nbk = ref_loader.load_svd_notebook()
    x = oc.synth_ece(5, 7, n=30_000).astype(np.float64)
    with tempfile.NamedTemporaryFile(suffix="_1.pkl", delete=False) as fh:
        pickle.dump({"\\tecef08": x}, fh)
    try:
        S0, f0, t0 = ref.specgr(fh.name, 8, dict(oc.DEFAULT_SPEC_PARAMS), 30_000 / 500000)
    finally:
        os.unlink(fh.name)
    S1, f1, t1 = oc.specgr_array(x)
    assert np.array_equal(f0, f1) and np.array_equal(t0, t1)
    np.testing.assert_allclose(S1, S0, rtol=0, atol=1e-10)
    assert np.array_equal(ref.quantfilt(S0, 0.8), oc.quantfilt(S0, 0.8))
    M = oc.synth_lowrank(32, 200, [50.0, 20.0, 5.0], 0.01, seed=1).astype(np.float64)
    for kw in ({}, dict(use_optimal=True), dict(start=0, stop=2)):
        np.testing.assert_allclose(oc.denoiseSignal(M, **kw), nbk["denoiseSignal"](M, **kw), atol=1e-10)
    np.testing.assert_allclose(oc.computeSignal(M), nbk["computeSignal"](M), atol=1e-10)
    assert oc.omega(0.3) == nbk["omega"](0.3) or abs(oc.omega(0.3) - nbk["omega"](0.3)) < 1e-15


# ---- cv2 chain restatement vs OpenCV itself and vs the reference's golden outputs ----------------
def test_cv2_chain_restatement_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(0)     # OpenCV's own thread pool once returned a different blur for a 9x9 image late in a long suite run
    rng = np.random.default_rng(0)
    for n in (3, 5, 7, 9, 15, 31, 41):          # Q8.8 taps: read them off a constant-column probe image
        img = np.zeros((9, 4 * n + 1), np.uint8)
        img[:, 2 * n] = 255
        taps = cv2.GaussianBlur(img, (n, 3), 0)[4, 2 * n - n // 2: 2 * n + n // 2 + 1].astype(int)
        assert np.array_equal(taps, oc.gaussian_taps_q8(n))
    se1 = cv2.getStructuringElement(cv2.MORPH_RECT, (4, 4))
    se2 = cv2.getStructuringElement(cv2.MORPH_RECT, (3, 1))
    for shape in ((256, 390), (17, 40), (5, 33), (9, 9)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        for ks in ((31, 3), (5, 5), (9, 3)):
            if shape[1] >= ks[0]:
                assert np.array_equal(cv2.GaussianBlur(img, ks, 0), oc.gaussian_blur_u8(img, ks))
        m = cv2.morphologyEx(cv2.morphologyEx(img, cv2.MORPH_CLOSE, se1), cv2.MORPH_OPEN, se2)
        assert np.array_equal(m, oc.morph_close_open_u8(img))


def test_cv2_chain_golden(golden):
    """quantfilt -> gaussblr -> meansub -> morph -> meansub of the reference itself (pipeline_data.py:101-110)."""
    g = golden("specgr_small.npz")
    gauss = oc.gaussblr(g["quant_f32"], (31, 3))
    assert np.array_equal(gauss, g["gauss"])
    mean = oc.meansub(gauss)
    assert np.array_equal(mean, g["mean"])
    mo = oc.morph(mean)
    assert np.array_equal(mo, g["morph"])
    assert np.array_equal(oc.meansub(mo), g["final"])
    assert np.array_equal(oc.filter_chain(g["S_f32"]), g["final"])
