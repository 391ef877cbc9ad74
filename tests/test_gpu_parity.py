"""Parity tests proper: libspecgpu.so (CUDA, sm_100a) through the C ABI / host API on cuda:0 versus the
CPU oracle on the same seeded inputs, the committed golden vectors produced by the reference's own
functions, and size-independent properties at BASELINE.json's full sizes.  Run with `-m gpu`."""
import zlib

import numpy as np
import pytest

import parity_cases as pc
from oracle import spec_oracle as oc
from spectrogram_enhancement_b200 import api

pytestmark = pytest.mark.gpu

SP = oc.DEFAULT_SPEC_PARAMS


# ---- K1/K2a: spectrogram front-end -----------------------------------------------------------------
@pytest.mark.parametrize("nperseg", [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_spectrogram_every_size(cuda_rt, nperseg):
    pc.case_spectrogram(cuda_rt, nperseg, nperseg // 2, max(40 * nperseg, 50_000), "linear", "hamm", "density", B=3)


@pytest.mark.parametrize("nperseg,noverlap,n,detrend,window,scaling", [
    (512, 256, 100_000, "constant", "hann", "spectrum"),
    (512, 63, 50_001, False, "boxcar", "density"),
    (64, 63, 7_000, "linear", "hann", "density"),
    (256, 0, 20_480, "constant", "hamming", "density"),
    (1024, 1000, 30_001, "linear", "hann", "density"),
])
def test_spectrogram_variants(cuda_rt, nperseg, noverlap, n, detrend, window, scaling):
    pc.case_spectrogram(cuda_rt, nperseg, noverlap, n, detrend, window, scaling)


def test_spectrogram_property_random_parameters(cuda_rt):
    """random (n, nperseg, noverlap, window, detrend, scaling, batch) against the oracle."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(log2n=st.integers(3, 13), ov=st.floats(0.0, 0.97), extra=st.integers(0, 5000), nseg=st.integers(1, 40),
           window=st.sampled_from(["hann", "hamm", "boxcar"]), detrend=st.sampled_from([False, "constant", "linear"]),
           scaling=st.sampled_from(["density", "spectrum"]), B=st.integers(1, 3))
    def run(log2n, ov, extra, nseg, window, detrend, scaling, B):
        nperseg = 1 << log2n
        noverlap = min(int(ov * nperseg), nperseg - 1)
        n = nperseg + (nseg - 1) * (nperseg - noverlap) + extra % (nperseg - noverlap)
        pc.case_spectrogram(cuda_rt, nperseg, noverlap, n, detrend, window, scaling, B=B)

    run()


def test_spectrogram_custom_window(cuda_rt):
    pc.case_spectrogram(cuda_rt, 128, 64, 30_000, "constant", np.hanning(130)[1:-1], "density")


def test_specgr_reference_defaults(cuda_rt):
    pc.case_specgr(cuda_rt, SP, 200_000, B=3)


def test_specgr_golden_small(cuda_rt, golden):
    """specgr() of the reference itself (pipeline_data.py:28-36) on 20 000 samples."""
    g = golden("specgr_small.npz")
    S, f, t = api.spectrogram_batch(g["x"], SP, runtime=cuda_rt)
    np.testing.assert_allclose(S, g["S_f64"], rtol=0, atol=pc.ATOL_IMAGE)
    np.testing.assert_allclose(S, g["S_f32"], rtol=0, atol=pc.ATOL_IMAGE)
    assert np.array_equal(f, g["f_f64"]) and np.array_equal(t, g["t_f64"])
    # quantfilt of the reference on its own float32 spectrogram: integer mask bit-exact
    out, thr, mask = api.quantfilt_mask(g["S_f32"], 0.9, runtime=cuda_rt)
    assert np.array_equal(thr, g["quant_thr_f32"]) and np.array_equal(out, g["quant_f32"])
    np.testing.assert_allclose(api.norm(g["S_f32"], runtime=cuda_rt), g["norm_f32"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(api.rescale(g["S_f32"] * 3 - 1, runtime=cuda_rt), g["rescale_f32"], rtol=0, atol=1e-6)


def test_specgr_golden_full_channel(cuda_rt, golden):
    """Config 1/2 size: the reference's specgr on the full 1 000 000-sample synthetic channel (shot 0, ch 0)."""
    g = golden("specgr_full_cols.npz")
    x = oc.synth_ece(0, 0)
    assert np.uint32(zlib.crc32(x.tobytes())) == g["x_crc"]
    S, f, t = api.spectrogram_batch(x, SP, runtime=cuda_rt)
    assert S.shape == tuple(g["shape"]) == (256, 3905)
    assert np.array_equal(f, g["f"]) and np.array_equal(t, g["t"])
    np.testing.assert_allclose(S[:, g["cols"]], g["S_f64"], rtol=0, atol=pc.ATOL_IMAGE)
    np.testing.assert_allclose(S.astype(np.float64).sum(axis=1), g["rowsum_f64"], rtol=1e-5)


def test_config1_stft_and_spectrogram(cuda_rt):
    """BASELINE config 1: single channel, 1M samples @500 kHz, nperseg=1024 hann 50% overlap."""
    x = oc.synth_ece(1, 0)
    f, t, Z = api.stft(x, fs=500000, window="hann", nperseg=1024, noverlap=512, runtime=cuda_rt)
    fr, tr, Zr = oc.stft(x.astype(np.float64), fs=500000, window="hann", nperseg=1024, noverlap=512)
    assert Z.shape == (513, 1955) and np.array_equal(f, fr)
    np.testing.assert_allclose(t, tr, rtol=0, atol=1e-12)
    pc.assert_spec_close(Z, Zr)
    f, t, P = api.spectrogram(x, fs=500000, window="hann", nperseg=1024, noverlap=512, runtime=cuda_rt)
    _, _, Pr = oc.spectrogram(x.astype(np.float64), fs=500000, window="hann", nperseg=1024, noverlap=512)
    assert P.shape == (513, 1952)
    pc.assert_spec_close(P, Pr)


def test_config1_rank_k_denoise_512_rows(cuda_rt):
    """Config 1's SVD leg: rank-k truncation of a 512-row image with a planted gap (cluster-of-8 Jacobi)."""
    pc.case_svd_range(cuda_rt, 512, 1952, [900.0, 300.0, 120.0], 0, 3, noise=0.05, seed=11)


@pytest.mark.parametrize("boundary,padded", [("zeros", True), (None, True), (None, False), ("zeros", False)])
def test_stft_boundaries(cuda_rt, boundary, padded):
    pc.case_stft(cuda_rt, 256, 128, 30_001, boundary, padded, B=2)


def test_spectrogram_linearity_and_welch_identity(cuda_rt):
    """Size-independent properties at full size: PSD(2x) = 4 PSD(x); mean over segments == csd diagonal."""
    x = pc.signals(2, 1_000_000)
    f, t, P1 = api.spectrogram(x, fs=500000, window="hann", nperseg=512, noverlap=256, detrend="constant", runtime=cuda_rt)
    _, _, P2 = api.spectrogram(2 * x, fs=500000, window="hann", nperseg=512, noverlap=256, detrend="constant", runtime=cuda_rt)
    np.testing.assert_allclose(P2, 4 * P1, rtol=2e-6, atol=1e-12)
    _, C = api.csd_allpairs(x, fs=500000, window="hann", nperseg=512, noverlap=256, runtime=cuda_rt)
    np.testing.assert_allclose(np.stack([C[0, 0].real, C[1, 1].real]), P1.astype(np.float64).mean(axis=-1), rtol=2e-5)


# ---- helpers, K4 ------------------------------------------------------------------------------------
def test_rescale_norm(cuda_rt):
    pc.case_rescale_norm(cuda_rt, (256, 3905))
    pc.case_rescale_norm(cuda_rt, (7,))


@pytest.mark.parametrize("rows,cols,thr", [(256, 3905, 0.9), (257, 1000, 0.9), (100, 333, 0.5), (33, 65, 0.123),
                                           (64, 64, 0.0), (64, 64, 1.0), (513, 700, 0.99), (1024, 129, 0.75)])
def test_quantfilt(cuda_rt, rows, cols, thr):
    pc.case_quantfilt(cuda_rt, rows, cols, thr)


def test_quantfilt_ties_and_3d(cuda_rt):
    pc.case_quantfilt(cuda_rt, 256, 500, 0.9, ties=True)
    pc.case_quantfilt_3d(cuda_rt, 256, 3905, 8)          # denoising_spectrogram.ipynb:115 layout [F, T, C]


def test_patch_roundtrip(cuda_rt):
    pc.case_patch(cuda_rt, 4, 256, 3905, 128, 30)        # the reference's hard-wired 30 x (256, 128)
    pc.case_patch(cuda_rt, 3, 16, 70, 8, 8)


# ---- K5: cv2 image chain --------------------------------------------------------------------------------
def test_cv2_chain_golden(cuda_rt, golden):
    """quantfilt -> gaussblr -> meansub -> morph -> meansub of the reference itself (pipeline_data.py:101-110)."""
    g = golden("specgr_small.npz")
    pc.case_filter_chain(cuda_rt, g["S_f32"])
    assert np.array_equal(api.gaussblr(g["quant_f32"], (31, 3), runtime=cuda_rt), g["gauss"])
    pc.assert_same_f64(api.meansub(g["gauss"], runtime=cuda_rt), g["mean"])
    assert np.array_equal(api.morph(g["mean"], runtime=cuda_rt), g["morph"])
    pc.assert_same_f64(api.filter_chain(g["S_f32"], runtime=cuda_rt), g["final"])


def test_cv2_chain_full_size(cuda_rt):
    S, _, _ = api.spectrogram_batch(oc.synth_ece(3, 1), SP, runtime=cuda_rt)       # [256, 3905]
    pc.case_filter_chain(cuda_rt, S)
    Sb = np.stack([S, S[::-1].copy()])
    out = api.gaussblr(Sb, (31, 3), runtime=cuda_rt)
    assert np.array_equal(out[1], oc.gaussblr(Sb[1], (31, 3)))


# ---- K3: SVD denoise --------------------------------------------------------------------------------
def test_svd_golden(cuda_rt, golden):
    """denoiseSignal / computeSignal / omega of the reference notebook on a planted-gap matrix."""
    g = golden("svd_small.npz")
    M = g["M"]
    np.testing.assert_allclose([api.omega(b) for b in g["omega_beta"]], g["omega"], rtol=1e-15)
    pc.assert_denoise_close(api.denoiseSignal(M, runtime=cuda_rt), g["M64_default"])
    pc.assert_denoise_close(api.denoiseSignal(M, method="jacobi", runtime=cuda_rt), g["M64_default"])
    d, s, info = api.denoiseSignal(M, use_optimal=True, return_info=True, runtime=cuda_rt)
    pc.assert_denoise_close(d, g["M64_optimal"])
    np.testing.assert_allclose(s, g["M64_s"], rtol=1e-4, atol=1e-5 * g["M64_s"][0])
    pc.assert_denoise_close(api.denoiseSignal(M, 0, 4, runtime=cuda_rt), g["M64_0_4"])
    pc.assert_denoise_close(api.denoiseSignal(M, 2, 9, runtime=cuda_rt), g["M64_2_9"])
    pc.assert_denoise_close(api.denoiseSignal(M, -3, 1000, runtime=cuda_rt), g["M64_m3_1000"])
    # computeSignal sums idx in range(1, 2*num_sing) = 1..7: the cut falls INSIDE the noise bulk (s[7] - s[8] = 2e-3 on
    # s[0] = 300).  The full route forms the Gram matrix and runs the Jacobi sweeps in float64 for exactly this case.
    cs, s2, info2 = api.computeSignal(M, return_info=True, runtime=cuda_rt)
    pc.assert_denoise_close(cs, g["M64_compute"])
    np.testing.assert_allclose(s2, g["M64_s"], rtol=2e-6)
    # the reference's own spectrogram (256 x 77: taller than wide, handled through the transpose)
    S = golden("specgr_small.npz")["S_f32"]
    pc.assert_denoise_close(api.denoiseSignal(S, runtime=cuda_rt), g["S64_default"])
    pc.assert_denoise_close(api.denoiseSignal(S, method="jacobi", runtime=cuda_rt), g["S64_default"])
    d, s3, info3 = api.denoiseSignal(S, use_optimal=True, return_info=True, runtime=cuda_rt)
    np.testing.assert_allclose(s3, g["S64_s"], rtol=1e-5, atol=1e-6 * g["S64_s"][0])
    pc.assert_denoise_close(d, g["S64_optimal"])


@pytest.mark.parametrize("rows,cols", [(64, 200), (128, 1000), (256, 3905), (200, 777)])
def test_svd_default(cuda_rt, rows, cols):
    pc.case_svd_default(cuda_rt, rows, cols, [50.0 * rows ** 0.5, 20.0, 10.0])
    pc.case_svd_default(cuda_rt, rows, cols, [50.0 * rows ** 0.5, 20.0, 10.0], clip=True)


def test_svd_tensor_core_gram_matches_simt(cuda_rt):
    """The tcgen05 TF32 Gram route (method='auto', rows 256) and the fp32 SIMT + Jacobi route agree."""
    m = oc.synth_lowrank(256, 3905, [700.0, 9.0, 8.0], 0.1, 21) + np.float32(0.4)
    a = api.denoiseSignal(m, runtime=cuda_rt)
    b = api.denoiseSignal(m, method="jacobi", runtime=cuda_rt)
    ref = oc.denoiseSignal(m.astype(np.float64))
    pc.assert_denoise_close(a, ref)
    pc.assert_denoise_close(b, ref)


def test_svd_range_optimal_compute(cuda_rt):
    pc.case_svd_range(cuda_rt, 256, 3905, [400.0, 200.0, 100.0, 50.0], 0, 3)
    pc.case_svd_range(cuda_rt, 32, 80, [40, 20, 10, 5], 1, -28)
    pc.case_svd_optimal(cuda_rt, 256, 3905, [400.0, 200.0, 100.0], noise=0.05)
    pc.case_svd_optimal(cuda_rt, 100, 333, [40, 20, 10])
    # a (nearly) repeated singular value inside the kept range: the values-first solver's inverse iteration has to
    # Gram-Schmidt the pair apart (or hand the matrix to the Jacobi solver)
    pc.case_svd_optimal(cuda_rt, 128, 600, [60.0, 60.0, 25.0, 10.0], noise=0.02)
    pc.case_svd_optimal(cuda_rt, 256, 1200, [300.0, 299.9999, 80.0, 79.0, 20.0], noise=0.02, seed=9)
    pc.case_compute_signal(cuda_rt, 64, 300, [40, 20])


@pytest.mark.gpu
def test_svd_optimal_values_first_route_alone(cuda_rt, monkeypatch):
    # SPECGPU_TRIDIAG_STRICT drops the Jacobi fallback: the values-first solver (cluster tridiagonalisation, bisection,
    # inverse iteration, back-transformation) has to carry these by itself, cluster sizes 1 (n <= 154), 2 and 3
    monkeypatch.setenv("SPECGPU_TRIDIAG_STRICT", "1")
    pc.case_svd_optimal(cuda_rt, 256, 3905, [400.0, 200.0, 100.0], noise=0.05)
    pc.case_svd_optimal(cuda_rt, 200, 900, [100.0, 50.0, 20.0], noise=0.03, seed=2)
    pc.case_svd_optimal(cuda_rt, 100, 333, [40, 20, 10])
    pc.case_svd_optimal(cuda_rt, 128, 600, [60.0, 60.0, 25.0, 10.0], noise=0.02)
    pc.case_svd_optimal(cuda_rt, 256, 1200, [300.0, 299.9999, 80.0, 79.0, 20.0], noise=0.02, seed=9)
    pc.case_svd_optimal(cuda_rt, 3, 40, [5.0, 2.0], noise=0.05)


def test_svd_batched_and_tall(cuda_rt):
    ms = np.stack([oc.synth_lowrank(128, 500, [300.0, 20.0], 0.05, 30 + i) for i in range(5)])
    d = api.denoiseSignal(ms, runtime=cuda_rt)
    for i in range(5):
        pc.assert_denoise_close(d[i], oc.denoiseSignal(ms[i].astype(np.float64)))
    m = oc.synth_lowrank(300, 64, [40, 20, 10], 0.05, 9)
    pc.assert_denoise_close(api.denoiseSignal(m, runtime=cuda_rt), oc.denoiseSignal(m.astype(np.float64)))


def test_cv2_chain_tiles_and_degenerate(cuda_rt):
    """Fused blur (16 x 512 tiles, IDP.4A) and morphology (32 x 256 tiles) across tile seams, ragged edges, the
    unpacked kw = 1 path, wide kernels and images smaller than the kernel; uint8 bit-exact."""
    pc.case_cv2_tiles(cuda_rt, [(256, 3905), (70, 1100), (33, 257)], [(31, 3), (5, 7), (1, 1)])
    pc.case_cv2_tiles(cuda_rt, [(300, 700)], [(63, 31), (3, 101)])
    pc.case_cv2_tiles(cuda_rt, [(3, 9), (1, 40), (40, 1), (2, 2)], [(1, 1), (3, 1), (9, 3), (31, 3)])
    pc.case_meansub_wide(cuda_rt)
    pc.case_cv2_many_rows(cuda_rt)
    pc.case_cv2_pitched(cuda_rt)
    pc.case_cv2_division_corners(cuda_rt)
    pc.case_cv2_many_rows(cuda_rt, (7, 333, 290))


def test_filter_chain_fused_full_size(cuda_rt):
    """specgpu_filter_chain on a [3, 256, 3905] stack (config 2's image size): bit-identical to chaining the five public
    calls, and within float64 round-off of the oracle per image."""
    rng = np.random.default_rng(21)
    S = rng.random((3, 256, 3905)).astype(np.float32) ** 4
    fused = api.filter_chain(S, runtime=cuda_rt)
    assert fused.dtype == np.float64 and fused.shape == S.shape
    assert np.array_equal(fused, api.filter_chain(S, runtime=cuda_rt, fused=False))
    for i in (0, 2):
        assert np.array_equal(fused[i], api.filter_chain(S[i], runtime=cuda_rt))          # batched == one by one
    pc.assert_same_f64(fused[1], oc.filter_chain(S[1]))


# ---- K2b: cross-power spectrum -------------------------------------------------------------------------
def test_config3_co2_csd(cuda_rt):
    """BASELINE config 3: 4 chords, 2 s @ 1.6 MHz, nperseg 4096, all pairs."""
    pc.case_csd(cuda_rt, 4, 3_200_000, 4096)


@pytest.mark.parametrize("nperseg", [256, 512, 1024, 2048, 4096, 8192])
def test_config5_nperseg_sweep_40ch(cuda_rt, nperseg):
    """BASELINE config 5 (single-GPU leg): 40-channel all-pairs CSD across the nperseg sweep."""
    pc.case_csd(cuda_rt, 40, 120_000, nperseg, fs=500000.0)


def test_csd_wide_stacks_and_row_blocks(cuda_rt):
    """8x4 tiles with folded warps (9..20 channels), the staged kernel with ragged tiles (27, 64 channels), and row
    blocks accumulated over segment blocks (the sharded call pattern)."""
    pc.case_csd(cuda_rt, 9, 60_000, 256)
    pc.case_csd(cuda_rt, 20, 60_000, 128)
    pc.case_csd(cuda_rt, 27, 50_000, 256)
    pc.case_csd(cuda_rt, 64, 40_000, 512)
    pc.case_csd(cuda_rt, 40, 200, 64)
    pc.case_csd(cuda_rt, 40, 64, 64)
    pc.case_csd(cuda_rt, 33, 100, 8)
    pc.case_csd_row_block(cuda_rt, 40, 120_000, 1024, 8, 24, nblocks=3)
    pc.case_csd_row_block(cuda_rt, 40, 120_000, 1024, 20, 20, nblocks=4)
    pc.case_csd_row_block(cuda_rt, 6, 30_000, 64, 2, 3, nblocks=2)


def test_csd_variants_and_kat(cuda_rt):
    pc.case_csd(cuda_rt, 5, 30_000, 64, detrend="linear", scaling="spectrum")
    pc.case_csd(cuda_rt, 3, 30_000, 512, detrend=False, window="hamm")
    x = np.zeros(16, np.float32)
    x[0] = 1
    x[8] = 1
    f, p = api.csd(x, x, nperseg=8, runtime=cuda_rt)     # scipy TestCSD.test_real_onesided_even
    np.testing.assert_allclose(p.real, [0.08333333, 0.15277778, 0.22222222, 0.22222222, 0.11111111], rtol=1e-5)


def test_ae_co2_time_resolved(cuda_rt):
    """interferometer/crosspowerspec.py:39 call shape: 2 s @ 1.6 MHz, frames of 8 half-overlapped 1024-sample segments."""
    pc.case_ae_co2(cuda_rt, 3_200_000, 1024, 8)
    pc.case_ae_co2(cuda_rt, 100_000, 4096, 3)


# ---- whole path ----------------------------------------------------------------------------------------
def test_pipeline_small(cuda_rt):
    pc.case_pipeline(cuda_rt, dict(SP, nperseg=32, noverlap=16), 9000, B=2, tile=64)
    pc.case_pipeline(cuda_rt, dict(SP, nperseg=256, noverlap=128), 60_000, B=3, tile=128)
    # tile widths that are not a multiple of the projection's 32-column CTA tile (lanes of a warp straddle two VAE tiles)
    pc.case_pipeline(cuda_rt, dict(SP, nperseg=256, noverlap=128), 60_000, B=2, tile=50)
    pc.case_pipeline(cuda_rt, dict(SP, nperseg=512, noverlap=256), 80_000, B=2, tile=33)


def test_config2_pipeline_40ch(cuda_rt):
    """BASELINE config 2: one shot, 40 ECE channels x 1M samples, reference defaults.  Oracle on 6 of
    the 40 channels (seconds each); invariants on all 40."""
    x = pc.signals(40, 1_000_000, shot=7)
    S, D, tiles, info = api.pipeline(x, SP, clip=True, tiles=True, return_info=True, runtime=cuda_rt)
    assert S.shape == D.shape == (40, 256, 3905) and tiles.shape == (40 * 30, 256, 128)
    assert np.array_equal(info[:, :2], np.tile([1, 256], (40, 1))) and (info[:, 3] == 0).all()
    assert (D >= 0).all() and S.min() >= 0 and S.max() <= 1
    assert np.array_equal(tiles, oc.patch(list(D), 128, 30).astype(np.float32))
    for c in (0, 1, 13, 26, 38, 39):
        Sr, _, _ = oc.specgr_array(x[c].astype(np.float64), SP)
        np.testing.assert_allclose(S[c], Sr, rtol=0, atol=pc.ATOL_IMAGE)
        pc.assert_denoise_close(D[c], oc.clip(oc.denoiseSignal(S[c].astype(np.float64))))
    # batch invariance: channel 5 alone gives the same bits as inside the batch of 40
    S5, D5 = api.pipeline(x[5:6], SP, clip=True, runtime=cuda_rt)
    assert np.array_equal(S5[0], S[5])
    np.testing.assert_allclose(D5[0], D[5], rtol=0, atol=2e-5)


def test_config2_denoise_vs_pure_oracle_chain(cuda_rt):
    """Config 2 at full size, end to end: D of the GPU pipeline against the oracle's own chain
    clip(denoiseSignal(specgr(x))) in float64 (NOT the oracle applied to the GPU's S), at the north star's tolerance
    for the denoised reconstruction (rtol 1e-3, atol 1e-3 max|D|), on 3 of the 40 channels."""
    x = pc.signals(40, 1_000_000, shot=11)
    S, D = api.pipeline(x, SP, clip=True, runtime=cuda_rt)
    for c in (2, 19, 37):
        Sr, _, _ = oc.specgr_array(x[c].astype(np.float64), SP)
        pc.assert_denoise_close(D[c], oc.clip(oc.denoiseSignal(Sr)))


def test_pipeline_fallback_route(cuda_rt):
    pc.case_pipeline_fallback(cuda_rt, dict(SP, nperseg=256, noverlap=128), 60_000, B=3)
    pc.case_pipeline_fallback(cuda_rt, SP, 300_000, B=4)       # 256 rows: the 4-CTA cluster float64 Jacobi


def test_svd_degenerate_leading_pair_and_null_start(cuda_rt):
    pc.case_svd_degenerate(cuda_rt, rows=64, cols=300)
    pc.case_svd_degenerate(cuda_rt, rows=256, cols=1500)
    pc.case_svd_degenerate(cuda_rt, rows=128, cols=777)


def test_clip(cuda_rt):
    """denoising_by_svd.ipynb:280-281 as a standalone call: negatives (and -0.0) -> 0, NaN stays NaN, any shape."""
    rng = np.random.default_rng(2)
    a = rng.standard_normal((3, 257, 131)).astype(np.float32)
    a[0, 0, :4] = [np.nan, -0.0, 0.0, -np.inf]
    got = api.clip(a, runtime=cuda_rt)
    assert np.array_equal(got, oc.clip(a), equal_nan=True)
    assert api.clip(np.zeros((0,), np.float32), runtime=cuda_rt).shape == (0,)
    one = api.clip(np.float32([-1.5]), runtime=cuda_rt)
    assert one.tolist() == [0.0]


def test_host_pipeline_streams(cuda_rt):
    """HostPipeline (pinned host in/out, channel groups on several streams, shots in flight) == api.pipeline."""
    import torch
    n, C = 300_000, 12
    xs = [torch.from_numpy(pc.signals(C, n, shot=20 + i)).pin_memory() for i in range(3)]
    hp = api.HostPipeline(SP, channels=C, samples=n, groups=5, streams=3, clip=True, want_S=True)
    outs = [torch.empty((C, hp.rows, hp.T)).pin_memory() for _ in range(3)]
    Ss = torch.empty((C, hp.rows, hp.T)).pin_memory()
    hp.run(xs[0], outs[0], Ss)
    handles = [hp.submit(xs[i], outs[i]) for i in (1, 2)]
    for h in handles:
        h.synchronize()
    for i in range(3):
        S, D = api.pipeline(xs[i].numpy(), SP, clip=True, runtime=cuda_rt)
        np.testing.assert_allclose(outs[i].numpy(), D, rtol=0, atol=2e-5)
        if i == 0:
            assert np.array_equal(Ss.numpy(), S)
    with pytest.raises(ValueError):
        hp.run(xs[0][:3], outs[0])


@pytest.mark.parametrize("interlock", [False, True])
def test_shot_streams_device_resident(cuda_rt, interlock):
    """ShotStreams (device-resident shots, several in flight on worker streams, with and without the STFT <-> projection
    interlock of specgpu_set_pipeline_interlock) == Runtime.pipeline_dev, bit for bit, including the tiles and the info
    rows (which the projection kernel copies itself), for more shots than streams."""
    import torch
    n, C = 300_000, 6
    dev = cuda_rt.device
    xs = [torch.from_numpy(pc.signals(C, n, shot=30 + i)).to(dev) for i in range(5)]
    pool = api.ShotStreams(SP, n=2, device=dev, interlock=interlock)
    plan = cuda_rt.plan_from_params(SP)
    T = int(cuda_rt.lib.plan_num_segments(plan, n))
    nt = T // 128
    Sp = [pool.empty_image(C, 256, T) for _ in xs]
    Dp = [pool.empty_image(C, 256, T) for _ in xs]
    tp = [torch.empty((C * nt, 256, 128), device=dev) for _ in xs]
    infos = [torch.full((C, 4), -7, dtype=torch.int32, device=dev) for _ in xs]
    before = pool.launch_count()
    for i, x in enumerate(xs):
        assert pool.submit(x, Sp[i], Dp[i], clip=True, tiles=tp[i], tile_w=128, ntiles=nt, info=infos[i]) == i % 2
    pool.join()
    torch.cuda.current_stream().synchronize()
    assert pool.launch_count() > before
    for i, x in enumerate(xs):
        S = cuda_rt.empty_image(C, 256, T)
        D = cuda_rt.empty_image(C, 256, T)
        tl = torch.empty_like(tp[i])
        cuda_rt.pipeline_dev(plan, x, S, D, clip=True, tiles=tl, tile_w=128, ntiles=nt)
        torch.cuda.synchronize()
        assert torch.equal(Sp[i][..., :T], S[..., :T]) and torch.equal(Dp[i][..., :T], D[..., :T]) and torch.equal(tp[i], tl)
        assert infos[i][:, 3].max().item() == 0 and infos[i][:, 0].min().item() == 1
        assert torch.equal(infos[i][:, :2].cpu(), torch.tensor([[1, 256]] * C, dtype=torch.int32))
    Sr, _, _ = oc.specgr_array(xs[4][0].cpu().numpy().astype(np.float64), SP)
    np.testing.assert_allclose(Sp[4][0][:, :T].cpu().numpy(), Sr, rtol=0, atol=pc.ATOL_IMAGE)


def test_torch_cuda_zero_copy(cuda_rt):
    import torch
    x = torch.from_numpy(pc.signals(2, 100_000)).cuda()
    S, f, t = api.spectrogram_batch(x, SP, runtime=cuda_rt)
    assert isinstance(S, torch.Tensor) and S.is_cuda and S.shape == (2, 256, 389)
    Sr, _, _ = oc.specgr_array(x.cpu().numpy().astype(np.float64), SP)
    np.testing.assert_allclose(S.cpu().numpy(), Sr, rtol=0, atol=pc.ATOL_IMAGE)
