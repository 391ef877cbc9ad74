"""Host logic + kernel index arithmetic on the CPU: the product kernel sources compiled with
-DSPECGPU_EMULATE (tests/emu/cuda_emu.h: CUDA threads -> OS threads) and driven through the same
ctypes/C-ABI/host-API stack as on the GPU, at sizes the emulation finishes in seconds.  The parity
tests proper are tests/test_gpu_parity.py (-m gpu); this suite exists so that indexing, tiling and
host bookkeeping bugs are caught without a device.  tcgen05 PTX is not emulated (gram_tc.cu carries a
numerically equivalent scalar stand-in under SPECGPU_EMULATE)."""
import os

import numpy as np
import pytest

import parity_cases as pc
from oracle import spec_oracle as oc
from spectrogram_enhancement_b200 import api


@pytest.mark.parametrize("nperseg", [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_spectrogram_every_size(emu_rt, nperseg):
    n = max(5 * nperseg, 3000) if nperseg < 4096 else 3 * nperseg
    pc.case_spectrogram(emu_rt, nperseg, nperseg // 2, n, "linear", "hamm", "density")


@pytest.mark.parametrize("nperseg,noverlap,n,detrend,window,scaling", [
    (512, 256, 20000, "constant", "hann", "spectrum"),
    (512, 63, 5001, False, "boxcar", "density"),
    (64, 63, 700, "linear", "hann", "density"),
    (256, 0, 2048, "constant", "hamming", "density"),
])
def test_spectrogram_variants(emu_rt, nperseg, noverlap, n, detrend, window, scaling):
    pc.case_spectrogram(emu_rt, nperseg, noverlap, n, detrend, window, scaling)


def test_spectrogram_property_random_parameters(emu_rt):
    from hypothesis import HealthCheck, given, settings, strategies as st

    @settings(max_examples=8, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(log2n=st.integers(3, 9), ov=st.floats(0.0, 0.9), nseg=st.integers(1, 12),
           window=st.sampled_from(["hann", "hamm", "boxcar"]), detrend=st.sampled_from([False, "constant", "linear"]),
           scaling=st.sampled_from(["density", "spectrum"]))
    def run(log2n, ov, nseg, window, detrend, scaling):
        nperseg = 1 << log2n
        noverlap = min(int(ov * nperseg), nperseg - 1)
        n = nperseg + (nseg - 1) * (nperseg - noverlap) + 3
        pc.case_spectrogram(emu_rt, nperseg, noverlap, n, detrend, window, scaling, B=1)

    run()


def test_spectrogram_custom_window(emu_rt):
    w = np.hanning(130)[1:-1]
    pc.case_spectrogram(emu_rt, 128, 64, 3000, "constant", w, "density")


def test_specgr_reference_defaults(emu_rt):
    pc.case_specgr(emu_rt, oc.DEFAULT_SPEC_PARAMS, 20000)


def test_specgr_golden_small(emu_rt, golden):
    g = golden("specgr_small.npz")
    S, f, t = api.spectrogram_batch(g["x"], oc.DEFAULT_SPEC_PARAMS, runtime=emu_rt)
    np.testing.assert_allclose(S, g["S_f64"], rtol=0, atol=pc.ATOL_IMAGE)
    assert np.array_equal(f, g["f_f64"]) and np.array_equal(t, g["t_f64"])


@pytest.mark.parametrize("boundary,padded", [("zeros", True), (None, True), (None, False), ("zeros", False)])
def test_stft(emu_rt, boundary, padded):
    pc.case_stft(emu_rt, 256, 128, 3001, boundary, padded)


def test_rescale_norm(emu_rt):
    pc.case_rescale_norm(emu_rt, (37, 101))


@pytest.mark.parametrize("rows,thr", [(256, 0.9), (257, 0.9), (100, 0.5), (33, 0.123), (64, 0.0), (64, 1.0), (513, 0.99)])
def test_quantfilt(emu_rt, rows, thr):
    pc.case_quantfilt(emu_rt, rows, 45, thr)


def test_quantfilt_ties_and_3d(emu_rt):
    pc.case_quantfilt(emu_rt, 128, 40, 0.9, ties=True)
    pc.case_quantfilt_3d(emu_rt, 64, 33, 3)


def test_clip(emu_rt):
    a = np.array([[-1.0, 0.0, 2.5], [np.nan, -0.0, -3.0]], np.float32)
    out = api.clip(a, runtime=emu_rt)
    assert np.array_equal(out, oc.clip(a), equal_nan=True) and np.signbit(out[1, 1])


def test_patch_roundtrip(emu_rt):
    pc.case_patch(emu_rt, 3, 16, 70, 8, 8)
    pc.case_patch(emu_rt, 1, 256, 130, 128, 1)


def test_svd_default_power(emu_rt):
    pc.case_svd_default(emu_rt, 64, 200, [50, 20, 10])
    pc.case_svd_default(emu_rt, 48, 90, [30, 5], clip=True)


def test_svd_default_tf32_gram(emu_rt):
    # rows = 128 takes the tensor-core Gram route (TF32 operand rounding, emulated bit-for-bit)
    pc.case_svd_default(emu_rt, 128, 300, [200, 20, 10])


def test_svd_jacobi_range_optimal_compute(emu_rt):
    pc.case_svd_range(emu_rt, 32, 80, [40, 20, 10, 5], 0, 3)
    pc.case_svd_range(emu_rt, 32, 80, [40, 20, 10, 5], 1, -28)
    pc.case_svd_optimal(emu_rt, 32, 90, [40, 20, 10])
    pc.case_compute_signal(emu_rt, 32, 90, [40, 20])


def test_svd_tall_matrix_is_transposed(emu_rt):
    m = oc.synth_lowrank(90, 32, [40, 20, 10], 0.05, 9)
    d = api.denoiseSignal(m, runtime=emu_rt)
    pc.assert_denoise_close(d, oc.denoiseSignal(m.astype(np.float64)))


def test_csd(emu_rt):
    pc.case_csd(emu_rt, 4, 6000, 256)
    pc.case_csd(emu_rt, 5, 3000, 64, detrend="linear", scaling="spectrum")


def test_csd_wide_stacks(emu_rt):
    # 8x4 tiles: folded interleaved-segment warps (9 and 16 channels), the shared-memory staged kernel (27 and 40
    # channels, ragged last tiles, odd segment counts) and a non-symmetric row block accumulated over segment blocks
    pc.case_csd(emu_rt, 9, 3000, 64)
    pc.case_csd(emu_rt, 16, 2100, 32)
    pc.case_csd(emu_rt, 27, 2100, 32)
    pc.case_csd(emu_rt, 40, 2000, 64)
    pc.case_csd(emu_rt, 40, 200, 64)          # fewer stages than the ring is deep
    pc.case_csd(emu_rt, 40, 64, 64)           # a single segment
    pc.case_csd(emu_rt, 33, 100, 8)           # five bins: most lanes of the frequency block idle
    pc.case_csd_row_block(emu_rt, 40, 2100, 32, 8, 24, nblocks=3)
    pc.case_csd_row_block(emu_rt, 40, 1500, 32, 20, 20, nblocks=1)
    pc.case_csd_row_block(emu_rt, 6, 1500, 32, 2, 3, nblocks=2)


def test_csd_scipy_kat(emu_rt):
    # scipy/signal/tests/test_spectral.py TestCSD.test_real_onesided_even
    x = np.zeros(16, np.float32)
    x[0] = 1
    x[8] = 1
    f, p = api.csd(x, x, nperseg=8, runtime=emu_rt)
    np.testing.assert_allclose(f, np.linspace(0, 0.5, 5))
    np.testing.assert_allclose(p.real, [0.08333333, 0.15277778, 0.22222222, 0.22222222, 0.11111111], rtol=1e-5)
    np.testing.assert_allclose(p.imag, 0, atol=1e-7)


def test_ae_co2_time_resolved(emu_rt):
    pc.case_ae_co2(emu_rt, 6000, 128, 4)


def test_pipeline_small(emu_rt):
    sp = dict(oc.DEFAULT_SPEC_PARAMS, nperseg=32, noverlap=16)
    pc.case_pipeline(emu_rt, sp, 9000, B=2, tile=64)


def test_pipeline_fused_normalise_route(emu_rt):
    # 128 frequency rows: the tensor-core Gram + rank-1 projection route with the min-max normalisation fused
    sp = dict(oc.DEFAULT_SPEC_PARAMS, nperseg=256, noverlap=128)
    pc.case_pipeline(emu_rt, sp, 20000, B=3, tile=64)
    # tile widths that are not a multiple of the projection's 32-column CTA tile: the lanes of one warp straddle two VAE
    # tiles, and the columns past the last tile are not exported (the projection writes the tiles itself)
    pc.case_pipeline(emu_rt, sp, 20000, B=2, tile=50)


def test_pipeline_fallback_route(emu_rt):
    # power iteration capped at one step: every channel goes through the float64 repair route inside specgpu_pipeline
    pc.case_pipeline_fallback(emu_rt, dict(oc.DEFAULT_SPEC_PARAMS, nperseg=64, noverlap=32), 6000, B=2)


def test_svd_optimal_repeated_value(emu_rt):
    # a repeated singular value inside the kept range (values-first solver: inverse iteration + Gram-Schmidt, or Jacobi)
    pc.case_svd_optimal(emu_rt, 48, 200, [30.0, 30.0, 12.0, 5.0], noise=0.02)


def test_svd_optimal_values_first_route_alone(emu_rt, monkeypatch):
    # SPECGPU_TRIDIAG_STRICT drops the Jacobi fallback: the cluster tridiagonalisation (3 CTAs under emulation), the
    # division-free bisection, inverse iteration and the back-transformation have to carry the case by themselves
    monkeypatch.setenv("SPECGPU_TRIDIAG_STRICT", "1")
    pc.case_svd_optimal(emu_rt, 40, 150, [30.0, 12.0, 5.0], noise=0.02)
    pc.case_svd_optimal(emu_rt, 33, 90, [20.0, 20.0, 6.0], noise=0.02, seed=4)


def test_svd_degenerate_leading_pair_and_null_start(emu_rt):
    pc.case_svd_degenerate(emu_rt, rows=32, cols=120)


def test_cv2_chain_small(emu_rt, golden):
    # a crop of the reference's spectrogram against the oracle (which test_oracle.py pins to cv2 and to the reference's
    # golden outputs); the full golden image runs in tests/test_gpu_parity.py -- the emulation is too slow for it
    S = np.ascontiguousarray(golden("specgr_small.npz")["S_f32"][:32, :70])
    pc.case_filter_chain(emu_rt, S)


def test_cv2_chain_batched_small(emu_rt):
    S = np.random.default_rng(5).random((2, 24, 40)).astype(np.float32)
    out = api.gaussblr(S, (9, 3), runtime=emu_rt)
    for i in range(2):
        assert np.array_equal(out[i], oc.gaussblr(S[i], (9, 3)))
    pc.case_meansub_wide(emu_rt)
    pc.case_cv2_many_rows(emu_rt, (3, 21, 33))
    pc.case_cv2_pitched(emu_rt)
    pc.case_cv2_division_corners(emu_rt)


def test_cv2_chain_tiles(emu_rt):
    # several blur tiles (16 x 512) and morphology tiles (32 x 256) with ragged edges, plus degenerate images
    # (every emulated launch costs a fixed ~0.2 s, so the case list is short; tests/test_gpu_parity.py runs the long one)
    pc.case_cv2_tiles(emu_rt, [(30, 560)], [(31, 3)])
    pc.case_cv2_tiles(emu_rt, [(17, 1030)], [(5, 7)])
    pc.case_cv2_tiles(emu_rt, [(3, 9), (12, 1), (2, 2)], [(1, 1), (9, 3)])


def test_specgr_from_pickle_and_load_shot(emu_rt, tmp_path):
    # the reference's file entry point (pipeline_data.py:28-36) and the load-once reader the batched calls use
    import pickle
    n = 3000
    data = {"\\tecef%.2i" % c: oc.synth_ece(0, c, n=n + 50).astype(np.float64) for c in (1, 2, 8)}
    fname = tmp_path / "shot.pkl"
    with open(fname, "wb") as fh:
        pickle.dump(data, fh)
    sp = dict(oc.DEFAULT_SPEC_PARAMS, nperseg=64, noverlap=32, fs=1500.0)      # cut_shot * fs = 3000 samples
    S, f, t = api.specgr(str(fname), 8, sp, cut_shot=2, runtime=emu_rt)
    Sr, fr, tr = oc.specgr_array(data["\\tecef08"][:n].astype(np.float32), sp)
    assert S.shape == Sr.shape and np.array_equal(f, fr) and np.array_equal(t, tr)
    np.testing.assert_allclose(S, Sr, rtol=0, atol=pc.ATOL_IMAGE)      # the log / min-max image: its own tolerance
    x = api.load_shot(str(fname), channels=(1, 2, 8), cut_shot=2, fs=1500.0)
    assert x.shape == (3, n) and x.dtype == np.float32
    assert np.array_equal(x[2], data["\\tecef08"][:n].astype(np.float32))
    Sb, _, _ = api.spectrogram_batch(x, sp, runtime=emu_rt)
    assert np.array_equal(Sb[2], S)                              # batched == per channel
    # the whole loop body for the shot: spec + f + t + pipeline_out, from the path or from the array
    res = api.process_shot(str(fname), sp, channels=(1, 2, 8), cut_shot=2, runtime=emu_rt)
    assert sorted(res) == ["f", "pipeline_out", "spec", "t"] and np.array_equal(res["spec"][2], S)
    assert res["pipeline_out"].dtype == np.float64 and res["pipeline_out"].shape == res["spec"].shape
    pc.assert_same_f64(res["pipeline_out"][2], oc.filter_chain(S))
    res2 = api.process_shot(x, sp, runtime=emu_rt)
    assert np.array_equal(res2["pipeline_out"], res["pipeline_out"])
    # a rewritten file is reloaded, not served from the cache
    data["\\tecef08"] = data["\\tecef08"] * 2.0 + 1.0
    with open(fname, "wb") as fh:
        pickle.dump(data, fh)
    os.utime(fname, ns=(1, 1))
    x2 = api.load_shot(str(fname), channels=(8,), cut_shot=2, fs=1500.0)
    assert np.array_equal(x2[0], data["\\tecef08"][:n].astype(np.float32))


# ---- host logic / error behaviour ---------------------------------------------------------------
def test_errors_mirror_scipy(emu_rt):
    x = np.zeros(100, np.float32)
    with pytest.raises(ValueError, match="noverlap must be less than nperseg"):
        api.spectrogram(x, nperseg=8, noverlap=8, runtime=emu_rt)
    with pytest.raises(ValueError, match="Unknown scaling"):
        api.spectrogram(x, nperseg=8, noverlap=4, scaling="foo", runtime=emu_rt)
    with pytest.raises(ValueError, match="power of two"):
        api.spectrogram(x, nperseg=100, noverlap=4, runtime=emu_rt)
    with pytest.raises(ValueError, match="Quantiles must be in the range"):
        api.quantfilt(np.zeros((4, 4), np.float32), 1.5, runtime=emu_rt)
    with pytest.raises(ValueError):
        api.stft(x, nperseg=8, boundary="even", runtime=emu_rt)
    with pytest.raises(ValueError, match="cols=120001"):
        api.meansub(np.zeros((1, 120001)), runtime=emu_rt)
    with pytest.raises(ValueError, match="Quantiles must be in the range"):
        api.filter_chain(np.zeros((4, 8), np.float32), thr=-0.1, runtime=emu_rt)


def test_empty_and_short_inputs(emu_rt):
    f, t, P = api.spectrogram(np.zeros((2, 100), np.float32), nperseg=256, noverlap=128, runtime=emu_rt)
    assert P.shape == (2, 129, 0) and t.shape == (0,)
    f, t, P = api.spectrogram(np.zeros((0, 1000), np.float32), nperseg=256, noverlap=128, runtime=emu_rt)
    assert P.shape == (0, 129, 6)
    out = api.patch([], runtime=emu_rt)
    assert out.shape[0] == 0


def test_torch_in_torch_out(emu_rt):
    import torch
    x = torch.from_numpy(pc.signals(1, 3000))
    f, t, P = api.spectrogram(x, fs=500000, nperseg=256, noverlap=128, runtime=emu_rt)
    assert isinstance(P, torch.Tensor) and P.shape == (1, 129, 22)


def test_launch_counter(emu_rt):
    before = emu_rt.launch_count()
    api.spectrogram(np.zeros(1000, np.float32), nperseg=64, noverlap=32, runtime=emu_rt)
    assert emu_rt.launch_count() == before + 1


# ---- shot files: BES keys and the HDF5 interchange layout ------------------------------------------------------------
class _FakeH5Group(dict):
    """The few h5py calls save_shot_hdf5 / load_hdf5_dataset make, on nested dicts (h5py is not in this image)."""

    def _walk(self, name, create=False):
        node = self
        for part in name.split("/"):
            if part not in dict.keys(node):
                if not create:
                    raise KeyError(name)
                dict.__setitem__(node, part, _FakeH5Group())
            node = dict.__getitem__(node, part)
        return node

    def __contains__(self, name):
        try:
            self._walk(name)
            return True
        except KeyError:
            return False

    def __getitem__(self, name):
        return self._walk(name)

    def __delitem__(self, name):
        head, _, tail = name.rpartition("/")
        dict.__delitem__(self._walk(head) if head else self, tail)

    def create_group(self, name):
        if name in self:
            raise ValueError("group exists")
        return self._walk(name, create=True)

    def create_dataset(self, name, data):
        dict.__setitem__(self, name, np.array(data))


def test_bes_keys_and_hdf5_layout(emu_rt, tmp_path):
    import pickle
    sp = dict(oc.DEFAULT_SPEC_PARAMS, nperseg=64, noverlap=32)
    n = 3000
    sig = {c: oc.synth_ece(9, c, n=n) for c in (1, 2)}
    ece, bes = tmp_path / "ece_123456.pkl", tmp_path / "bes_123456.pkl"
    pickle.dump({"\\tecef%.2i" % c: sig[c] for c in sig}, open(ece, "wb"))
    pickle.dump({"besfu{:02d}".format(c): {"data.BES": sig[c]} for c in sig}, open(bes, "wb"))
    cut = n / sp["fs"]
    S_e, f_e, t_e = api.specgr(str(ece), 2, sp, cut, runtime=emu_rt)
    S_b, f_b, t_b = api.specgr(str(bes), 2, sp, cut, runtime=emu_rt, key="bes")        # denoising_by_svd.ipynb:49-63
    assert np.array_equal(S_e, S_b) and np.array_equal(t_e, t_b)
    Sr, fr, tr = oc.specgr_array(sig[2].astype(np.float64), sp)
    np.testing.assert_allclose(S_b, Sr, rtol=0, atol=pc.ATOL_IMAGE)
    assert np.array_equal(api.load_shot(str(bes), channels=(1, 2), cut_shot=cut, fs=sp["fs"], key="bes"), np.stack([sig[1], sig[2]]))
    with pytest.raises(ValueError):
        api.specgr(str(ece), 1, sp, cut, runtime=emu_rt, key="co2")
    # the interchange layout of pipeline_data.py:112-116, read back the way VAE/manual_scan.py:137-148 does
    res = api.process_shot(str(ece), sp, channels=(1, 2), cut_shot=cut, runtime=emu_rt)
    root = _FakeH5Group()
    api.save_shot_hdf5(root, "123456", res, channels=(1, 2))
    api.save_shot_hdf5(root, "123456", res, channels=(1, 2))          # re-runnable (the reference's create_group is not)
    assert sorted(dict.keys(root)) == ["ece_123456"] and sorted(dict.keys(root["ece_123456"])) == ["chn_1", "chn_2"]
    g = root["ece_123456/chn_2"]
    assert sorted(dict.keys(g)) == ["f", "pipeline_out", "spec", "t"]
    assert np.array_equal(g["spec"], res["spec"][1]) and g["pipeline_out"].dtype == np.float64
    assert np.array_equal(g["f"], res["f"]) and np.array_equal(g["t"], res["t"])
    specs, final = api.load_hdf5_dataset(root, n_channels=20)
    assert len(specs) == 2 and np.array_equal(final[1], res["pipeline_out"][1])
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            api.save_shot_hdf5(str(tmp_path / "x.hdf5"), "1", res, channels=(1, 2))
    else:
        path = str(tmp_path / "x.hdf5")
        api.save_shot_hdf5(path, "123456", res, channels=(1, 2))
        specs2, final2 = api.load_hdf5_dataset(path)
        assert np.array_equal(specs2[0], res["spec"][0]) and np.array_equal(final2[1], res["pipeline_out"][1])
