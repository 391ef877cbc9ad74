"""Parity cases shared by the CPU-emulation suite (tests/test_emulated.py, tiny sizes) and the GPU
suite (tests/test_gpu_parity.py, -m gpu, through the real libspecgpu.so).  Each case drives the
product host API (spectrogram_enhancement_b200.api) on a Runtime and checks it against the CPU
oracle (oracle/spec_oracle.py) run in float64 on the same float32 input.

Tolerances (BASELINE.json north_star; SURVEY.md section 7.2 "precision contract"):
  * linear spectra (PSD, STFT, CSD):    rtol 1e-4, atol 1e-6 * max|oracle|
  * log/min-max normalised image:       atol 1e-4 on the [0, 1] image
  * denoised reconstruction:            rtol 1e-3, atol 1e-3 * max|oracle|
  * integer outputs (segment counts, axes, masks, start/stop/num_sing, tile copies): bit-exact
"""
import numpy as np

from oracle import spec_oracle as oc
from spectrogram_enhancement_b200 import api

RTOL_SPEC = 1e-4
ATOL_SPEC_REL = 1e-6
ATOL_IMAGE = 1e-4
RTOL_DENOISE = 1e-3
ATOL_DENOISE_REL = 1e-3


def assert_same_f64(got, ref):
    """The cv2 chain's float64 outputs are reproduced bit for bit (numpy's pairwise row sums included); NaN == NaN for
    constant images, whose rescale is 0 / 0 on both sides."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.dtype == np.float64 and got.shape == ref.shape
    with np.errstate(invalid="ignore"):
        same = np.array_equal(got, ref, equal_nan=True)
    assert same, "max |diff| = %g on %d elements" % (np.nanmax(np.abs(got - ref)), int((got != ref).sum()))


def assert_spec_close(got, ref):
    ref = np.asarray(ref)
    np.testing.assert_allclose(got, ref, rtol=RTOL_SPEC, atol=ATOL_SPEC_REL * float(np.abs(ref).max()))


def assert_denoise_close(got, ref):
    ref = np.asarray(ref)
    np.testing.assert_allclose(got, ref, rtol=RTOL_DENOISE, atol=ATOL_DENOISE_REL * float(np.abs(ref).max()))


def signals(B, n, shot=1, fs=500000.0, offset=0.0):
    return np.stack([oc.synth_ece(shot, c, n=n, fs=fs) for c in range(B)]) + np.float32(offset)


# ---- spectrogram front-end ---------------------------------------------------------------------
def case_spectrogram(rt, nperseg, noverlap, n, detrend, window, scaling, B=2):
    x = signals(B, n, offset=0.7)
    f, t, P = api.spectrogram(x, fs=500000, window=window, nperseg=nperseg, noverlap=noverlap, detrend=detrend,
                              scaling=scaling, runtime=rt)
    fr, tr, Pr = oc.spectrogram(x.astype(np.float64), fs=500000, window=window, nperseg=nperseg, noverlap=noverlap,
                                detrend=detrend, scaling=scaling)
    assert P.shape == Pr.shape and P.dtype == np.float32
    assert np.array_equal(f, fr) and np.array_equal(t, tr)          # index-derived: exact
    assert_spec_close(P, Pr)


def case_specgr(rt, sp, n, B=2):
    x = signals(B, n)
    S, f, t = api.spectrogram_batch(x, sp, runtime=rt)
    Sr, fr, tr = oc.specgr_array(x.astype(np.float64), sp)
    assert S.shape == Sr.shape
    assert np.array_equal(f, fr) and np.array_equal(t, tr)
    np.testing.assert_allclose(S, Sr, rtol=0, atol=ATOL_IMAGE)
    assert S.min() >= 0.0 and S.max() <= 1.0
    # the exported (min, max) of the log image are natural logs (the kernel keeps the image in base 2 internally)
    import torch
    xd, _ = rt.to_device(x)
    mm = torch.empty((B, 2), dtype=torch.float32, device=rt.device)
    rt.specgr_dev(rt.plan_from_params(sp), xd, minmax=mm)
    _, _, P = oc.spectrogram(x.astype(np.float64), fs=sp["fs"], window=sp["window"], nperseg=sp["nperseg"],
                             noverlap=sp["noverlap"], detrend=sp["detrend"], scaling=sp["scaling"])
    Lr = np.log(P + sp["eps"])
    np.testing.assert_allclose(mm.cpu().numpy(), np.stack([Lr.min(axis=(1, 2)), Lr.max(axis=(1, 2))], axis=1), rtol=0, atol=2e-4)
    return S, Sr


def case_stft(rt, nperseg, noverlap, n, boundary, padded, window="hann", B=1):
    x = signals(B, n)
    f, t, Z = api.stft(x, fs=500000, window=window, nperseg=nperseg, noverlap=noverlap, boundary=boundary, padded=padded,
                       runtime=rt)
    fr, tr, Zr = oc.stft(x.astype(np.float64), fs=500000, window=window, nperseg=nperseg, noverlap=noverlap,
                         boundary=boundary, padded=padded)
    assert Z.shape == Zr.shape and Z.dtype == np.complex64
    assert np.array_equal(f, fr)
    np.testing.assert_allclose(t, tr, rtol=0, atol=1e-12)
    assert_spec_close(Z, Zr)


# ---- helpers -------------------------------------------------------------------------------------
def case_rescale_norm(rt, shape, seed=0):
    a = (np.random.default_rng(seed).standard_normal(shape) * 3 + 1).astype(np.float32)
    np.testing.assert_allclose(api.rescale(a, runtime=rt), oc.rescale(a.astype(np.float64)), rtol=0, atol=2e-7)
    np.testing.assert_allclose(api.norm(a, runtime=rt), oc.norm(a.astype(np.float64)), rtol=1e-5, atol=1e-5)


def case_quantfilt(rt, rows, cols, thr, seed=0, ties=False):
    a = np.random.default_rng(seed).random((rows, cols)).astype(np.float32)
    if ties:
        a = np.round(a * 8).astype(np.float32) / 8      # heavy ties: many equal order statistics
    out, q, mask = api.quantfilt_mask(a, thr, runtime=rt)
    qr = np.quantile(a, thr, axis=0)
    assert q.dtype == np.float32 and np.array_equal(q, qr)                    # bit-exact threshold
    assert np.array_equal(out, oc.quantfilt(a, thr))                          # bit-exact image
    assert np.array_equal(mask.astype(bool), ~(a < qr))                       # bit-exact mask
    assert np.array_equal(api.quantfilt(a, thr, runtime=rt), out)


def case_quantfilt_3d(rt, rows, cols, chans, seed=0):
    a = np.random.default_rng(seed).random((rows, cols, chans)).astype(np.float32)
    assert np.array_equal(api.quantfilt(a, 0.9, runtime=rt), oc.quantfilt(a, 0.9))


def case_patch(rt, n, rows, cols, tile, ntiles, seed=0):
    arr = [np.random.default_rng(seed + i).random((rows, cols)).astype(np.float32) for i in range(n)]
    p = api.patch(arr, tile=tile, ntiles=ntiles, runtime=rt)
    pr = oc.patch(arr, tile=tile, ntiles=ntiles)
    assert p.dtype == np.float64 and np.array_equal(p, pr)
    p32 = api.patch(np.stack(arr), tile=tile, ntiles=ntiles, dtype=np.float32, runtime=rt)
    assert p32.dtype == np.float32 and np.array_equal(p32, pr.astype(np.float32))
    u = api.unpatch(p, ntiles=ntiles, runtime=rt)
    assert np.array_equal(u, oc.unpatch(pr, ntiles=ntiles))
    assert np.array_equal(u, np.stack(arr)[:, :, :tile * ntiles])             # round trip


# ---- SVD denoise -----------------------------------------------------------------------------------
def case_svd_default(rt, rows, cols, sv, noise=0.05, seed=3, method="auto", clip=False):
    m = oc.synth_lowrank(rows, cols, sv, noise, seed)
    d = api.denoiseSignal(m, clip=clip, method=method, runtime=rt)
    dr = oc.denoiseSignal(m.astype(np.float64))
    if clip:
        dr = oc.clip(dr)
    assert d.shape == dr.shape and d.dtype == np.float32
    assert_denoise_close(d, dr)


def case_svd_range(rt, rows, cols, sv, start, stop, noise=0.05, seed=4):
    m = oc.synth_lowrank(rows, cols, sv, noise, seed)
    d, s, info = api.denoiseSignal(m, start=start, stop=stop, return_info=True, runtime=rt)
    m64 = m.astype(np.float64)
    dr = oc.denoiseSignal(m64, start=start, stop=stop)
    sr = np.linalg.svd(m64, compute_uv=False)
    a, b, _ = oc.svd_plan(m.shape, sr, start, stop)
    ea, eb, _ = slice(a, b).indices(len(sr))            # what the python slice u[:, a:b] resolves to
    assert (int(info[0]), int(info[1])) == (ea, eb) and int(info[2]) == -1
    np.testing.assert_allclose(s, sr, rtol=1e-4, atol=1e-5 * sr[0])
    assert_denoise_close(d, dr)


def case_svd_optimal(rt, rows, cols, sv, noise=0.05, seed=5):
    m = oc.synth_lowrank(rows, cols, sv, noise, seed)
    d, s, info = api.denoiseSignal(m, use_optimal=True, return_info=True, runtime=rt)
    m64 = m.astype(np.float64)
    dr = oc.denoiseSignal(m64, use_optimal=True)
    sr = np.linalg.svd(m64, compute_uv=False)
    a, b, num_sing = oc.svd_plan(m.shape, sr, use_optimal=True)
    assert int(info[2]) == num_sing                                           # integer: exact
    ea, eb, _ = slice(a, b).indices(len(sr))
    assert (int(info[0]), int(info[1])) == (ea, eb)
    np.testing.assert_allclose(s, sr, rtol=1e-4, atol=1e-5 * sr[0])
    assert_denoise_close(d, dr)


def case_compute_signal(rt, rows, cols, sv, noise=0.05, seed=6):
    m = oc.synth_lowrank(rows, cols, sv, noise, seed)
    out, s, info = api.computeSignal(m, return_info=True, runtime=rt)
    ref = oc.computeSignal(m.astype(np.float64))
    assert out.dtype == np.float64
    assert_denoise_close(out, ref)


# ---- cross-power spectrum ----------------------------------------------------------------------------
def case_csd(rt, C, n, nperseg, fs=1.6e6, detrend="constant", window="hann", scaling="density"):
    x = signals(C, n, shot=2, fs=fs, offset=0.3)
    f, P = api.csd_allpairs(x, fs=fs, window=window, nperseg=nperseg, noverlap=nperseg // 2, detrend=detrend,
                            scaling=scaling, runtime=rt)
    fr, Pr = oc.csd_allpairs(x.astype(np.float64), fs=fs, window=window, nperseg=nperseg, noverlap=nperseg // 2,
                             detrend=detrend, scaling=scaling)
    assert np.array_equal(f, fr) and P.shape == Pr.shape
    assert_spec_close(P, Pr)
    # Hermitian in the pair index, real diagonal
    np.testing.assert_allclose(P, np.conj(np.swapaxes(P, 0, 1)), rtol=1e-5, atol=1e-7 * np.abs(Pr).max())
    f2, P01 = api.csd(x[0], x[1], fs=fs, window=window, nperseg=nperseg, noverlap=nperseg // 2, detrend=detrend,
                      scaling=scaling, runtime=rt)
    assert_spec_close(P01, Pr[0, 1])


def case_csd_row_block(rt, C, n, nperseg, i0, ni, nblocks=2, fs=1.6e6):
    """Rows [i0, i0 + ni) of the pair matrix through specgpu_csd_pairs_block, accumulated over `nblocks` segment blocks
    (the call pattern of parallel.csd_allpairs_sharded) against the oracle's full matrix."""
    import torch
    x = signals(C, n, shot=3, fs=fs, offset=0.1)
    plan = rt.plan(nperseg, nperseg // 2, fs, "hann", "density", "constant")
    xd, _ = rt.to_device(x)
    F = rt.lib.plan_num_freqs(plan)
    T = rt.lib.plan_num_segments(plan, n)
    hop = nperseg // 2
    ldf = (F + 1) & ~1
    P = rt.empty((ni, C, F, 2))
    edges = [T * b // nblocks for b in range(nblocks + 1)]
    for b in range(nblocks):
        t0, t1 = edges[b], edges[b + 1]
        X = rt.empty((C, t1 - t0, ldf, 2))
        xs = xd[:, t0 * hop:(t1 - 1) * hop + nperseg]
        rt.check(rt.lib.csd_spectra(rt._ctx, plan, xs.data_ptr(), C, xs.shape[1], api._ld(xd), X.data_ptr(), ldf, rt.stream()))
        rt.check(rt.lib.csd_pairs_block(rt._ctx, plan, X.data_ptr(), C, t1 - t0, T, ldf, i0, ni, 1 if b else 0, P.data_ptr(),
                                        rt.stream()))
    got = torch.view_as_complex(P).cpu().numpy()
    _, Pr = oc.csd_allpairs(x.astype(np.float64), fs=fs, nperseg=nperseg, noverlap=nperseg // 2)
    assert_spec_close(got, Pr[i0:i0 + ni])


def case_cv2_tiles(rt, shapes, ksizes, seed=11):
    """gaussblr / morph on noise images that span several shared-memory tiles (and on degenerate ones), uint8 results
    bit-exact against the oracle; meansub to float64 round-off."""
    rng = np.random.default_rng(seed)
    for shape in shapes:
        img = rng.random(shape)
        img[rng.random(shape) < 0.02] *= 4.0          # a few outliers so the quantised image keeps low values
        for ks in ksizes:
            g, g8 = api.gaussblr(img, ks, return_uint8=True, runtime=rt)
            q = (oc.rescale(img) * 255).astype("uint8")
            assert np.array_equal(g8, oc.gaussian_blur_u8(q, ks)), (shape, ks)
            with np.errstate(invalid="ignore"):          # a constant blurred image rescales to 0 / 0 = NaN on both sides
                assert np.array_equal(g, oc.gaussblr(img, ks), equal_nan=True), (shape, ks)
        mo, mo8 = api.morph(img, return_uint8=True, runtime=rt)
        assert np.array_equal(mo8, oc.morph_close_open_u8((oc.rescale(img) * 255).astype("uint8"))), shape
        with np.errstate(invalid="ignore"):
            assert np.array_equal(mo, oc.morph(img), equal_nan=True), shape
        assert_same_f64(api.meansub(img, runtime=rt), oc.meansub(img))
        f32 = img.astype(np.float32)
        with np.errstate(invalid="ignore"):
            assert np.array_equal(api.morph(f32, runtime=rt), oc.morph(f32), equal_nan=True), shape


def case_cv2_many_rows(rt, shape=(19, 63, 33)):
    """Stacks with more than 148 * 8 rows: the row kernels take several rows per CTA (ragged last group)."""
    S = np.random.default_rng(8).random(shape).astype(np.float32)
    g = api.gaussblr(S, (5, 3), runtime=rt)
    m = api.meansub(g, runtime=rt)
    fin = api.filter_chain(S, runtime=rt)
    for i in (0, shape[0] - 1):
        assert np.array_equal(g[i], oc.gaussblr(S[i], (5, 3)))
        assert_same_f64(m[i], oc.meansub(g[i]))
        assert_same_f64(fin[i], oc.filter_chain(S[i]))


def case_cv2_pitched(rt):
    """The C ABI takes row pitches: source rows padded to ld > cols and a padded float64 output."""
    import torch
    rng = np.random.default_rng(1)
    B, R, C, ld, ldo = 2, 20, 37, 45, 50
    buf = torch.from_numpy(rng.random((B, R, ld)).astype(np.float32)).to(rt.device)
    S = buf[:, :, :C].cpu().numpy().copy()
    out = torch.empty((B, R, C), dtype=torch.float64, device=rt.device)
    rt.check(rt.lib.filter_chain(rt._ctx, buf.data_ptr(), B, R, C, ld, 0.9, 31, 3, out.data_ptr(), C, rt.stream()))
    g = torch.empty((B, R, C), dtype=torch.float64, device=rt.device)
    rt.check(rt.lib.gaussblr(rt._ctx, buf.data_ptr(), 0, B, R, C, ld, 5, 3, g.data_ptr(), C, None, rt.stream()))
    mo = torch.zeros((B, R, ldo), dtype=torch.float64, device=rt.device)
    rt.check(rt.lib.morph(rt._ctx, buf.data_ptr(), 0, B, R, C, ld, mo.data_ptr(), ldo, None, rt.stream()))
    out, g, mo = out.cpu().numpy(), g.cpu().numpy(), mo.cpu().numpy()
    for i in range(B):
        assert_same_f64(out[i], oc.filter_chain(S[i]))
        assert np.array_equal(g[i], oc.gaussblr(S[i], (5, 3)))
        assert np.array_equal(mo[i, :, :C], oc.morph(S[i]))
    assert not mo[:, :, C:].any()                    # the padding of the output rows is left alone


def case_cv2_division_corners(rt):
    """Denominators the reciprocal-based quotient hands to the divider: an all-ones significand and a subnormal range."""
    rng = np.random.default_rng(12)
    base = rng.random((9, 40))
    base[0, 0], base[-1, -1] = 0.0, 1.0
    for dt, top in ((np.float32, np.float32(1.9999999)), (np.float64, np.nextafter(2.0, 0.0)),
                    (np.float32, np.float32(3e-41)), (np.float64, 5e-310)):
        img = (base * float(top)).astype(dt)
        img[-1, -1] = top
        assert float(img.max() - img.min()) == float(top)
        g, g8 = api.gaussblr(img, (5, 3), return_uint8=True, runtime=rt)
        assert np.array_equal(g8, oc.gaussian_blur_u8((oc.rescale(img) * 255).astype("uint8"), (5, 3))), (dt, top)
        mo, mo8 = api.morph(img, return_uint8=True, runtime=rt)
        assert np.array_equal(mo8, oc.morph_close_open_u8((oc.rescale(img) * 255).astype("uint8"))), (dt, top)
    m = rng.random((6, 50)) * float(np.nextafter(2.0, 0.0))
    assert_same_f64(api.meansub(m, runtime=rt), oc.meansub(m))


def case_meansub_wide(rt):
    """Rows wider than the 4096 columns the row-statistics kernel keeps in registers."""
    img = np.random.default_rng(3).random((3, 5000))
    assert_same_f64(api.meansub(img, runtime=rt), oc.meansub(img))


# ---- whole path ------------------------------------------------------------------------------------
def case_pipeline(rt, sp, n, B=2, tile=None):
    x = signals(B, n)
    T = oc.segment_count(n, sp["nperseg"], sp["noverlap"])
    if tile:
        S, D, tiles, info = api.pipeline(x, sp, clip=True, tiles=True, tile=tile, return_info=True, runtime=rt)
    else:
        S, D, info = api.pipeline(x, sp, clip=True, return_info=True, runtime=rt)
    Sr, _, _ = oc.specgr_array(x.astype(np.float64), sp)
    np.testing.assert_allclose(S, Sr, rtol=0, atol=ATOL_IMAGE)
    # stage-isolated: denoise OUR S with the oracle, so the SVD stage is judged on identical input
    Dr = np.stack([oc.clip(oc.denoiseSignal(s.astype(np.float64))) for s in S])
    assert_denoise_close(D, Dr)
    assert (D >= 0).all()
    rows = sp["nperseg"] // 2
    assert np.array_equal(info[:, :2], np.tile([1, rows], (B, 1)))
    if tile:
        nt = T // tile
        assert np.array_equal(tiles, oc.patch(list(D), tile=tile, ntiles=nt).astype(np.float32))
    # end to end against the pure oracle chain, at the north star's tolerance for the denoised reconstruction
    De = np.stack([oc.clip(oc.denoiseSignal(s)) for s in Sr])
    assert_denoise_close(D, De)


def case_pipeline_fallback(rt, sp, n, B=3):
    """The repair route of specgpu_pipeline: with the power iteration capped at one step no channel converges, every
    channel is flagged and redone by the float64 Gram + Jacobi solver in the stream; the result must equal the oracle
    like the regular route.  Without the fallback flag the status is reported in info[:, 3] instead."""
    x = signals(B, n, shot=7)
    plan = rt.plan_from_params(sp)
    xd, _ = rt.to_device(x)
    info = np.zeros((B, 4), np.int32)
    import torch
    info_d = torch.zeros((B, 4), dtype=torch.int32, device=rt.device)
    rt.set_power_iterations(1)
    try:
        S0, D0 = rt.pipeline_dev(plan, xd, clip=True, info=info_d, fallback=False)
        assert (info_d.cpu().numpy()[:, 3] == 1).all()          # reported, not repaired
        S1, D1 = rt.pipeline_dev(plan, xd, clip=True, info=info_d, fallback=True)
        assert (info_d.cpu().numpy()[:, 3] == 0).all()          # repaired in the stream
        S2, D2 = rt.pipeline_dev(plan, xd, clip=True, info=None)   # no info: the repair is always on
    finally:
        rt.set_power_iterations(0)
    Sr, _, _ = oc.specgr_array(x.astype(np.float64), sp)
    De = np.stack([oc.clip(oc.denoiseSignal(s)) for s in Sr])
    for S, D in ((S1, D1), (S2, D2)):
        np.testing.assert_allclose(S.contiguous().cpu().numpy(), Sr, rtol=0, atol=ATOL_IMAGE)
        assert_denoise_close(D.contiguous().cpu().numpy(), De)
    S3, D3 = rt.pipeline_dev(plan, xd, clip=True, info=info_d)     # the regular route, same data
    assert (info_d.cpu().numpy()[:, 3] == 0).all()
    assert_denoise_close(D3.contiguous().cpu().numpy(), De)


def case_svd_degenerate(rt, rows=64, cols=300, seed=21):
    """Default denoiseSignal on matrices the power iteration cannot handle by itself:
    (a) a leading pair 0.3 % apart (the iteration hits its cap; the float64 fallback must separate the pair);
    (b) a dominant component whose columns all sum to zero (G 1 = 0: a start vector of ones would fall into the null
        space and a zero iterate used to be reported as converged)."""
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((rows, rows)))
    V, _ = np.linalg.qr(rng.standard_normal((cols, rows)))
    sv = np.concatenate([[10.0, 9.97, 3.0, 1.0], 0.05 * rng.random(rows - 4)])
    A = ((U * sv) @ V.T).astype(np.float32)
    out, s, info = api.denoiseSignal(A, return_info=True, method="auto", runtime=rt)
    assert_denoise_close(out, oc.denoiseSignal(A.astype(np.float64)))
    assert np.array_equal(info[:2], [1, rows])
    a = np.concatenate([np.ones(rows // 2), -np.ones(rows - rows // 2)])
    if rows % 2:
        a[-1] = 0.0
    a /= np.linalg.norm(a)
    b = rng.standard_normal(rows)
    b -= a * (a @ b)
    b /= np.linalg.norm(b)
    v1, v2 = V[:, 0], V[:, 1]
    Z = (10.0 * np.outer(a, v1) + 1.0 * np.outer(b, v2)).astype(np.float32)
    ref = oc.denoiseSignal(Z.astype(np.float64))
    assert_denoise_close(api.denoiseSignal(Z, runtime=rt), ref)
    assert_denoise_close(api.denoiseSignal(np.stack([Z, A]), runtime=rt)[0], ref)      # batched, mixed statuses


# ---- cv2 image chain ---------------------------------------------------------------------------------
def case_filter_chain(rt, S):
    """Stage-isolated and chained parity of gaussblr / meansub / morph: uint8 intermediates bit-exact, float64
    outputs to 1e-12 (the row means are summed in another order than numpy's pairwise sum)."""
    q = oc.quantfilt(S, 0.9)
    g_ref = oc.gaussblr(q, (31, 3))
    g, g8 = api.gaussblr(q, (31, 3), return_uint8=True, runtime=rt)
    assert g.dtype == np.float64
    assert np.array_equal(g8, oc.gaussian_blur_u8((oc.rescale(q) * 255).astype("uint8"), (31, 3)))
    assert np.array_equal(g, g_ref)
    m_ref = oc.meansub(g_ref)
    m = api.meansub(g_ref, runtime=rt)
    assert_same_f64(m, m_ref)
    mo_ref = oc.morph(m_ref)
    mo, mo8 = api.morph(m_ref, return_uint8=True, runtime=rt)
    assert np.array_equal(mo8, oc.morph_close_open_u8((oc.rescale(m_ref) * 255).astype("uint8")))
    assert np.array_equal(mo, mo_ref)
    fin = api.filter_chain(S, runtime=rt)
    assert_same_f64(fin, oc.filter_chain(S))
    # the single fused call (uint8 planes between the stages) and the five chained public calls agree bit for bit
    assert np.array_equal(fin, api.filter_chain(S, runtime=rt, fused=False))
    return g, fin


# ---- time-resolved cross-power amplitude (ae_co2; PARITY-UNPINNED upstream: defined as frame-wise scipy.signal.csd) ----
def case_ae_co2(rt, n, nperseg, navg, fs=1.6e6):
    x = signals(2, n, shot=4, fs=fs, offset=0.2)
    t_ms = np.arange(n) / fs * 1e3 + 5.0
    amp, freq, time = api.ae_co2(x[0], x[1], t_ms, nperseg=nperseg, navg=navg, runtime=rt)
    hop = nperseg // 2
    frame = nperseg + (navg - 1) * hop
    nframes = n // frame
    assert amp.shape == (nframes, nperseg // 2 + 1) and amp.dtype == np.float32
    ref = np.stack([np.abs(oc.csd(x[0, k * frame:(k + 1) * frame].astype(np.float64),
                                  x[1, k * frame:(k + 1) * frame].astype(np.float64), fs=fs, nperseg=nperseg)[1])
                    for k in range(nframes)])
    assert_spec_close(amp, ref)
    np.testing.assert_allclose(freq, np.fft.rfftfreq(nperseg, 1 / fs) / 1e3, rtol=1e-9)      # kHz
    np.testing.assert_allclose(time, 5.0 + (np.arange(nframes) * frame + frame / 2) / fs * 1e3, rtol=1e-9)   # ms
