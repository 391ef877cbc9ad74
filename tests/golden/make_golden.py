"""Generate the golden fixtures in this directory by running the UNMODIFIED reference functions
(imported from /root/reference through oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container only (the reference tree does not travel):

    python tests/golden/make_golden.py

Outputs (all small, committed):
  specgr_small.npz    specgr() of pipeline_data.py on 20 000 samples, f32 and f64 input, + the
                      cv2 chain quantfilt/gaussblr/meansub/morph/meansub on it
  specgr_full_cols.npz  specgr() on the full 1 000 000-sample synthetic channel (shot 0, ch 0),
                      sampled columns only + min/max + per-row sums
  svd_small.npz       omega / denoiseSignal (default, explicit ranges, use_optimal) /
                      computeSignal of denoising_by_svd.ipynb on a planted-gap matrix and on the
                      small spectrogram
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader, spec_oracle as oc  # noqa: E402


def _pickle_signal(sig, ecen):
    f = tempfile.NamedTemporaryFile(suffix="_123456.pkl", delete=False)
    pickle.dump({"\\tecef%.2i" % ecen: sig}, f)
    f.close()
    return f.name


def main():
    assert ref_loader.available(), "reference tree not found"
    ref = ref_loader.load_pipeline_data()
    nbk = ref_loader.load_svd_notebook()
    sp = dict(oc.DEFAULT_SPEC_PARAMS)

    # ---- specgr, small ------------------------------------------------------------------
    n_small = 20000
    x32 = oc.synth_ece(0, 3, n=n_small)
    out = {"x": x32, "cut_shot": np.float64(n_small / sp["fs"])}
    for tag, sig in (("f32", x32), ("f64", x32.astype(np.float64))):
        fname = _pickle_signal(sig, 4)
        S, f, t = ref.specgr(fname, 4, sp, n_small / sp["fs"])
        os.unlink(fname)
        out[f"S_{tag}"], out[f"f_{tag}"], out[f"t_{tag}"] = S, f, t
    S = out["S_f32"]
    q = ref.quantfilt(S, 0.9)
    g = ref.gaussblr(q, (31, 3))
    m = ref.meansub(g)
    mo = ref.morph(m)
    fin = ref.meansub(mo)
    out.update(quant_f32=q, quant_thr_f32=np.quantile(S, 0.9, axis=0), gauss=g, mean=m, morph=mo, final=fin)
    out["quant_f64"] = ref.quantfilt(out["S_f64"], 0.9)
    out["norm_f32"] = ref.norm(S)
    out["rescale_f32"] = ref.rescale(S * 3 - 1)
    np.savez_compressed(os.path.join(HERE, "specgr_small.npz"), **out)

    # ---- specgr, full size, sampled ------------------------------------------------------
    xfull = oc.synth_ece(0, 0)
    fname = _pickle_signal(xfull, 1)
    S, f, t = ref.specgr(fname, 1, sp, 2)
    os.unlink(fname)
    fname = _pickle_signal(xfull.astype(np.float64), 1)
    S64, _, _ = ref.specgr(fname, 1, sp, 2)
    os.unlink(fname)
    cols = np.unique(np.concatenate([np.arange(0, 3905, 61), [1, 2, 3903, 3904]]))
    np.savez_compressed(
        os.path.join(HERE, "specgr_full_cols.npz"),
        x_crc=np.uint32(zlib.crc32(xfull.tobytes())), cols=cols, S_f32=S[:, cols], S_f64=S64[:, cols],
        rowsum_f64=S64.sum(axis=1), shape=np.array(S.shape), f=f, t=t,
        dtype_f32=str(S.dtype))

    # ---- SVD denoise ---------------------------------------------------------------------
    svd = {}
    M = oc.synth_lowrank(64, 300, [300.0, 120.0, 60.0, 30.0], 0.05, seed=7)
    svd["M"] = M
    svd["omega_beta"] = np.array([64 / 300, 256 / 3905, 1.0, 0.5])
    svd["omega"] = np.array([nbk["omega"](b) for b in svd["omega_beta"]])
    for tag, mat in (("M32", M), ("M64", M.astype(np.float64)), ("S32", out["S_f32"]),
                     ("S64", out["S_f64"])):
        svd[f"{tag}_s"] = np.linalg.svd(mat, compute_uv=False)
        svd[f"{tag}_default"] = nbk["denoiseSignal"](mat)
        svd[f"{tag}_optimal"] = nbk["denoiseSignal"](mat, use_optimal=True)
        if tag == "S64":
            continue
        svd[f"{tag}_0_4"] = nbk["denoiseSignal"](mat, 0, 4)
        svd[f"{tag}_2_9"] = nbk["denoiseSignal"](mat, 2, 9)
        svd[f"{tag}_m3_1000"] = nbk["denoiseSignal"](mat, -3, 1000)
        svd[f"{tag}_compute"] = nbk["computeSignal"](mat)
    np.savez_compressed(os.path.join(HERE, "svd_small.npz"), **svd)

    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
