#!/usr/bin/env python
"""Device-resident timings of the other rows of the hot-path table: quantfilt, tile export,
STFT (config 1), SVD full-decomposition modes, all-pairs CSD (configs 3 and 5).  One JSON line per case; the
headline metric stays bench.py's.  CUDA events, 3 warm-ups, inputs rotated so they do not sit in L2."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spectrogram_enhancement_b200 import api  # noqa: E402

PEAK = 6543.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = float(json.load(open(p))["hbm_gbs"])
rt = api.Runtime()
dev = rt.device


def timeit(fn, iters=10, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(case, ms, algo_bytes, **kw):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(case=case, ms=round(ms, 4), algorithmic_MB=round(algo_bytes / 1e6, 1), GBps=round(gbs, 1),
                          frac_of_measured_hbm_peak=round(gbs / PEAK, 3), **kw)), flush=True)


g = torch.Generator(device=dev)
g.manual_seed(0)
NB = 3
# ---- quantfilt on 40 x [256, 3905] ----
imgs = [torch.rand((40, 256, 3905), device=dev, generator=g) for _ in range(NB)]
out = torch.empty_like(imgs[0])
thr = torch.empty((40, 3905), device=dev)


def qf(i):
    rt.check(rt.lib.quantfilt(rt._ctx, imgs[i % NB].data_ptr(), 40, 256, 3905, 3905, 0.9, out.data_ptr(), thr.data_ptr(), None,
                              rt.stream()))


report("quantfilt 40x[256x3905] thr=0.9", timeit(qf), 40 * 256 * 3905 * 8)
# ---- patch: 40 x [256, 3905] -> [1200, 256, 128] float32 / float64 ----
t32 = torch.empty((1200, 256, 128), device=dev)
t64 = torch.empty((1200, 256, 128), device=dev, dtype=torch.float64)
report("patch f32 tiles 40x30x[256x128]", timeit(lambda i: rt.check(rt.lib.patch(rt._ctx, imgs[i % NB].data_ptr(), 40, 256, 3905, 128, 30, t32.data_ptr(), 0, rt.stream()))),
       1200 * 256 * 128 * 8)
report("patch f64 tiles (reference dtype)", timeit(lambda i: rt.check(rt.lib.patch(rt._ctx, imgs[i % NB].data_ptr(), 40, 256, 3905, 128, 30, t64.data_ptr(), 1, rt.stream()))),
       1200 * 256 * 128 * 12)
# ---- cv2 chain on 40 x [256, 3905] (quantfilt output as input) ----
g64 = torch.empty((40, 256, 3905), device=dev, dtype=torch.float64)
m64 = torch.empty_like(g64)
o64 = torch.empty_like(g64)
report("gaussblr (31,3) 40x[256x3905] f32 -> f64", timeit(lambda i: rt.check(rt.lib.gaussblr(rt._ctx, imgs[i % NB].data_ptr(), 0, 40, 256, 3905, 3905, 31, 3, g64.data_ptr(), 3905, None, rt.stream())), iters=5),
       40 * 256 * 3905 * 12)
report("meansub 40x[256x3905] f64", timeit(lambda i: rt.check(rt.lib.meansub(rt._ctx, g64.data_ptr(), 40, 256, 3905, 3905, m64.data_ptr(), 3905, rt.stream())), iters=5),
       40 * 256 * 3905 * 16)
report("morph 40x[256x3905] f64", timeit(lambda i: rt.check(rt.lib.morph(rt._ctx, m64.data_ptr(), 1, 40, 256, 3905, 3905, o64.data_ptr(), 3905, None, rt.stream())), iters=5),
       40 * 256 * 3905 * 16)
report("filter_chain fused (quantfilt -> gaussblr -> meansub -> morph -> meansub) 40x[256x3905] f32 -> f64",
       timeit(lambda i: rt.check(rt.lib.filter_chain(rt._ctx, imgs[i % NB].data_ptr(), 40, 256, 3905, 3905, 0.9, 31, 3, o64.data_ptr(), 3905,
                                                     rt.stream())), iters=5), 40 * 256 * 3905 * 12)
# ---- specgr only (spectrogram + log + min-max), 40 channels ----
xs = [torch.randn((40, 1_000_000), device=dev, generator=g) for _ in range(NB)]
plan = rt.plan_from_params(api.DEFAULT_SPEC_PARAMS)
S = rt.empty((40, 256, 3905))
report("specgr 40ch x 1M (STFT+log+min-max), dense 3905-float rows", timeit(lambda i: rt.specgr_dev(plan, xs[i % NB], S)), 40 * (4e6 + 4 * 256 * 3905))
Sp = rt.empty_image(40, 256, 3905)
report("specgr 40ch x 1M (STFT+log+min-max), pitched rows (Runtime.empty_image, the API default)",
       timeit(lambda i: rt.specgr_dev(plan, xs[i % NB], Sp)), 40 * (4e6 + 4 * 256 * 3905))
# ---- config 1: STFT 1024/512 hann, complex output, 40 channels batched ----
p1 = rt.plan(1024, 512, 500000, "hann", "spectrum", False)
T1 = rt.lib.stft_num_segments(p1, 1_000_000, 1, 1)
Z = rt.empty((40, 513, T1, 2))
report("config1 stft 1024/512 hann, 40ch x 1M, complex64 out",
       timeit(lambda i: rt.check(rt.lib.stft(rt._ctx, p1, xs[i % NB].data_ptr(), 40, 1_000_000, 1_000_000, 1, 1, Z.data_ptr(), T1, rt.stream()))),
       40 * (4e6 + 8 * 513 * T1))
# ---- SVD modes on 40 x [256, 3905] ----
D = torch.empty_like(imgs[0])
info = torch.zeros((40, 4), dtype=torch.int32, device=dev)
s_out = rt.empty((40, 256))


def svd(mode_opt):
    def f(i):
        rt.check(rt.lib.svd_denoise(rt._ctx, imgs[i % NB].data_ptr(), 40, 256, 3905, 3905, 1, 256, mode_opt, 1, 0 if mode_opt == 0 else 1,
                                    D.data_ptr(), 3905, None if mode_opt == 0 else s_out.data_ptr(), info.data_ptr(), rt.stream()))
    return f


report("denoiseSignal default (power route) 40x[256x3905]", timeit(svd(0)), 40 * 256 * 3905 * 8)
report("denoiseSignal use_optimal (float64 Gram + values-first eigensolver) 40x[256x3905]", timeit(svd(1), iters=3, warm=1), 40 * 256 * 3905 * 8)
# ---- config 3: CSD 4 chords x 3.2 M, nperseg 4096 ----
x3 = [torch.randn((4, 3_200_000), device=dev, generator=g) for _ in range(NB)]
p3 = rt.plan(4096, 2048, 1.6e6, "hann", "density", "constant")
P3 = rt.empty((4, 4, 2049, 2))
report("config3 csd all-pairs 4 x 3.2M nperseg 4096",
       timeit(lambda i: rt.check(rt.lib.csd_allpairs(rt._ctx, p3, x3[i % NB].data_ptr(), 4, 3_200_000, 3_200_000, P3.data_ptr(), rt.stream()))),
       4 * 3_200_000 * 4 + 16 * 2049 * 8)
# ---- config 5: 40 channels x 1 M, nperseg sweep ----
for nps in (256, 512, 1024, 2048, 4096, 8192):
    p5 = rt.plan(nps, nps // 2, 500000, "hann", "density", "constant")
    P5 = rt.empty((40, 40, nps // 2 + 1, 2))
    ms = timeit(lambda i: rt.check(rt.lib.csd_allpairs(rt._ctx, p5, xs[i % NB].data_ptr(), 40, 1_000_000, 1_000_000, P5.data_ptr(), rt.stream())),
                iters=5, warm=2)
    T = rt.lib.plan_num_segments(p5, 1_000_000)
    flops = 8.0 * 1600 * (nps // 2 + 1) * T
    report(f"config5 csd all-pairs 40 x 1M nperseg {nps}", ms, 40 * 4e6 + 1600 * (nps // 2 + 1) * 8, pair_TFLOPs=round(flops / ms / 1e9, 2))
