"""Per-kernel timing of config 3 (4 chords x 3.2 M samples, nperseg 4096) through the library's event profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
rt = api.Runtime()
g = torch.Generator(device=rt.device); g.manual_seed(0)
for (C, n, nps, fs) in ((4, 3_200_000, 4096, 1.6e6), (40, 1_000_000, 1024, 5e5)):
    xs = [torch.randn((C, n), device=rt.device, generator=g) for _ in range(3)]
    plan = rt.plan(nps, nps // 2, fs, "hann", "density", "constant")
    P = rt.empty((C, C, nps // 2 + 1, 2))
    def run(i): rt.check(rt.lib.csd_allpairs(rt._ctx, plan, xs[i % 3].data_ptr(), C, n, n, P.data_ptr(), rt.stream()))
    for i in range(3): run(i)
    torch.cuda.synchronize()
    rt.profile(True)
    for i in range(10): run(i)
    prof = rt.profile_read(); rt.profile(False)
    print(C, n, nps, {k: round(v[0] / v[1], 4) for k, v in prof.items()})
