"""Event timing of quantfilt on 40 x [256 x 3905] (inputs rotated)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
rt = api.Runtime()
g = torch.Generator(device=rt.device); g.manual_seed(0)
xs = [torch.rand((40, 256, 3905), device=rt.device, generator=g) for _ in range(3)]
out = torch.empty_like(xs[0])
for thr in (0.9, 0.5, 0.99, 0.05):
    run = lambda i: rt.check(rt.lib.quantfilt(rt._ctx, xs[i % 3].data_ptr(), 40, 256, 3905, 3905, thr, out.data_ptr(), None, None, rt.stream()))
    for i in range(3): run(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20): run(i)
    b.record(); torch.cuda.synchronize()
    print("quantfilt thr", thr, round(a.elapsed_time(b) / 20, 4), "ms", flush=True)
