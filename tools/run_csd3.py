#!/usr/bin/env python
"""Config 3 (4 chords x 3.2M samples, nperseg 4096) per-kernel timing."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
rt = api.Runtime()
g = torch.Generator(device=rt.device); g.manual_seed(0)
xs = [torch.randn((4, 3_200_000), device=rt.device, generator=g) for _ in range(4)]
kw = dict(fs=1.6e6, nperseg=4096, runtime=rt)
for i in range(3): api.csd_allpairs(xs[i % 4], **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): api.csd_allpairs(xs[i % 4], **kw)
e1.record(); torch.cuda.synchronize()
rt.profile(True)
for i in range(4): api.csd_allpairs(xs[i % 4], **kw)
prof = rt.profile_read(); rt.profile(False)
print(json.dumps({"ms": round(e0.elapsed_time(e1) / 20, 4), "kernels_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items()}}))
