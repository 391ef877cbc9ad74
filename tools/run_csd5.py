#!/usr/bin/env python
"""Config 5 (all-pairs CSD, 40 channels x 1M samples) timing with the library's per-kernel CUDA events: one JSON line per nperseg."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from spectrogram_enhancement_b200 import api  # noqa: E402

rt = api.Runtime()
g = torch.Generator(device=rt.device)
g.manual_seed(0)
x = torch.randn((40, 1_000_000), device=rt.device, generator=g)
for nps in [int(a) for a in sys.argv[1:]] or [1024]:
    kw = dict(fs=500000.0, nperseg=nps, runtime=rt)
    for _ in range(3):
        api.csd_allpairs(x, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        api.csd_allpairs(x, **kw)
    e1.record()
    torch.cuda.synchronize()
    rt.profile(True)
    api.csd_allpairs(x, **kw)
    prof = rt.profile_read()
    rt.profile(False)
    print(json.dumps({"nperseg": nps, "ms": round(e0.elapsed_time(e1) / 10, 4), "kernels_ms": {k: round(v[0], 4) for k, v in prof.items()}}))
