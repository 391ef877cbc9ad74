#!/usr/bin/env python
"""Per-kernel timing of denoiseSignal(use_optimal=True) on 40 x [256 x 3905] (BASELINE config 2's second mode)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import spec_oracle as oc  # noqa: E402
from spectrogram_enhancement_b200 import api  # noqa: E402

rt = api.Runtime()
x = np.stack([oc.synth_ece(3, c, n=1_000_000) for c in range(40)])
S, _, _ = api.spectrogram_batch(torch.from_numpy(x).cuda(), oc.DEFAULT_SPEC_PARAMS, runtime=rt)
S = S.contiguous()
for mode in ("optimal", "default"):
    kw = dict(use_optimal=True) if mode == "optimal" else {}
    for _ in range(2):
        D = api.denoiseSignal(S, runtime=rt, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        D = api.denoiseSignal(S, runtime=rt, **kw)
    e1.record()
    torch.cuda.synchronize()
    rt.profile(True)
    D, s, info = api.denoiseSignal(S, return_info=True, runtime=rt, **kw)
    prof = rt.profile_read()
    rt.profile(False)
    print(json.dumps({"mode": mode, "ms_per_call": e0.elapsed_time(e1) / 3, "kernels_ms": {k: round(v[0], 3) for k, v in prof.items()},
                      "num_sing": info[:4, 2].tolist()}))
