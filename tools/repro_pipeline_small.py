"""Debug aid: the smallest pipeline case, for compute-sanitizer runs on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import parity_cases as pc  # noqa: E402
from oracle import spec_oracle as oc  # noqa: E402
from spectrogram_enhancement_b200 import api  # noqa: E402

rt = api.Runtime()
sp = dict(oc.DEFAULT_SPEC_PARAMS, nperseg=int(sys.argv[1]) if len(sys.argv) > 1 else 32)
sp["noverlap"] = sp["nperseg"] // 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
x = pc.signals(2, n)
S, D, tiles, info = api.pipeline(x, sp, clip=True, tiles=True, tile=64, return_info=True, runtime=rt)
torch.cuda.synchronize()
print("ok", S.shape, D.shape, tiles.shape, info)
