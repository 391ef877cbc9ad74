"""One pass over the widened rows (config 5 CSD at 40 channels, the fused cv2 filter chain) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
rt = api.Runtime()
g = torch.Generator(device=rt.device); g.manual_seed(0)
x = torch.randn((40, 1_000_000), device=rt.device, generator=g)
S = torch.rand((40, 256, 3905), device=rt.device, generator=g)
for _ in range(int(os.environ.get("ITERS", "2"))):
    api.csd_allpairs(x, fs=5e5, nperseg=1024, runtime=rt)
    api.filter_chain(S, runtime=rt)
torch.cuda.synchronize()
print("ok")
