import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
rt = api.Runtime()
g = torch.Generator(device=rt.device); g.manual_seed(0)
x = torch.rand((40, 256, 3905), device=rt.device, generator=g)
out = torch.empty_like(x); thr = torch.empty((40, 3905), device=rt.device)
for i in range(3):
    rt.check(rt.lib.quantfilt(rt._ctx, x.data_ptr(), 40, 256, 3905, 3905, 0.9, out.data_ptr(), thr.data_ptr(), None, rt.stream()))
torch.cuda.synchronize()
print("ok")
