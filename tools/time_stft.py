"""Time specgr's STFT kernel alone (device-resident, 40 x 1M, rotating inputs); for A/B runs of library builds (SPECGPU_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
rt = api.Runtime()
plan = rt.plan_from_params(api.DEFAULT_SPEC_PARAMS)
g = torch.Generator(device=rt.device); g.manual_seed(0)
xs = [torch.randn((40, 1_000_000), device=rt.device, generator=g) for _ in range(3)]
ldt = int(os.environ.get("LDT", 3936))
S = rt.empty((40, 257, ldt))[:, :, :3905]
def run(i): rt.spectrogram_dev(plan, xs[i % 3], S)
for i in range(3): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): run(i)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("SPECGPU_LIB", "default"), "PSD-mode stft ms:", round(e0.elapsed_time(e1) / 20, 4))
