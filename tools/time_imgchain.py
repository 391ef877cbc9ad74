"""Full-size cv2-chain calls (40 x [256 x 3905]) for a per-kernel launch list under ncu, plus event timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
rt = api.Runtime()
g = torch.Generator(device=rt.device); g.manual_seed(0)
dt = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
xs = [torch.rand((40, 256, 3905), device=rt.device, generator=g, dtype=dt) for _ in range(2)]
fns = dict(gaussblr=lambda x: api.gaussblr(x, (31, 3), runtime=rt), meansub=lambda x: api.meansub(x, runtime=rt),
           morph=lambda x: api.morph(x, runtime=rt))
iters = int(os.environ.get("ITERS", "10"))
if dt == torch.float32:
    fns["filter_chain_fused"] = lambda x: api.filter_chain(x, runtime=rt)
    fns["filter_chain_5_calls"] = lambda x: api.filter_chain(x, runtime=rt, fused=False)
for name, fn in fns.items():
    if name == "meansub" and dt != torch.float64:
        continue
    for i in range(2): fn(xs[i % 2])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(xs[i % 2])
    b.record(); torch.cuda.synchronize()
    print(name, str(dt), round(a.elapsed_time(b) / iters, 4), "ms", flush=True)
