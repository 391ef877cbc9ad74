"""Experiment: two shots in flight on two streams (two contexts) vs one stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
dev = torch.device("cuda", 0)
NS = int(os.environ.get("NSTREAMS", 2))
rts = [api.Runtime(device=dev) for _ in range(NS)]
plans = [rt.plan_from_params(api.DEFAULT_SPEC_PARAMS) for rt in rts]
streams = [torch.cuda.Stream() for _ in range(NS)]
g = torch.Generator(device=dev); g.manual_seed(0)
nb = 4
xs = [torch.randn((40, 1_000_000), device=dev, generator=g) for _ in range(nb)]
S = [rts[0].empty((40, 256, 3905)) for _ in range(nb)]
D = [rts[0].empty((40, 256, 3905)) for _ in range(nb)]
def step(i):
    k = i % NS
    with torch.cuda.stream(streams[k]):
        rts[k].pipeline_dev(plans[k], xs[i % nb], S[i % nb], D[i % nb])
for i in range(8): step(i)
torch.cuda.synchronize()
K = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for st in streams: st.wait_event(e0)
for i in range(K): step(i)
for st in streams: torch.cuda.current_stream().wait_stream(st)
e1.record(); torch.cuda.synchronize()
print("streams", NS, "ms/shot", round(e0.elapsed_time(e1) / K, 4))
