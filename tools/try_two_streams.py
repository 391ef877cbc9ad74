"""Experiment: shots in flight on several streams (api.ShotStreams), with / without the STFT <-> projection interlock.
   NSTREAMS=2 INTERLOCK=1 SPECGPU_STFT_SMEM_PAD=20000 python tools/try_two_streams.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
dev = torch.device("cuda", 0)
NS = int(os.environ.get("NSTREAMS", 2))
ILK = os.environ.get("INTERLOCK", "0") == "1"
pool = api.ShotStreams(api.DEFAULT_SPEC_PARAMS, n=NS, device=dev, interlock=ILK)
g = torch.Generator(device=dev); g.manual_seed(0)
nb = 4
xs = [torch.randn((40, 1_000_000), device=dev, generator=g) for _ in range(nb)]
S = [pool.empty_image(40, 256, 3905) for _ in range(nb)]
D = [pool.empty_image(40, 256, 3905) for _ in range(nb)]
def step(i):
    pool.submit(xs[i % nb], S[i % nb], D[i % nb], clip=True)
for i in range(8): step(i)
pool.join(); torch.cuda.synchronize()
K = int(os.environ.get("K", 40))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record()
t0 = time.perf_counter()
for i in range(K): step(i)
t_host = time.perf_counter() - t0
pool.join()
e1.record(); torch.cuda.synchronize()
print("host enqueue us/shot", round(1e6 * t_host / K, 1))
print("streams", NS, "interlock", int(ILK), "pad", os.environ.get("SPECGPU_STFT_SMEM_PAD", "0"), "ms/shot", round(e0.elapsed_time(e1) / K, 4))
