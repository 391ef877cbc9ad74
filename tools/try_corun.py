"""Experiment: does a DRAM-bound elementwise kernel hide under the (instruction-bound) STFT when both run at once?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
dev = torch.device("cuda", 0)
rt = api.Runtime(device=dev)
plan = rt.plan_from_params(api.DEFAULT_SPEC_PARAMS)
g = torch.Generator(device=dev); g.manual_seed(0)
x = torch.randn((40, 1_000_000), device=dev, generator=g)
S = rt.empty_image(40, 256, 3905)
a = torch.randn((40, 256, 3936), device=dev); b = torch.empty_like(a); c = torch.empty_like(a)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
K = 20
def stft():
    with torch.cuda.stream(s1):
        for _ in range(K): rt.specgr_dev(plan, x, S)
def mem():
    with torch.cuda.stream(s2):
        for _ in range(K):
            torch.mul(a, 2.0, out=b); torch.add(a, 1.0, out=c)      # 1 read + 1 write, twice: 640 MB per iteration
def timed(*fns):
    for f in fns: f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for f in fns: f()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
print("pad", os.environ.get("SPECGPU_STFT_SMEM_PAD", "0"), "specgr alone %.4f ms" % timed(stft), "mem alone %.4f ms" % timed(mem), "both %.4f ms" % timed(stft, mem))
