"""Sweep HostPipeline (groups, streams) for the host-to-host throughput of one 40-channel shot."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
N_CH, N = 40, 1_000_000
xh = torch.randn((N_CH, N)).pin_memory()
dh = torch.empty((N_CH, 256, 3905)).pin_memory()
# raw PCIe: H2D alone, D2H alone, both at once
xd = torch.empty((N_CH, N), device="cuda"); dd = torch.empty((N_CH, 256, 3905), device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
print("H2D 160MB: %.2f ms" % (1e3 * t(lambda: xd.copy_(xh, non_blocking=True))))
print("D2H 160MB: %.2f ms" % (1e3 * t(lambda: dh.copy_(dd, non_blocking=True))))
def both():
    with torch.cuda.stream(s1): xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2): dh.copy_(dd, non_blocking=True)
print("both     : %.2f ms" % (1e3 * t(both)))
xh2 = torch.randn((N_CH, N)).pin_memory()
dh2 = torch.empty((N_CH, 256, 3905)).pin_memory()
for groups, streams in ((2, 2), (3, 2), (3, 3), (4, 2), (4, 3), (4, 4), (5, 2), (5, 3), (6, 3), (8, 3)):
    hp = api.HostPipeline(api.DEFAULT_SPEC_PARAMS, channels=N_CH, samples=N, groups=groups, streams=streams)
    ms = 1e3 * t(lambda: hp.run(xh, dh), n=8)
    # as bench.py times it: shots submitted back to back (two host buffer sets), one wait at the end
    def many(k=10):
        hs = [hp.submit(xh if i % 2 == 0 else xh2, dh if i % 2 == 0 else dh2) for i in range(k)]
        for h in hs:
            h.synchronize()
    many(2)
    torch.cuda.synchronize(); t0 = time.perf_counter(); many(10); dt = (time.perf_counter() - t0) / 10
    print(f"groups {groups:2d} streams {streams}: one at a time {ms:.2f} ms/shot = {N_CH * N / ms / 1e6:.2f} G samples/s; "
          f"back to back {1e3 * dt:.2f} ms/shot = {N_CH * N / dt / 1e9:.2f} G samples/s")
    del hp
