"""Stress test of the STFT's dynamic tile walk: thousands of back-to-back pipeline calls (one context, and two contexts on
two streams with the dynamic walk forced on in both), results compared bit for bit with a first reference run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_enhancement_b200 import api
dev = torch.device("cuda", 0)
N = int(os.environ.get("N", 3000))
g = torch.Generator(device=dev); g.manual_seed(3)
xs = [torch.randn((40, 1_000_000), device=dev, generator=g) for _ in range(2)] + [torch.randn((7, 333_333), device=dev, generator=g)]
rt = api.Runtime(device=dev)
plan = rt.plan_from_params(api.DEFAULT_SPEC_PARAMS)
ref = []
for x in xs:
    S, D = rt.pipeline_dev(plan, x, clip=True, static_tiles=True)
    torch.cuda.synchronize()
    ref.append((S.clone(), D.clone()))
bad = 0
for i in range(N):
    k = i % len(xs)
    S, D = rt.pipeline_dev(plan, xs[k], clip=True)            # dynamic walk
    if i % 50 == 0:
        torch.cuda.synchronize()
        if not (torch.equal(S, ref[k][0]) and torch.equal(D, ref[k][1])):
            bad += 1
torch.cuda.synchronize()
print("one context:", N, "calls, mismatches", bad)
rts = [api.Runtime(device=dev) for _ in range(2)]
plans = [r.plan_from_params(api.DEFAULT_SPEC_PARAMS) for r in rts]
sts = [torch.cuda.Stream() for _ in range(2)]
outs = [[(r.empty_image(x.shape[0], 256, ref[k][0].shape[-1]), r.empty_image(x.shape[0], 256, ref[k][0].shape[-1])) for k, x in enumerate(xs)] for r in rts]
bad2 = 0
for i in range(N):
    j, k = i % 2, (i // 2) % len(xs)
    with torch.cuda.stream(sts[j]):
        rts[j].pipeline_dev(plans[j], xs[k], outs[j][k][0], outs[j][k][1], clip=True, static_tiles=False)
    if i % 100 == 99:
        torch.cuda.synchronize()
        for jj in range(2):
            for kk in range(len(xs)):
                if i > 2 * len(xs) * 2 and not (torch.equal(outs[jj][kk][0], ref[kk][0]) and torch.equal(outs[jj][kk][1], ref[kk][1])):
                    bad2 += 1
torch.cuda.synchronize()
print("two contexts, two streams:", N, "calls, mismatches", bad2)
assert bad == 0 and bad2 == 0
