import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
from spectrogram_enhancement_b200 import api
rt = api.Runtime()
plan = rt.plan_from_params(api.DEFAULT_SPEC_PARAMS)
g = torch.Generator(device=rt.device); g.manual_seed(1234)
import bench
x = bench.synth_on_device(torch, rt.device, 0, g)
S = rt.empty_image(40, 256, 3905); D = rt.empty_image(40, 256, 3905)
info = torch.zeros((40, 4), dtype=torch.int32, device=rt.device)
for it in (1, 2, 3, 4, 6, 8, 12, 200):
    rt.set_power_iterations(it)
    for _ in range(3): rt.pipeline_dev(plan, x, S, D, clip=True, info=info, fallback=False)
    rt.profile(True)
    for _ in range(10): rt.pipeline_dev(plan, x, S, D, clip=True, info=info, fallback=False)
    prof = rt.profile_read(); rt.profile(False)
    print(it, "gram_eig ms", round(prof["gram_eig"][0] / prof["gram_eig"][1], 4), "unconverged", int(info[:, 3].sum().item()))
