#!/usr/bin/env python
"""Multi-GPU check on real devices (torchrun, NCCL): channel-block-sharded all-pairs CSD (config 5 style) against
the single-GPU result and the oracle, plus timing of the sharded path.  Prints one JSON line from rank 0.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import spec_oracle as oc  # noqa: E402
from spectrogram_enhancement_b200 import api, parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rt = api.Runtime(device=dev)
C, n, nps = 40, 400_000, 1024
x = np.stack([oc.synth_ece(5, c, n=n) for c in range(C)])
lo, hi = parallel.channel_block(rank, world, C)
xl = torch.from_numpy(x[lo:hi]).to(dev)
f, P = parallel.csd_allpairs_sharded(xl, fs=500000.0, nperseg=nps, runtime=rt)
torch.cuda.synchronize()
gather = [torch.empty_like(P) for _ in range(world)]
dist.all_gather(gather, P)
res = {}
if rank == 0:
    Pall = torch.cat(gather).cpu().numpy()
    _, P1 = api.csd_allpairs(torch.from_numpy(x).to(dev), fs=500000.0, nperseg=nps, runtime=rt)
    P1 = P1.cpu().numpy()
    _, Pr = oc.csd_allpairs(x.astype(np.float64), fs=500000.0, nperseg=nps)
    res["max_rel_err_vs_oracle"] = float(np.abs(Pall - Pr).max() / np.abs(Pr).max())
    res["max_rel_diff_vs_single_gpu"] = float(np.abs(Pall - P1).max() / np.abs(P1).max())
    np.testing.assert_allclose(Pall, Pr, rtol=1e-4, atol=1e-6 * np.abs(Pr).max())
# timing: full-size config 5 (40 x 1M), sharded vs blocks=1 (no overlap)
n2 = 1_000_000
g = torch.Generator(device=dev)
g.manual_seed(rank)
xs = torch.randn((hi - lo, n2), device=dev, generator=g)
for blocks in (1, 4):
    for it in range(2):
        parallel.csd_allpairs_sharded(xs, fs=500000.0, nperseg=nps, blocks=blocks, runtime=rt)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5):
        parallel.csd_allpairs_sharded(xs, fs=500000.0, nperseg=nps, blocks=blocks, runtime=rt)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[f"ms_sharded_blocks{blocks}"] = float(t.item())
# segment sharding: every rank holds the whole record on its device, one all-reduce of the [C, C, F] matrix
xt = torch.from_numpy(x).to(dev)
_, Ps = parallel.csd_allpairs_segment_sharded(xt, fs=500000.0, nperseg=nps, runtime=rt)
if rank == 0:
    Ps = Ps.cpu().numpy()
    res["segment_sharded_max_rel_err_vs_oracle"] = float(np.abs(Ps - Pr).max() / np.abs(Pr).max())
    np.testing.assert_allclose(Ps, Pr, rtol=1e-4, atol=1e-6 * np.abs(Pr).max())
g.manual_seed(7)
xfull = [torch.randn((C, n2), device=dev, generator=g) for _ in range(2)]      # same record on every rank


def timed(fn, iters=10):
    for it in range(3):
        fn(it)
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for it in range(iters):
        fn(it)
    b.record()
    torch.cuda.synchronize()
    tt = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


res["ms_segment_sharded"] = timed(lambda i: parallel.csd_allpairs_segment_sharded(xfull[i % 2], fs=500000.0, nperseg=nps,
                                                                                   runtime=rt))
res["ms_single_gpu"] = timed(lambda i: api.csd_allpairs(xfull[i % 2], fs=500000.0, nperseg=nps, runtime=rt))
if rank == 0:
    res.update(world=world, C=C, nperseg=nps, n_check=n, n_timed=n2)
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
