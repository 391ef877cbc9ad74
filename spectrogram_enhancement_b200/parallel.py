"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* STFT -> denoise -> tiles is embarrassingly parallel over (shot, channel): `shot_range` hands every rank
  a contiguous range of shots (BASELINE config 4); there is no data-path collective.
* All-pairs CSD over a channel stack sharded by channel block (config 5) has ONE exchange step: every rank
  transforms its own channels (specgpu_csd_spectra), the spectra are all-gathered (NCCL over NVLink on the
  GPUs; gloo in the CPU tests), and every rank forms its row block of pairs (specgpu_csd_pairs).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import api

__all__ = ["shot_range", "channel_block", "csd_allpairs_sharded", "pipeline_sharded"]


def shot_range(rank: int, world: int, n_shots: int):
    """Contiguous, balanced shot range [lo, hi) of `rank`: the first n_shots % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world) or n_shots < 0:
        raise ValueError("bad rank/world/n_shots")
    base, extra = divmod(n_shots, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def channel_block(rank: int, world: int, n_channels: int):
    """Channel block [lo, hi) of `rank` for the CSD exchange; all blocks must have the same size."""
    if n_channels % world != 0:
        raise ValueError(f"{n_channels} channels do not split evenly over {world} ranks")
    per = n_channels // world
    return rank * per, (rank + 1) * per


def csd_allpairs_sharded(x_local, fs=1.0, window="hann", nperseg=256, noverlap=None, detrend="constant",
                         scaling="density", group=None, runtime=None):
    """Rows [rank*Cl, (rank+1)*Cl) of the all-pairs Welch CSD of the channel stack whose block `x_local[Cl, N]`
    this rank holds.  Returns (f, P_rows[Cl, C, F]) with C = world * Cl; P_rows[i, j] = csd(x_i, x_j)."""
    rt = runtime if runtime is not None else api.default_runtime()
    if noverlap is None:
        noverlap = int(nperseg) // 2
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    plan = rt.plan(nperseg, noverlap, fs, window, scaling, detrend)
    xd, as_torch = rt.to_device(x_local)
    if xd.dim() != 2:
        raise ValueError("csd_allpairs_sharded expects x_local[Cl, N]")
    Cl, n = xd.shape
    F = rt.lib.plan_num_freqs(plan)
    T = rt.lib.plan_num_segments(plan, n)
    if T == 0:
        raise ValueError("record shorter than nperseg")
    ldf = (F + 1) & ~1
    X_all = rt.empty((world * Cl, T, ldf, 2))
    X_loc = X_all[rank * Cl:(rank + 1) * Cl]          # transform straight into this rank's slot
    rt.check(rt.lib.csd_spectra(rt._ctx, plan, xd.data_ptr(), Cl, n, api._ld(xd), X_loc.data_ptr(), ldf, rt.stream()))
    if world > 1:
        dist.all_gather_into_tensor(X_all, X_loc.clone(), group=group)
    P = rt.empty((Cl, world * Cl, F, 2))
    rt.check(rt.lib.csd_pairs(rt._ctx, plan, X_all.data_ptr(), world * Cl, T, ldf, rank * Cl, Cl, P.data_ptr(), rt.stream()))
    f = np.fft.rfftfreq(int(nperseg), 1.0 / fs)
    return f, rt.ret(torch.view_as_complex(P), as_torch)


def pipeline_sharded(load_shot, n_shots, spec_params=api.DEFAULT_SPEC_PARAMS, clip=True, tiles=False, runtime=None):
    """Run `api.pipeline` over this rank's shots: `load_shot(i) -> x[C, N]`.  Yields (shot index, results)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shot_range(rank, world, n_shots)
    for i in range(lo, hi):
        yield i, api.pipeline(load_shot(i), spec_params, clip=clip, tiles=tiles, runtime=runtime)
