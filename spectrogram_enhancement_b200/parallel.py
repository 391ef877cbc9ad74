"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* STFT -> denoise -> tiles is embarrassingly parallel over (shot, channel): `shot_range` hands every rank
  a contiguous range of shots (BASELINE config 4); there is no data-path collective.
* All-pairs CSD over a channel stack sharded by channel block (config 5) has ONE exchange step: every rank
  transforms its own channels (specgpu_csd_spectra), the spectra are all-gathered (NCCL over NVLink on the
  GPUs; gloo in the CPU tests), and every rank forms its row block of pairs (specgpu_csd_pairs).
* When every rank can read the whole record (one shot file), the same matrix shards by SEGMENT instead
  (`csd_allpairs_segment_sharded`): the Welch mean is a sum over segments, so each rank sums its own contiguous range
  of segments of all channels and the only exchange is an all-reduce of the [C, C, F] result (6.5 MB at 40 channels,
  nperseg 1024, against 160 MB of spectra per rank for the channel-block form).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import api

__all__ = ["shot_range", "channel_block", "segment_range", "frequency_block", "csd_allpairs_sharded",
           "csd_allpairs_freq_sharded", "csd_allpairs_segment_sharded", "pipeline_sharded"]


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
                         scaling="density", blocks=4, group=None, runtime=None):
    """Rows [rank*Cl, (rank+1)*Cl) of the all-pairs Welch CSD of the channel stack whose block `x_local[Cl, N]`
    this rank holds.  Returns (f, P_rows[Cl, C, F]) with C = world * Cl; P_rows[i, j] = csd(x_i, x_j).

    The segment axis is cut into `blocks`: block n's spectra are transformed and all-gathered on a side stream
    while the pair products of block n-1 run on the caller's stream (specgpu_csd_pairs_block accumulates)."""
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
    C = world * Cl
    F = rt.lib.plan_num_freqs(plan)
    T = rt.lib.plan_num_segments(plan, n)
    if T == 0:
        raise ValueError("record shorter than nperseg")
    hop = int(nperseg) - int(noverlap)
    ldf = (F + 15) & ~15                      # 128-byte rows: the pair kernel stages them with 16-byte copies
    nb = max(1, min(int(blocks), T))
    edges = [T * i // nb for i in range(nb + 1)]
    P = rt.empty((Cl, C, F, 2))
    cuda = rt.device.type == "cuda"
    main = torch.cuda.current_stream(rt.device) if cuda else None
    side = torch.cuda.Stream(device=rt.device) if cuda else None
    if cuda:
        side.wait_stream(main)
    ready, bufs = [], []
    for bi in range(nb):                      # producer: spectra of this rank's channels for block bi, then the exchange
        t0, t1 = edges[bi], edges[bi + 1]
        tb = t1 - t0
        X_all = rt.empty((C, tb, ldf, 2))
        X_loc = X_all[rank * Cl:(rank + 1) * Cl]
        xs = xd[:, t0 * hop:(t1 - 1) * hop + int(nperseg)]      # exactly segments t0 .. t1-1
        ctxm = torch.cuda.stream(side) if cuda else _null()
        with ctxm:
            rt.check(rt.lib.csd_spectra(rt._ctx, plan, xs.data_ptr(), Cl, xs.shape[1], api._ld(xd), X_loc.data_ptr(), ldf,
                                        rt.stream()))
            if world > 1:
                # in place: rank r's block of the gather buffer is X_loc itself (no staging copy)
                dist.all_gather_into_tensor(X_all.view(-1), X_loc.reshape(-1), group=group)
            if cuda:
                ev = torch.cuda.Event()
                ev.record(side)
                ready.append(ev)
        bufs.append((X_all, tb))
    for bi, (X_all, tb) in enumerate(bufs):   # consumer: pair products of block bi on the caller's stream
        if cuda:
            main.wait_event(ready[bi])
        rt.check(rt.lib.csd_pairs_block(rt._ctx, plan, X_all.data_ptr(), C, tb, T, ldf, rank * Cl, Cl, 1 if bi else 0,
                                        P.data_ptr(), rt.stream()))
        if cuda:
            X_all.record_stream(main)
    f = np.fft.rfftfreq(int(nperseg), 1.0 / fs)
    return f, rt.ret(torch.view_as_complex(P), as_torch)


def frequency_block(rank: int, world: int, n_freqs: int):
    """Bins [f0, f1) of `rank` when the one-sided spectrum is cut into `world` blocks of ceil(n_freqs / world) bins (the
    last blocks may be short or empty)."""
    w = -(-n_freqs // world)
    f0 = min(rank * w, n_freqs)
    return f0, min(f0 + w, n_freqs)


def csd_allpairs_freq_sharded(x_local, fs=1.0, window="hann", nperseg=256, noverlap=None, detrend="constant",
                              scaling="density", group=None, runtime=None):
    """All-pairs Welch CSD of a channel stack sharded by channel block, with the OUTPUT sharded by frequency: this rank
    holds `x_local[Cl, N]` and returns (f[f0:f1], P[C, C, f1 - f0]) for its block of bins `frequency_block(rank, ...)`,
    C = world * Cl.

    Pair products are independent per frequency bin, so instead of gathering every rank's full spectra everywhere
    (`csd_allpairs_sharded`: C T F values received per rank) each rank transforms its channels and sends rank h only
    bins `frequency_block(h)` of them -- one all-to-all in which a rank receives C T F / world values -- and then forms
    ALL pairs for its bins (specgpu_csd_pairs_bins)."""
    rt = runtime if runtime is not None else api.default_runtime()
    if noverlap is None:
        noverlap = int(nperseg) // 2
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    plan = rt.plan(nperseg, noverlap, fs, window, scaling, detrend)
    xd, as_torch = rt.to_device(x_local)
    if xd.dim() != 2:
        raise ValueError("csd_allpairs_freq_sharded expects x_local[Cl, N]")
    Cl, n = xd.shape
    C = world * Cl
    F = rt.lib.plan_num_freqs(plan)
    T = rt.lib.plan_num_segments(plan, n)
    if T == 0:
        raise ValueError("record shorter than nperseg")
    wf = -(-F // world)                         # bins per block
    ldb = (wf + 15) & ~15                       # 128-byte rows of a block
    # spectra of this rank's channels, written block-major by the transform itself: send[h] = bins of block h
    send = rt.empty((world, Cl, T, ldb, 2))
    rt.check(rt.lib.csd_spectra_blocked(rt._ctx, plan, xd.data_ptr(), Cl, n, api._ld(xd), send.data_ptr(), wf, ldb, rt.stream()))
    if world > 1:
        recv = torch.empty_like(send)            # [world (source rank)][Cl, T, ldb] == X_f[C, T, ldb], channels in rank order
        dist.all_to_all_single(recv, send, group=group)
        Xf = recv
    else:
        Xf = send
    f0, f1 = frequency_block(rank, world, F)
    P = rt.empty((C, C, max(f1 - f0, 0), 2))
    if f1 > f0:
        rt.check(rt.lib.csd_pairs_bins(rt._ctx, plan, Xf.data_ptr(), C, T, T, ldb, f0, f1 - f0, 0, P.data_ptr(), rt.stream()))
    f = np.fft.rfftfreq(int(nperseg), 1.0 / fs)[f0:f1]
    return f, rt.ret(torch.view_as_complex(P), as_torch)


def segment_range(rank: int, world: int, n_segments: int):
    """Contiguous, balanced range [t0, t1) of Welch segments owned by `rank` (same rule as `shot_range`)."""
    return shot_range(rank, world, n_segments)


def csd_allpairs_segment_sharded(x, fs=1.0, window="hann", nperseg=256, noverlap=None, detrend="constant",
                                 scaling="density", n_samples=None, n_channels=None, group=None, runtime=None):
    """All-pairs Welch CSD of the record `x[C, N]` sharded over segments.  `x` is the whole record (numpy or torch,
    readable by every rank) or a callable `x(lo, hi) -> [C, hi - lo]` that loads a sample range (then `n_samples` and
    `n_channels` describe the record).  Every rank returns the full (f, P[C, C, F]): rank r transforms and multiplies
    segments `segment_range(r, world, T)` of all channels, scaled by the total segment count, and the partial
    matrices are summed with one all-reduce."""
    rt = runtime if runtime is not None else api.default_runtime()
    if noverlap is None:
        noverlap = int(nperseg) // 2
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    plan = rt.plan(nperseg, noverlap, fs, window, scaling, detrend)
    hop = int(nperseg) - int(noverlap)
    if callable(x):
        if n_samples is None or n_channels is None:
            raise ValueError("a loader needs n_samples and n_channels")
        n, C = int(n_samples), int(n_channels)
    else:
        if x.ndim != 2:
            raise ValueError("csd_allpairs_segment_sharded expects x[C, N]")
        C, n = int(x.shape[0]), int(x.shape[1])
    F = rt.lib.plan_num_freqs(plan)
    T = rt.lib.plan_num_segments(plan, n)
    if T == 0:
        raise ValueError("record shorter than nperseg")
    t0, t1 = segment_range(rank, world, T)
    as_torch = (not callable(x)) and api._is_torch(x)
    P = rt.empty((C, C, F, 2))
    if t1 > t0:
        lo, hi = t0 * hop, (t1 - 1) * hop + int(nperseg)          # exactly segments t0 .. t1-1
        part = x(lo, hi) if callable(x) else x[:, lo:hi]
        if api._is_torch(part) and part.device == rt.device and part.dtype == torch.float32 and part.stride(-1) == 1:
            xs = part                                              # a view of the resident record: no copy
        else:
            xs, _ = rt.to_device(part)                             # only this rank's samples are uploaded
        ldf = (F + 15) & ~15
        X = rt.empty((C, t1 - t0, ldf, 2))
        rt.check(rt.lib.csd_spectra(rt._ctx, plan, xs.data_ptr(), C, xs.shape[1], api._ld(xs), X.data_ptr(), ldf, rt.stream()))
        rt.check(rt.lib.csd_pairs_block(rt._ctx, plan, X.data_ptr(), C, t1 - t0, T, ldf, 0, C, 0, P.data_ptr(), rt.stream()))
    else:
        P.zero_()
    if world > 1:
        dist.all_reduce(P, op=dist.ReduceOp.SUM, group=group)
    f = np.fft.rfftfreq(int(nperseg), 1.0 / fs)
    return f, rt.ret(torch.view_as_complex(P), as_torch)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def pipeline_sharded(load_shot, n_shots, spec_params=api.DEFAULT_SPEC_PARAMS, clip=True, tiles=False, runtime=None):
    """Run `api.pipeline` over this rank's shots: `load_shot(i) -> x[C, N]`.  Yields (shot index, results)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shot_range(rank, world, n_shots)
    for i in range(lo, hi):
        yield i, api.pipeline(load_shot(i), spec_params, clip=clip, tiles=tiles, runtime=runtime)
