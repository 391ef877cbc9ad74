#!/usr/bin/env python
"""Headline benchmark: input samples/s from raw ECE signal to denoised spectrogram (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the whole hot path (STFT -> |X|^2 -> log -> min-max -> Nyquist drop -> SVD
denoise -> clip) over one synthetic shot of 40 ECE channels x 1 000 000 float32 samples with the
reference's spec_params (BASELINE config 2).  With N > 1 (torchrun, one rank per GPU) every rank
processes its own shots (config 4's shot sharding: no data-path collective, weak scaling) and the
value is the whole-job aggregate over the max-over-ranks device time.

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the
public host API with pinned HOST buffers, H2D of the shot and D2H of the denoised spectrogram inside
the timed region; `roofline` is for the dominant kernel, timed live with CUDA events recorded by the
library around each launch during the timed steps; `cpu_baseline` is the reference's per-channel
arithmetic (scipy.signal.spectrogram + numpy SVD, oracle.reference_channel) on a bounded sample of
the same workload on this host.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_CH = 40
N_SAMP = 1_000_000
SP = {"nperseg": 512, "noverlap": 256, "fs": 500000, "window": "hamm", "scaling": "density", "detrend": "linear",
      "eps": 1e-11}
ROWS = SP["nperseg"] // 2
NSEG = (N_SAMP - SP["noverlap"]) // (SP["nperseg"] - SP["noverlap"])
METRIC = "input samples/sec -> denoised spectrogram"
UNIT = "samples/s"
WORKLOAD = ("config2: one shot, 40 ECE channels x 1M samples @500 kHz, nperseg=512 noverlap=256 hamm linear-detrend "
            "density; STFT+log+min-max+SVD(top-1 removed)+clip")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self.ok:
            self.th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.th.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ---------------------------------------------------------------------------------------------------
def _ref_worker(args):
    shot, ch = args
    from oracle import spec_oracle as oc
    try:
        from threadpoolctl import threadpool_limits
        ctxm = threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        import contextlib
        ctxm = contextlib.nullcontext()
    x = oc.synth_ece(shot, ch, n=N_SAMP)
    t0 = time.perf_counter()
    with ctxm:
        S, D = oc.reference_channel(x, SP)
    return time.perf_counter() - t0, float(D.sum())


def cpu_serial_baseline(nch):
    """The reference as it runs: one process, serial channel loop (pipeline_data.py:92-95), BLAS threads as found."""
    from oracle import spec_oracle as oc
    xs = [oc.synth_ece(0, c, n=N_SAMP) for c in range(nch)]
    t0 = time.perf_counter()
    for x in xs:
        oc.reference_channel(x, SP)
    dt = time.perf_counter() - t0
    return nch * N_SAMP / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU arithmetic on all host cores (one channel per worker process,
    BLAS pinned to 1 thread per worker); each step is a bounded sample of `workers` channels."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_step = workers
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for w in range(args.warmup):
            pool.map(_ref_worker, [(100 + w, c) for c in range(min(per_step, workers))])
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_ref_worker, [(s, c) for c in range(per_step)])
        dt = time.perf_counter() - t0
    value = args.steps * per_step * N_SAMP / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "channels_per_step": per_step, "samples_per_channel": N_SAMP},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{per_step} channels x 1M samples per step ({workers} worker processes, 1 BLAS thread "
                                   "each): scipy.signal.spectrogram + log/min-max + np.linalg.svd truncation + clip, "
                                   "exactly the reference's calls (oracle.reference_channel)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
ALGO_BYTES = {   # algorithmic HBM bytes of one launch over B channels (DESIGN.md "kernels")
    "stft_kernel": lambda B: B * (4 * N_SAMP + 4 * ROWS * NSEG),          # read x, write log-PSD (Nyquist dropped)
    "lognorm": lambda B: B * 8 * ROWS * NSEG,                              # read + write the image
    "gram_tc": lambda B: B * 4 * ROWS * NSEG,                              # read the image once
    "gram_simt": lambda B: B * 4 * ROWS * NSEG,
    "svd_project": lambda B: B * 8 * ROWS * NSEG,                          # read S, write D
    "svd_rank1": lambda B: B * 12 * ROWS * NSEG,                           # read log image, write S and D
    "patch": lambda B: B * 8 * ROWS * (NSEG // 128) * 128,
}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed ncu summary
    (profiles/rNN*_ncu_raw.csv, written by tools/summarize_profiles.py from an `ncu --set full` capture)."""
    import csv
    import glob
    import re
    best = None
    for path in glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_raw.csv")):
        m = re.match(r"r(\d+)([a-z]*)_", os.path.basename(path))
        if not m:
            continue
        try:
            rows = list(csv.DictReader(open(path)))
        except Exception:
            continue
        for r in rows:
            name = r.get("Kernel Name", "")
            if kernel in name and "dram__bytes_read.sum [Mbyte]" in r:
                try:
                    val = (float(r["dram__bytes_read.sum [Mbyte]"]) + float(r["dram__bytes_write.sum [Mbyte]"])) * 1e6
                except ValueError:
                    continue
                key = (int(m.group(1)), m.group(2))
                if best is None or key > best[0]:
                    best = (key, val, os.path.basename(path))
                break
    return (best[1], best[2]) if best else (None, None)


def synth_on_device(torch, device, shot, gen):
    """Device-side synthetic shot with the oracle's signal model (0.5*chirp + N(0,1)); parity on these exact
    bytes is checked in run_ours() by reading a channel back."""
    t = torch.arange(N_SAMP, device=device, dtype=torch.float64) / SP["fs"]
    f0, f1 = 20e3, 120e3
    base = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / float(t[-1]) * t * t)
    ph = 2 * np.pi * torch.arange(N_CH, device=device, dtype=torch.float64)[:, None] / N_CH
    x = 0.5 * torch.cos(base[None, :] + ph).to(torch.float32)
    x += torch.randn((N_CH, N_SAMP), device=device, dtype=torch.float32, generator=gen)
    return x.contiguous()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from spectrogram_enhancement_b200 import api, parallel

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libspecgpu has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rt = api.Runtime(device=device)
    plan = rt.plan_from_params(SP)
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + rank)
    nbuf = 3   # rotate over 3 resident shots (3 x 160 MB in, 2 x 160 MB out each) so no step finds its data in L2
    xs = [synth_on_device(torch, device, rank * 1000 + i, gen) for i in range(nbuf)]
    # the buffers the public API itself allocates (Runtime.empty_image: rows pitched to 32 floats; the C ABI takes any ld)
    S = [rt.empty_image(N_CH, ROWS, NSEG) for _ in range(nbuf)]
    D = [rt.empty_image(N_CH, ROWS, NSEG) for _ in range(nbuf)]
    ldt = int(S[0].stride(1))
    info = torch.zeros((N_CH, 4), dtype=torch.int32, device=device)
    fallback = os.environ.get("SPECGPU_BENCH_FALLBACK", "1") != "0"     # A/B knob; the API default is on

    def step(i):
        b = i % nbuf
        rt.pipeline_dev(plan, xs[b], S[b], D[b], clip=True, info=info, fallback=fallback)

    # The timed job keeps `inflight` shots in flight (api.ShotStreams: shot i on worker stream i % inflight, one library
    # context each), as the reference's shot loop allows -- shots are independent.  One shot at a time is reported next to it.
    inflight = max(1, env_int("SPECGPU_BENCH_INFLIGHT", 2))
    pool = api.ShotStreams(SP, n=inflight, device=device)
    infos = [torch.zeros((N_CH, 4), dtype=torch.int32, device=device) for _ in range(inflight)]

    def pstep(i):
        b = i % nbuf
        pool.submit(xs[b], S[b], D[b], clip=True, info=infos[i % inflight], fallback=fallback)

    # ---- correctness on these exact bytes (rank 0, untimed): one channel against the oracle ----
    step(0)
    torch.cuda.synchronize()
    parity = None
    if rank == 0:
        from oracle import spec_oracle as oc
        xc = xs[0][3].cpu().numpy()
        Sr, _, _ = oc.specgr_array(xc.astype(np.float64), SP)
        Sg = S[0][3].contiguous().cpu().numpy()
        Dg = D[0][3].contiguous().cpu().numpy()
        Dr = oc.clip(oc.denoiseSignal(Sr))                 # the pure oracle chain, not the oracle on our S
        parity = {"spec_max_abs_err": float(np.abs(Sg - Sr).max()),
                  "denoise_max_err_rel_to_max": float(np.abs(Dg - Dr).max() / np.abs(Dr).max())}
        if not os.environ.get("SPECGPU_BENCH_NOASSERT"):       # (timing ablations produce wrong results on purpose)
            assert parity["spec_max_abs_err"] < 1e-4 and parity["denoise_max_err_rel_to_max"] < 1e-3, parity
            assert int(info[:, 3].max().item()) == 0

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    # ---- device-resident timing (config 2) ----
    clk = ClockSampler(local)
    clk.__enter__()                      # sampled every 2 ms from here (before the warm-up: the sampler thread's start-up
                                         # stays out of the 5 ms timed region) to the end of the e2e loop
    for i in range(max(args.warmup, 3)):
        pstep(i)
    pool.join()
    barrier()
    l0 = pool.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        pstep(i)
    pool.join()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pool.launch_count() - l0
    assert os.environ.get("SPECGPU_BENCH_NOASSERT") or all(int(t[:, 3].max().item()) == 0 for t in infos)
    # the same K steps one shot at a time on one stream
    ms_one = max_over_ranks(timed(step, args.steps, 3))
    # per-kernel durations: the same K steps once more with the library's per-launch CUDA events switched on
    # (the event pairs cost ~8 % of a step, so the timed region above runs without them)
    rt.profile(True)
    for i in range(args.steps):
        step(i)
    prof = rt.profile_read()
    rt.profile(False)
    lt = torch.tensor([launches], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms_max = max_over_ranks(ms)
    value = world * args.steps * N_CH * N_SAMP / (ms_max * 1e-3)

    # ---- the same with DENSE image rows (3905 floats: no 16-byte alignment -> plain-store / 4-byte-copy paths) ----
    Sd = [torch.empty((N_CH, ROWS, NSEG), device=device) for _ in range(nbuf)]
    Dd = [torch.empty((N_CH, ROWS, NSEG), device=device) for _ in range(nbuf)]
    ms_dense = max_over_ranks(timed(lambda i: rt.pipeline_dev(plan, xs[i % nbuf], Sd[i % nbuf], Dd[i % nbuf], clip=True, info=info,
                                                              fallback=fallback), args.steps, 3))
    value_dense = world * args.steps * N_CH * N_SAMP / (ms_dense * 1e-3)
    del Sd, Dd

    # ---- config 4: 1000 shots x 40 channels streamed through resident buffers, sharded by shot, tiles written ----
    cfg4 = None
    if not args.no_config4:
        shots_total = int(os.environ.get("SPECGPU_BENCH_SHOTS", 1000))
        lo, hi = parallel.shot_range(rank, world, shots_total)
        ntile = NSEG // 128

        tl = [rt.empty((N_CH * ntile, ROWS, 128)) for _ in range(nbuf)]

        def shot(i):
            b = i % nbuf
            pool.submit(xs[b], S[b], D[b], clip=True, tiles=tl[b], tile_w=128, ntiles=ntile, info=infos[i % inflight],
                        fallback=fallback)

        def shots(_):
            for i in range(hi - lo):
                shot(i)
            pool.join()

        for i in range(3):
            shot(i)
        pool.join()
        l4 = pool.launch_count()
        clk4 = ClockSampler(local)
        with clk4:
            ms4 = max_over_ranks(timed(shots, 1, 0))
        launches4 = pool.launch_count() - l4
        ok4 = None
        if rank == 0:
            from oracle import spec_oracle as oc
            last = (hi - lo - 1)
            c = 17
            Sr, _, _ = oc.specgr_array(xs[last % nbuf][c].cpu().numpy().astype(np.float64), SP)
            Tr = oc.patch([oc.clip(oc.denoiseSignal(Sr))], 128, ntile)
            Tg = tl[last % nbuf][c * ntile:(c + 1) * ntile].cpu().numpy()
            err = float(np.abs(Tg - Tr).max() / np.abs(Tr).max())
            ok4 = {"shot": int(lo + last), "channel": c, "tiles_max_err_rel_to_max": err}
            assert err < 1e-3, ok4
        cfg4 = {"workload": "config4: 1000 synthetic shots x 40 channels x 1M samples, contiguous shot ranges per GPU "
                            "(parallel.shot_range), pipeline + clip + 30 tiles of 256x128 per channel written; "
                            f"{nbuf} resident input shots rotated, every shot's S, D and tiles written to HBM; "
                            f"{inflight} shots in flight (api.ShotStreams)",
                "shots_total": shots_total, "shots_this_rank": hi - lo, "value": shots_total * N_CH * N_SAMP / (ms4 * 1e-3),
                "unit": UNIT, "ms_per_shot_per_gpu": ms4 / max(hi - lo, 1), "seconds": ms4 * 1e-3, "scaling": "strong",
                "gpu_launches_this_rank": int(launches4), "bytes_per_sample": 16, "clocks": clk4.summary(),
                "parity_sampled": ok4}
        del tl

    # ---- config 2's second mode: denoiseSignal(use_optimal=True) on the 40 spectrogram images of shot 0 (rank 0 at N = 1) ----
    useopt = None
    if world == 1 and not args.no_config4:
        Simg = S[0]
        kw_opt = dict(use_optimal=True, runtime=rt)
        for _ in range(2):
            api.denoiseSignal(Simg, **kw_opt)
        ms_opt = timed(lambda i: api.denoiseSignal(Simg, **kw_opt), 5, 0) / 5
        _, _, info_opt = api.denoiseSignal(Simg, return_info=True, **kw_opt)
        ns = info_opt[:, 2].tolist() if hasattr(info_opt, "tolist") else [int(v) for v in info_opt[:, 2]]
        useopt = {"workload": "denoiseSignal(use_optimal=True) (denoising_by_svd.ipynb:210-217) on 40 x [256 x 3905]: float64 Gram, "
                              "all singular values (tridiagonalisation + bisection), Gavish-Donoho cut, leading vectors, projection",
                  "ms_per_shot": ms_opt, "num_sing_min_max": [int(min(ns)), int(max(ns))]}

    # ---- config 5: all-pairs Welch CSD of 40 channels x 1M samples, nperseg 1024 ----
    cfg5 = None
    if not args.no_csd5:
        g5 = torch.Generator(device=device)
        g5.manual_seed(77)                                  # the same record on every rank
        x5 = synth_on_device(torch, device, 5, g5)
        nps = 1024
        kw = dict(fs=float(SP["fs"]), nperseg=nps, runtime=rt)
        res5 = {}
        P1 = None
        if rank == 0 or world == 1:
            _, P1 = api.csd_allpairs(x5, **kw)
        res5["single_gpu_ms"] = timed(lambda i: api.csd_allpairs(x5, **kw), 10, 3) / 10
        if world > 1:
            clo, chi = parallel.channel_block(rank, world, N_CH)
            xl = x5[clo:chi].contiguous()
            legs = {"channel_block_allgather": lambda i: parallel.csd_allpairs_sharded(xl, **kw),
                    "frequency_block_alltoall": lambda i: parallel.csd_allpairs_freq_sharded(xl, **kw),
                    "segment_allreduce": lambda i: parallel.csd_allpairs_segment_sharded(x5, **kw)}
            for name, fn in legs.items():
                res5[name + "_ms"] = max_over_ranks(timed(fn, 10, 3) / 10)
            # parity of every sharding against the single-GPU matrix (itself checked against the oracle in tests/)
            _, Pa = parallel.csd_allpairs_sharded(xl, **kw)
            _, Pf = parallel.csd_allpairs_freq_sharded(xl, **kw)
            _, Ps = parallel.csd_allpairs_segment_sharded(x5, **kw)
            ga = [torch.empty_like(Pa) for _ in range(world)]
            dist.all_gather(ga, Pa.contiguous())
            F = nps // 2 + 1
            wf = -(-F // world)
            Pfp = torch.zeros((N_CH, N_CH, wf), dtype=Pf.dtype, device=device)
            Pfp[:, :, :Pf.shape[-1]] = Pf
            gf = [torch.empty_like(Pfp) for _ in range(world)]
            dist.all_gather(gf, Pfp)
            if rank == 0:
                scale = float(P1.abs().max().item())
                res5["parity_rel_to_max"] = {
                    "channel_block_allgather": float((torch.cat(ga) - P1).abs().max().item()) / scale,
                    "frequency_block_alltoall": float((torch.cat(gf, dim=-1)[:, :, :F] - P1).abs().max().item()) / scale,
                    "segment_allreduce": float((Ps - P1).abs().max().item()) / scale}
                assert max(res5["parity_rel_to_max"].values()) < 1e-5, res5
        cfg5 = {"workload": f"config5: all-pairs CSD, 40 channels x 1M samples, hann, nperseg {nps}, noverlap {nps // 2}; "
                            "times include the collective (CUDA events, max over ranks)", **res5}
        del x5

    # ---- end to end through the public API with host buffers ----
    # api.HostPipeline: the shot sits in pinned host memory; per step H2D of the 40 channels, the whole path, D2H of
    # the denoised spectrogram (channel groups on 3 streams so uploads, kernels and downloads overlap).
    e2e_steps = max(2, min(args.steps, 20))
    if args.no_e2e:
        e2e_steps = 0
    e2e_ok, e2e_launches, e2e_value, e2e_ceiling = None, 0, None, None
    if e2e_steps:
        xh = [torch.empty((N_CH, N_SAMP), dtype=torch.float32).pin_memory() for _ in range(2)]
        for i in range(2):
            xh[i].copy_(xs[i])
        dh = [torch.empty((N_CH, ROWS, NSEG), dtype=torch.float32).pin_memory() for _ in range(2)]
        hp = api.HostPipeline(SP, channels=N_CH, samples=N_SAMP, groups=8, streams=3, clip=True, device=device)
        for i in range(2):
            hp.run(xh[i % 2], dh[i % 2])
        e2e_ok = bool(torch.allclose(dh[1][3], D[1][3].cpu(), rtol=0, atol=2e-5))   # dh[1] holds shot xs[1], as D[1] does

        def e2e_loop(copy_only):
            barrier()
            t0 = time.perf_counter()
            # shots are submitted back to back (double-buffered host results): every step uploads its shot from pinned
            # memory and downloads its denoised spectrogram; the download of step i overlaps the upload of step i+1, and
            # step i is waited for right after step i+1 has been enqueued
            prev = None
            for i in range(e2e_steps):
                ev = hp.submit(xh[i % 2], dh[i % 2], copy_only=copy_only)
                if prev is not None:
                    prev.synchronize()
                prev = ev
            prev.synchronize()
            barrier()
            return max_over_ranks(time.perf_counter() - t0)

        l0e = hp.launch_count()
        e2e_s = e2e_loop(False)
        e2e_launches = hp.launch_count() - l0e
        e2e_value = world * e2e_steps * N_CH * N_SAMP / e2e_s
        # the same buffers, streams and copies with the kernels skipped: what the host link alone allows
        e2e_ceiling = world * e2e_steps * N_CH * N_SAMP / e2e_loop(True)
    clk.__exit__()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    kernels = {k: {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1], "share": v[0] / max(sum(x[0] for x in prof.values()), 1e-9)}
               for k, v in prof.items()}
    dom = max((k for k in prof if k in ALGO_BYTES), key=lambda k: prof[k][0], default=None)
    roof = None
    if dom is not None:
        dur_s = prof[dom][0] / max(prof[dom][1], 1) * 1e-3
        ach = ALGO_BYTES[dom](N_CH) / dur_s / 1e9
        traffic, traffic_src = ncu_traffic(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": f"ncu --set full capture, profiles/{traffic_src}" if traffic_src else None,
                "peak_source": peak_src, "timing": "CUDA events around each launch, second pass of the same K steps",
                "algorithmic_bytes_per_launch": ALGO_BYTES[dom](N_CH), "ms_per_launch": dur_s * 1e3}
    pipeline_gbs = 12.0 * N_CH * N_SAMP * args.steps / (ms * 1e-3) / 1e9      # 12 B/sample

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        nch = 8
        v, dt = cpu_serial_baseline(nch)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"{nch} of the 40 channels x 1M samples, one process, serial channel loop as the reference runs "
                         f"(scipy.signal.spectrogram + np.linalg.svd, BLAS threads as found; {dt:.1f} s)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "channels": N_CH, "samples_per_channel": N_SAMP, "shots_per_step_per_gpu": 1,
                   "sharding": "by shot, no data-path collective" if world > 1 else "single GPU",
                   "l2": f"inputs larger than L2: {nbuf} resident shots rotated (160 MB in + 320 MB out per step)",
                   "image_row_pitch_floats": ldt, "buffers": "allocated by the public API (Runtime.empty_image)",
                   "in_stream_fallback": fallback,
                   "shots_in_flight": f"{inflight} (api.ShotStreams: shot i on worker stream i % {inflight}, one context each; "
                                      "all K steps start after the first event and finish before the second)"},
        "ms_per_step_one_shot_at_a_time": ms_one / args.steps,
        "value_dense_layout": value_dense, "ms_per_step_dense_layout": ms_dense / args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_CH * N_SAMP * 4,
                "d2h_bytes_per_step": N_CH * ROWS * NSEG * 4, "steps": e2e_steps, "api": "HostPipeline(groups=8, streams=3).submit (one upload stream, one download stream, 3 staging sets with their own compute stream), shots submitted back to back",
                "gpu_launches": int(e2e_launches), "matches_device_path": e2e_ok,
                "copy_only_ceiling": e2e_ceiling,
                "frac_of_copy_ceiling": (e2e_value / e2e_ceiling) if e2e_value and e2e_ceiling else None},
        "gpu_launches": int(lt.item()),
        "clocks": clk.summary(),
        "roofline": roof,
        "pipeline_hbm": {"algorithmic_gbs": pipeline_gbs, "frac_of_peak": pipeline_gbs / peak, "bytes_per_sample": 12},
        "kernels": kernels,
        "config4": cfg4,
        "csd5": cfg5,
        "use_optimal": useopt,
        "cpu_baseline": cpu,
        "parity_on_bench_bytes": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling aid: skip the host-to-host leg (e2e is then null)")
    ap.add_argument("--no-config4", action="store_true", help="profiling aid: skip the 1000-shot streamed leg")
    ap.add_argument("--no-csd5", action="store_true", help="profiling aid: skip the all-pairs CSD leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
