"""world-size-2 tests of the sharded paths on CPU (gloo): shot sharding has no collective; the channel-block
CSD has one all-gather.  The kernels run through the CPU-emulation build (tests/emu); on the GPU box the same
host code runs over NCCL (bench.py --gpus N, tests/test_gpu_parity.py for the single-GPU legs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_path, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import spec_oracle as oc
        from spectrogram_enhancement_b200 import _ffi, api, parallel
        rt = api.Runtime(_ffi.Library(emu_path), "cpu")
        C, n = 4, 3000
        x = np.stack([oc.synth_ece(2, c, n=n, fs=1.6e6) for c in range(C)])
        lo, hi = parallel.channel_block(rank, world, C)
        f, P = parallel.csd_allpairs_sharded(x[lo:hi], fs=1.6e6, nperseg=128, noverlap=64, runtime=rt)
        np.save(os.path.join(outdir, f"P{rank}.npy"), P)
        # frequency-block exchange: channel-sharded input, one all-to-all, output sharded by frequency (65 bins over 2 ranks)
        ff, Pf = parallel.csd_allpairs_freq_sharded(x[lo:hi], fs=1.6e6, nperseg=128, noverlap=64, runtime=rt)
        np.save(os.path.join(outdir, f"F{rank}.npy"), Pf)
        np.save(os.path.join(outdir, f"Ff{rank}.npy"), ff)
        # segment sharding: every rank sees the record, sums its own segments, one all-reduce; odd segment count and
        # a loader callable on the second call
        x9 = np.stack([oc.synth_ece(4, c, n=2100, fs=1.6e6) for c in range(9)])
        f, Pa = parallel.csd_allpairs_segment_sharded(x9, fs=1.6e6, nperseg=64, runtime=rt)
        f, Pb = parallel.csd_allpairs_segment_sharded(lambda lo, hi: x9[:, lo:hi], fs=1.6e6, nperseg=64, n_samples=2100,
                                                      n_channels=9, runtime=rt)
        assert np.array_equal(Pa, Pb)
        np.save(os.path.join(outdir, f"S{rank}.npy"), Pa)
        # shot sharding: every rank runs the pipeline on its own shots, no communication
        sp = dict(oc.DEFAULT_SPEC_PARAMS, nperseg=32, noverlap=16)
        got = {}
        for i, (S, D) in parallel.pipeline_sharded(lambda i: np.stack([oc.synth_ece(i, c, n=2000) for c in range(2)]), 5,
                                                   sp, runtime=rt):
            got[i] = D
        np.savez(os.path.join(outdir, f"D{rank}.npz"), **{str(k): v for k, v in got.items()})
    finally:
        dist.destroy_process_group()


def test_shot_range_partitions():
    from spectrogram_enhancement_b200 import parallel
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 1000):
            rs = [parallel.shot_range(r, world, n) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.channel_block(0, 3, 40)
    assert [parallel.segment_range(r, 3, 7) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    assert [parallel.frequency_block(r, 4, 513) for r in range(4)] == [(0, 129), (129, 258), (258, 387), (387, 513)]
    assert [parallel.frequency_block(r, 8, 5) for r in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 5), (5, 5), (5, 5)]


def test_world2_csd_and_shot_sharding(tmp_path, emu_rt):
    from oracle import spec_oracle as oc
    from emu.build_emu import EMU_LIB
    world = 2
    for attempt in range(3):          # the rendezvous port is picked optimistically: retry on a collision
        try:
            mp.spawn(_worker, args=(world, _free_port(), EMU_LIB, str(tmp_path)), nprocs=world, join=True)
            break
        except Exception:
            if attempt == 2:
                raise
    C, n = 4, 3000
    x = np.stack([oc.synth_ece(2, c, n=n, fs=1.6e6) for c in range(C)])
    _, Pr = oc.csd_allpairs(x.astype(np.float64), fs=1.6e6, nperseg=128, noverlap=64)
    P = np.concatenate([np.load(tmp_path / f"P{r}.npy") for r in range(world)])
    np.testing.assert_allclose(P, Pr, rtol=1e-4, atol=1e-6 * np.abs(Pr).max())
    # sharded (segment blocks exchanged and accumulated one by one) vs unsharded: same sums in another order
    from spectrogram_enhancement_b200 import api
    _, P1 = api.csd_allpairs(x, fs=1.6e6, nperseg=128, noverlap=64, runtime=emu_rt)
    np.testing.assert_allclose(P, P1, rtol=1e-5, atol=1e-7 * np.abs(P1).max())
    Pf = np.concatenate([np.load(tmp_path / f"F{r}.npy") for r in range(world)], axis=-1)      # frequency blocks side by side
    ff = np.concatenate([np.load(tmp_path / f"Ff{r}.npy") for r in range(world)])
    assert Pf.shape == Pr.shape and np.array_equal(ff, np.fft.rfftfreq(128, 1 / 1.6e6))
    np.testing.assert_allclose(Pf, Pr, rtol=1e-4, atol=1e-6 * np.abs(Pr).max())
    x9 = np.stack([oc.synth_ece(4, c, n=2100, fs=1.6e6) for c in range(9)])
    _, P9 = oc.csd_allpairs(x9.astype(np.float64), fs=1.6e6, nperseg=64, noverlap=32)
    S0, S1 = np.load(tmp_path / "S0.npy"), np.load(tmp_path / "S1.npy")
    assert np.array_equal(S0, S1)                       # the all-reduce leaves every rank with the same matrix
    np.testing.assert_allclose(S0, P9, rtol=1e-4, atol=1e-6 * np.abs(P9).max())
    d0, d1 = np.load(tmp_path / "D0.npz"), np.load(tmp_path / "D1.npz")
    assert sorted(d0.files) == ["0", "1", "2"] and sorted(d1.files) == ["3", "4"]
    sp = dict(oc.DEFAULT_SPEC_PARAMS, nperseg=32, noverlap=16)
    for i in range(5):
        xs = np.stack([oc.synth_ece(i, c, n=2000) for c in range(2)])
        _, D = api.pipeline(xs, sp, runtime=emu_rt)
        assert np.array_equal((d0 if i < 3 else d1)[str(i)], D)


# ---- the same sharded paths on real GPUs over NCCL (2 ranks; skipped with fewer than 2 devices) ---------------------
def _nccl_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import spec_oracle as oc
        from spectrogram_enhancement_b200 import api, parallel
        rt = api.Runtime(device=dev)
        C, n, nps = 8, 120_000, 1024
        x = np.stack([oc.synth_ece(6, c, n=n) for c in range(C)])
        lo, hi = parallel.channel_block(rank, world, C)
        xl = torch.from_numpy(x[lo:hi]).to(dev)
        _, P = parallel.csd_allpairs_sharded(xl, fs=500000.0, nperseg=nps, blocks=3, runtime=rt)
        _, Pf = parallel.csd_allpairs_freq_sharded(xl, fs=500000.0, nperseg=nps, runtime=rt)
        _, Ps = parallel.csd_allpairs_segment_sharded(torch.from_numpy(x).to(dev), fs=500000.0, nperseg=nps, runtime=rt)
        torch.cuda.synchronize()
        np.save(os.path.join(outdir, f"P{rank}.npy"), P.cpu().numpy())
        np.save(os.path.join(outdir, f"F{rank}.npy"), Pf.cpu().numpy())
        np.save(os.path.join(outdir, f"S{rank}.npy"), Ps.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_world2_csd_shardings(tmp_path):
    """BASELINE config 5's exchange step on two real GPUs: channel-block all-gather, frequency-block all-to-all and
    segment-sharded all-reduce all reproduce the oracle's all-pairs matrix."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    from oracle import spec_oracle as oc
    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    C, n, nps = 8, 120_000, 1024
    x = np.stack([oc.synth_ece(6, c, n=n) for c in range(C)])
    _, Pr = oc.csd_allpairs(x.astype(np.float64), fs=500000.0, nperseg=nps)
    tol = dict(rtol=1e-4, atol=1e-6 * np.abs(Pr).max())
    np.testing.assert_allclose(np.concatenate([np.load(tmp_path / f"P{r}.npy") for r in range(world)]), Pr, **tol)
    np.testing.assert_allclose(np.concatenate([np.load(tmp_path / f"F{r}.npy") for r in range(world)], axis=-1), Pr, **tol)
    S0, S1 = np.load(tmp_path / "S0.npy"), np.load(tmp_path / "S1.npy")
    assert np.array_equal(S0, S1)
    np.testing.assert_allclose(S0, Pr, **tol)
