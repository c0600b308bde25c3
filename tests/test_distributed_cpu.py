"""N > 1 host logic on CPU: shard planners and the gloo all_gather plumbing of
iqwaveform_b200.distributed, with the numpy oracle standing in for the per-rank CUDA compute
(world_size 2 and 3, 127.0.0.1 rendezvous).  Sharded == unsharded, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iqwaveform_b200 import distributed as D
from oracle import iqw_oracle as orc
from oracle.make_golden import synth
from _shard_ops_numpy import NumpyShardOps, float_to_key


def test_shard_planners_cover_everything_once():
    for n, world in [(8, 8), (8, 3), (5, 8), (1000, 7)]:
        got = [c for r in range(world) for c in D.channel_shard(n, world, r)]
        assert got == list(range(n))
    for N, nfft, nov, world in [(100000, 256, 128, 4), (100001, 256, 192, 3), (4096 * 9 + 5, 4096, 0, 2),
                                (1000, 64, 48, 8), (300, 256, 128, 4)]:
        hop = nfft - nov
        T = orc.frame_count(N, nfft, nov)
        frames = []
        for r in range(world):
            s = D.frame_shard(N, nfft, nov, world, r)
            assert s.n_frames == T
            frames += list(range(s.frame0, s.frame1))
            if s.frame1 > s.frame0:
                assert s.sample0 == s.frame0 * hop and s.sample1 == (s.frame1 - 1) * hop + nfft <= N
        assert frames == list(range(T))
    for N, nb, world in [(1000, 10, 3), (1005, 10, 4), (7, 10, 2)]:
        bins = []
        for r in range(world):
            s = D.bin_shard(N, nb, world, r)
            bins += list(range(s.bin0, s.bin1))
            assert (s.sample0, s.sample1) == (s.bin0 * nb, s.bin1 * nb)
        assert bins == list(range(N // nb))
    with pytest.raises(ValueError):
        D.frame_shard(100, 16, 16, 2, 0)
    with pytest.raises(ValueError):
        D.channel_shard(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # --- channel-sharded persistence spectrum (config 3 layout) ---
        C, N = 5, 40000
        x = synth(3, (C, N))
        kw = dict(fs=1e6, window='hann', resolution=1e6 / 256, fractional_overlap=0.5,
                  statistics=[0.1, 0.5, 'mean', 'max'], dB=True)
        mine = D.channel_shard(C, world, rank)
        full = D.persistence_spectrum_sharded(torch.from_numpy(x[mine.start:mine.stop]), n_channels=C,
                                              compute=lambda a, **k: orc.persistence_spectrum(a.numpy(), **k), **kw)
        want = orc.persistence_spectrum(x, axis=1, **kw)
        assert np.array_equal(full.numpy(), want)

        # --- time-sharded spectrogram with halo (config 2 layout) ---
        N, nfft, nov = 50001, 256, 192
        x1 = synth(4, (N,))
        sh = D.frame_shard(N, nfft, nov, world, rank)
        spg = D.spectrogram_time_sharded(
            torch.from_numpy(x1[sh.sample0:sh.sample1]), n_samples=N, nperseg=nfft, noverlap=nov, gather=True,
            fs=1e6, window='hann', compute=lambda a, **k: orc.spectrogram(a.numpy(), **k))
        want = orc.spectrogram(x1, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=0,
                               return_axis_arrays=False)
        assert np.array_equal(spg.numpy(), want)

        # --- bin-sharded iq_to_bin_power (config 4 layout), ragged bin counts ---
        N, Ts, Tbin = 10007, 1e-6, 1e-4
        x2 = synth(5, (N,))
        bs = D.bin_shard(N, 100, world, rank)
        for kind in ('mean', 'max'):
            pw = D.iq_to_bin_power_sharded(torch.from_numpy(x2[bs.sample0:bs.sample1]), Ts, Tbin, n_samples=N,
                                           kind=kind, compute=lambda a, *p, **k: orc.iq_to_bin_power(a.numpy(), *p, **k))
            want = orc.iq_to_bin_power(x2, Ts, Tbin, kind=kind, truncate=True)
            assert np.array_equal(pw.numpy(), want)
        # --- time-sharded persistence spectrum: mean/max/min by one all_reduce each, quantiles and
        #     median EXACT by the distributed radix select (4 all_reduce of digit counts) ---
        for N, nfft, ov, bw in [(70001, 256, 0.5, float('inf')), (300, 256, 0.75, 0.5e6), (1000, 256, 0.75, 0.5e6)]:
            x3 = synth(6, (N,))
            pk = dict(fs=1e6, window='hann', resolution=1e6 / nfft, fractional_overlap=ov, dB=True, bandwidth=bw)
            stats = ['mean', 0.1, 'max', 'min', 0.5, 'median', 'peak', 0.999, 1.0, 0.0]
            fsh = D.frame_shard(N, nfft, round(ov * nfft), world, rank)
            ops = NumpyShardOps()
            got = D.persistence_spectrum_time_sharded(
                torch.from_numpy(x3[fsh.sample0:fsh.sample1]), n_samples=N, statistics=stats, ops=ops, **pk).numpy()
            assert ops.collected and (N < 70001 or not ops.overflowed)     # stationary capture: the bracketed path
            want = orc.persistence_spectrum(x3, statistics=stats, axis=0, **pk)
            assert got.shape == want.shape
            assert np.array_equal(got[1:], want[1:])                      # order statistics, max, min: exact
            np.testing.assert_allclose(got[0], want[0], atol=1e-3)        # mean of dB: fp32 summation order
        # digital silence in most of the capture: heavy ties overflow the candidate store and the
        # select falls back to counting over the whole local spectrogram; same exact result
        N = 40000
        x4 = synth(8, (N,))
        x4[3000:33000] = 0
        pk = dict(fs=1e6, window='hann', resolution=1e6 / 128, fractional_overlap=0.5, dB=True)
        fsh = D.frame_shard(N, 128, 64, world, rank)
        p4 = NumpyShardOps().power_spectrogram(x4[fsh.sample0:fsh.sample1], window='hann', nfft=128, noverlap=64,
                                               nzero=0, bin_lo=0, bin_hi=128)
        for sel, stored in [([fsh.n_frames // 2, fsh.n_frames // 2 + 1], False), ([fsh.n_frames - 3], None)]:
            info = {}
            keys = D.select_order_statistics(p4, sel, fsh.n_frames, ops=NumpyShardOps(), info=info)
            assert stored is None or info['candidate_store'] is stored
            full = orc.spectrogram(x4, fs=1e6, window='hann', nperseg=128, noverlap=64, axis=0, return_axis_arrays=False)
            assert np.array_equal(keys.numpy().view(np.uint32), np.sort(float_to_key(full), axis=0)[sel])
        with pytest.raises(ValueError):
            D.persistence_spectrum_time_sharded(torch.from_numpy(x3[:5]), n_samples=N, statistics=[0.5],
                                                ops=NumpyShardOps(), **pk)
        open(os.path.join(out_dir, f'ok{rank}'), 'w').close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_equals_unsharded_gloo(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))
