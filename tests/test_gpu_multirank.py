"""N > 1 on real GPUs (skipped with fewer than 2 devices): one process per GPU over NCCL; the
channel-, frame- and bin-sharded results equal the single-GPU ones bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import iqwaveform_b200 as iqw
    from iqwaveform_b200 import distributed as D
    from oracle.make_golden import synth

    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        C, N = 2 * world + 1, 1 << 18
        x = torch.from_numpy(synth(21, (C, N))).to(dev)
        kw = dict(fs=1e6, window='hann', resolution=1e6 / 1024, fractional_overlap=0.5,
                  statistics=[0.1, 0.5, 0.999, 'max', 'mean'], dB=True)
        mine = D.channel_shard(C, world, rank)
        full = D.persistence_spectrum_sharded(x[mine.start:mine.stop], n_channels=C, **kw)
        single = iqw.persistence_spectrum(x, axis=1, **kw)
        assert torch.equal(full, single)

        x1 = x[0]
        sh = D.frame_shard(N, 2048, 1024, world, rank)
        spg = D.spectrogram_time_sharded(x1[sh.sample0:sh.sample1].contiguous(), n_samples=N, nperseg=2048,
                                         noverlap=1024, gather=True, fs=1e6, window='blackmanharris')
        assert torch.equal(spg, iqw.spectrogram(x1, fs=1e6, window='blackmanharris', nperseg=2048, noverlap=1024,
                                                return_axis_arrays=False))
        bs = D.bin_shard(N, 1000, world, rank)
        pw = D.iq_to_bin_power_sharded(x1[bs.sample0:bs.sample1].contiguous(), 1e-6, 1e-3, n_samples=N, kind='peak')
        assert torch.equal(pw, iqw.iq_to_bin_power(x1, 1e-6, 1e-3, kind='peak', truncate=True))
        # one capture split in time: exact quantiles through 4 NCCL all_reduce of digit counts
        kw = dict(fs=1e6, window='hann', resolution=1e6 / 1024, fractional_overlap=0.5, dB=True,
                  statistics=['mean', 0.1, 0.5, 'median', 0.999, 'max', 'min'])
        sh = D.frame_shard(N, 1024, 512, world, rank)
        got = D.persistence_spectrum_time_sharded(x1[sh.sample0:sh.sample1].contiguous(), n_samples=N, **kw)
        want = iqw.persistence_spectrum(x1, axis=0, **kw)
        assert torch.equal(got[1:], want[1:])
        assert torch.allclose(got[0], want[0], atol=1e-3)
        open(os.path.join(out_dir, f'ok{rank}'), 'w').close()
    finally:
        dist.destroy_process_group()


def test_sharded_equals_single_gpu_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least 2 GPUs')
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


def test_tensor_on_non_current_device(cuda_device):
    """every C entry point launches on the device that owns its buffers: a capture on cuda:1 while
    cuda:0 is the current device gives the same bits as the same capture on cuda:0"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least 2 GPUs')
    import iqwaveform_b200 as iqw
    from oracle.make_golden import synth
    x = synth(3, (2, 1 << 17))
    kw = dict(fs=1e6, window='hann', resolution=1e6 / 1024, fractional_overlap=0.5,
              statistics=[0.1, 0.5, 'max', 'mean'], dB=True, axis=1)
    torch.cuda.set_device(0)
    a = iqw.persistence_spectrum(torch.from_numpy(x).to('cuda:0'), **kw)
    b = iqw.persistence_spectrum(torch.from_numpy(x).to('cuda:1'), **kw)       # cuda:0 still current
    assert torch.cuda.current_device() == 0 and b.device.index == 1
    assert torch.equal(a.cpu(), b.cpu())
    pa = iqw.iq_to_bin_power(torch.from_numpy(x).to('cuda:0'), 1e-6, 1e-3, kind='peak', axis=1, truncate=True)
    pb = iqw.iq_to_bin_power(torch.from_numpy(x).to('cuda:1'), 1e-6, 1e-3, kind='peak', axis=1, truncate=True)
    assert torch.equal(pa.cpu(), pb.cpu())
