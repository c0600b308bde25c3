#!/usr/bin/env python
"""experiment: persistence spectrum of several channels with the STFT -> statistics chain of
consecutive channels on two alternating streams (does channel c+1's STFT overlap channel c's
statistics?).  python tools/exp/two_stream_channels.py [channels]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib, fourier

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = 1_000_000_000
dev = torch.device('cuda', 0)
x = torch.stack([bench.device_capture(torch, n, 100 + c, dev) for c in range(C)])
kw = dict(fs=100e6, window='hann', resolution=100e6 / 4096, fractional_overlap=0.5, statistics=[0.1, 0.5, 0.9, 0.999], dB=True, axis=1)
stats = kw['statistics']


def sequential():
    return iqw.persistence_spectrum(x, **kw)


streams = [torch.cuda.Stream(), torch.cuda.Stream()]
T = (n - 4096) // 2048 + 1
spg = [torch.empty((1, T, 4096), dtype=torch.float32, device=dev) for _ in range(2)]


def two_streams():
    out = torch.empty((C, 4, 4096), dtype=torch.float32, device=dev)
    main = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(main)
    for c in range(C):
        s = streams[c % 2]
        with torch.cuda.stream(s):
            p = fourier._stft_device(x[c:c + 1], window='hann', nfft=4096, noverlap=2048, nzero=0, norm='power',
                                     truncate=True, mode=_lib.STFT_POWER, out=spg[c % 2])
            fourier.time_statistics(p, stats, dB=True, eps=1e-25, out=out[c:c + 1])
    for s in streams:
        main.wait_stream(s)
    return out


def one_by_one():
    return torch.stack([iqw.persistence_spectrum(x[c], **dict(kw, axis=0)) for c in range(C)])


for name, fn in (('library, (C, N) input', sequential), ('two streams by hand', two_streams), ('channel by channel', one_by_one), ('library, (C, N) input', sequential)):
    ref = fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        o = fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1) / 5 / C:.3f} ms per channel')
print('equal:', torch.equal(sequential(), one_by_one()))
