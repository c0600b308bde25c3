#!/usr/bin/env python
"""experiment: timeline of the STFT (stream 0) and statistics (stream 1) stages of consecutive channels.
Result (round 1): the stages do run side by side, but they share the bottleneck -- next to an STFT the
statistics chain takes 8.2 ms instead of 3.65 and the STFT 5.0-6.5 ms instead of 4.6-4.9 -- so the
channel rate does not improve ("measured and rejected")."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from iqwaveform_b200 import _lib, fourier

C = 4
n = 1_000_000_000
dev = torch.device('cuda', 0)
x = torch.stack([bench.device_capture(torch, n, 100 + c, dev) for c in range(C)])
stats = [0.1, 0.5, 0.9, 0.999]
T = (n - 4096) // 2048 + 1
spg = [torch.empty((1, T, 4096), dtype=torch.float32, device=dev) for _ in range(2)]
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def run(log):
    out = torch.empty((C, 4, 4096), dtype=torch.float32, device=dev)
    main = torch.cuda.current_stream()
    t0 = torch.cuda.Event(enable_timing=True); t0.record(main)
    sa.wait_stream(main); sb.wait_stream(main)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    a0, a1, b0, b1 = [ev() for _ in range(C)], [ev() for _ in range(C)], [ev() for _ in range(C)], [ev() for _ in range(C)]
    for c in range(C):
        with torch.cuda.stream(sa):
            if c >= 2:
                sa.wait_event(b1[c - 2])
            a0[c].record(sa)
            p = fourier._stft_device(x[c:c + 1], window='hann', nfft=4096, noverlap=2048, nzero=0, norm='power',
                                     truncate=True, mode=_lib.STFT_POWER, out=spg[c % 2])
            a1[c].record(sa)
        with torch.cuda.stream(sb):
            sb.wait_event(a1[c])
            b0[c].record(sb)
            fourier.time_statistics(p, stats, dB=True, eps=1e-25, out=out[c:c + 1])
            b1[c].record(sb)
    main.wait_stream(sa); main.wait_stream(sb)
    torch.cuda.synchronize()
    if log:
        for c in range(C):
            print(f'ch {c}: stft {t0.elapsed_time(a0[c]):7.2f} .. {t0.elapsed_time(a1[c]):7.2f}   '
                  f'stats {t0.elapsed_time(b0[c]):7.2f} .. {t0.elapsed_time(b1[c]):7.2f}')


run(False); run(False); run(True)
