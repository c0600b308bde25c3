#!/usr/bin/env python
"""time kernel 4 (istft) and the stft -> istft / ola_filter chains on one GPU:
python tools/istft_probe.py [n_samples] [nffts] [overlap divisors R]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
nffts = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [256, 1024, 4096]
Rs = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [2, 4]
PEAK, _ = bench.measured_peak()
x = torch.randn(n, dtype=torch.complex64, device='cuda')


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)); del out
    return best


for nfft in nffts:
    for R in Rs:
        nov = nfft - nfft // R
        y = iqw.stft(x, fs=1e6, window='hamming', nperseg=nfft, noverlap=nov, truncate=False, return_axis_arrays=False)
        ms = timed(lambda: iqw.istft(y, n, nfft=nfft, noverlap=nov))
        by = 8 * R + 8
        print(f'istft nfft {nfft} R {R}: {ms:.3f} ms  {n / ms / 1e6:.1f} GS/s  {n * by / ms / 1e6:.0f} GB/s '
              f'({n * by / ms / 1e6 / PEAK:.2f} of the measured HBM peak, {by} B/sample)')
        del y
    ms = timed(lambda: iqw.ola_filter(x, fs=1e6, nfft=nfft, window='hamming', passband=(-2e5, 2e5), fused=False))
    print(f'ola_filter nfft {nfft} hamming, stft + istft kernels: {ms:.3f} ms  {n / ms / 1e6:.1f} GS/s  '
          f'({n * 48 / ms / 1e6 / PEAK:.2f} of the HBM peak at 48 B/sample: 8 in + 16 stft out + 16 in + 8 out)')
    ms = timed(lambda: iqw.ola_filter(x, fs=1e6, nfft=nfft, window='hamming', passband=(-2e5, 2e5)))
    print(f'ola_filter nfft {nfft} hamming, one kernel: {ms:.3f} ms  {n / ms / 1e6:.1f} GS/s  '
          f'({n * 16 / ms / 1e6 / PEAK:.2f} of the HBM peak at 16 B/sample: 8 in + 8 out)')

for up, down in ((512, 1024), (2048, 1024), (1024, 4096)):
    ms = timed(lambda: iqw.oaresample(x, up, down, 1e6, axis=0))
    print(f'oaresample down={down} up={up}: {ms:.3f} ms  {n / ms / 1e6:.1f} GS/s in')
ms = timed(lambda: iqw.oaresample(x, 512, 1024, 1e6, axis=0, filter_bandwidth=0.3e6, transition_bandwidth=50e3))
print(f'oaresample down=1024 up=512 with the FIR low-pass: {ms:.3f} ms  {n / ms / 1e6:.1f} GS/s in')
