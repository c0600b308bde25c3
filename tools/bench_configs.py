#!/usr/bin/env python
"""GPU timings of the BASELINE.json configs other than the headline one (bench.py measures
configs[2]); device-resident synthetic input, CUDA events, best of 3 after a warm-up.

    python tools/bench_configs.py [--out profiles/configs_<tag>.json] [--quick]

configs[0]  persistence_spectrum, 15.36 MS/s x 1 s, nfft 1024 Hann 50 %, q = [0.5, 0.99]
configs[1]  stft / spectrogram dB, 100 MS/s x 10 s, nfft 2048 Blackman-Harris 50 %
configs[3]  iq_to_bin_power mean / peak, 1 ms bins at 245.76 MS/s (one 8 s slice: 2e9 samples, 16 GB)
configs[4]  spectrogram sweep nfft in {64, 256, 1024, 8192, 65536} x overlap {50 %, 75 %}, 2^28 samples
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import iqwaveform_b200 as iqw  # noqa: E402

PEAK, _ = bench.measured_peak()


def timed(fn, reps=3):
    out = fn(); del out
    torch.cuda.synchronize()
    best = float('inf')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)); del out
    return best


def row(name, n, ms, bytes_per_sample):
    gbs = n * bytes_per_sample / ms / 1e6
    r = {'config': name, 'samples': n, 'ms': round(ms, 4), 'GS_per_s': round(n / ms / 1e6, 2),
         'algorithmic_bytes_per_sample': bytes_per_sample, 'algorithmic_GBps': round(gbs, 1),
         'frac_of_measured_hbm_peak': round(gbs / PEAK, 4)}
    print(json.dumps(r), flush=True)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--quick', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    rows = []

    # configs[0]
    n = 15_360_000
    x = bench.device_capture(torch, n, 1, dev).view(1, n)
    ms = timed(lambda: iqw.persistence_spectrum(x, fs=15.36e6, window='hann', resolution=15e3, fractional_overlap=0.5,
                                                statistics=[0.5, 0.99], dB=True, axis=1))
    rows.append(row('configs[0] persistence_spectrum 15.36 MS/s x 1 s nfft 1024 hann 50% q=[0.5,0.99]', n, ms, 24))
    del x

    # configs[2] with several channels on ONE GPU (the bench line is one channel per GPU): the chains of
    # consecutive channels run on two alternating streams
    if not a.quick:
        n = 1_000_000_000
        xs = torch.stack([bench.device_capture(torch, n, 10 + c, dev) for c in range(4)])
        ms = timed(lambda: iqw.persistence_spectrum(xs, fs=100e6, window='hann', resolution=100e6 / 4096,
                                                    fractional_overlap=0.5, statistics=[0.1, 0.5, 0.9, 0.999], dB=True, axis=1))
        rows.append(row('configs[2] persistence_spectrum, 4 channels x 1e9 samples on one GPU (two-stream channel pipeline)',
                        4 * n, ms, 24))
        del xs

    # configs[1]
    n = 100_000_000 if a.quick else 1_000_000_000
    x = bench.device_capture(torch, n, 2, dev)
    kw = dict(fs=100e6, window='blackmanharris', nperseg=2048, noverlap=1024, return_axis_arrays=False)
    ms = timed(lambda: iqw.spectrogram(x, dB=True, **kw))
    rows.append(row(f'configs[1] spectrogram dB 100 MS/s x {n / 100e6:g} s nfft 2048 blackmanharris 50%', n, ms, 16))
    ms = timed(lambda: iqw.spectrogram(x, **kw))
    rows.append(row(f'configs[1] spectrogram power, same capture', n, ms, 16))
    if a.quick or True:
        m = n // 2      # complex output is 16 B/sample at 50 %: half the capture keeps it within memory
        ms = timed(lambda: iqw.stft(x[:m], norm='power', **kw))
        rows.append(row(f'configs[1] stft complex, first {m} samples', m, ms, 24))
    del x

    # configs[3]: bins are independent; one 8 s slice of the 60 s capture per GPU pass
    n = 245_760 * (400 if a.quick else 8138)
    x = torch.empty(n, dtype=torch.complex64, device=dev)
    torch.view_as_real(x).normal_(0, 0.7)
    for kind in ('mean', 'peak'):
        ms = timed(lambda: iqw.iq_to_bin_power(x, 1 / 245.76e6, 1e-3, kind=kind))
        rows.append(row(f'configs[3] iq_to_bin_power {kind} 1 ms bins 245.76 MS/s x {n / 245.76e6:.1f} s', n, ms, 8))
    del x

    # configs[4]
    n = 1 << (26 if a.quick else 28)
    x = torch.randn(n, dtype=torch.complex64, device=dev)
    for nfft in (64, 256, 1024, 8192, 65536):
        for ov in (0.5, 0.75):
            nov = int(nfft * ov)
            ms = timed(lambda: iqw.spectrogram(x, fs=1e8, window='hann', nperseg=nfft, noverlap=nov,
                                               return_axis_arrays=False))
            rows.append(row(f'configs[4] spectrogram nfft {nfft} overlap {ov:.2f}', n, ms, 8 + 4 * nfft / (nfft - nov)))
    # inverse STFT and the overlap-add filter (hamming, 50 % overlap), same capture
    for nfft in (256, 1024, 4096):
        nov = nfft // 2
        y = iqw.stft(x, fs=1e8, window='hamming', nperseg=nfft, noverlap=nov, truncate=False, return_axis_arrays=False)
        ms = timed(lambda: iqw.istft(y, n, nfft=nfft, noverlap=nov))
        rows.append(row(f'(f3) istft nfft {nfft} overlap 0.50', n, ms, 24))
        del y
        ms = timed(lambda: iqw.ola_filter(x, fs=1e8, nfft=nfft, window='hamming', passband=(-2e7, 2e7)))
        rows.append(row(f'(f3) ola_filter nfft {nfft} hamming, one kernel', n, ms, 16))
        ms = timed(lambda: iqw.ola_filter(x, fs=1e8, nfft=nfft, window='hamming', passband=(-2e7, 2e7), fused=False))
        rows.append(row(f'(f3) ola_filter nfft {nfft} hamming, stft + istft kernels', n, ms, 48))
    # row A4's public pair: batched fft / ifft of (rows, n) complex64, 8 B read + 8 B written per sample
    for nfft in (256, 1024, 4096, 65536):
        xr = x.view(-1, nfft)
        ms = timed(lambda: iqw.fft(xr, axis=1))
        rows.append(row(f'(A4) fft rows of {nfft}', n, ms, 16))
        if nfft <= 8192:
            ms = timed(lambda: iqw.ifft(xr, axis=1))
            rows.append(row(f'(A4) ifft rows of {nfft}', n, ms, 16))
    if a.out:
        json.dump({'hbm_peak_GBps': PEAK, 'rows': rows}, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
