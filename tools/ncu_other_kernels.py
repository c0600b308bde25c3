#!/usr/bin/env python
"""one launch of the kernels the headline step does not use (other STFT sizes and modes, the large-nfft
pair, bin power, elementwise), for an ncu capture:
ncu --set full --clock-control none -k regex:"stft_kernel|columns_kernel|rows_kernel|bin_power|ew_real|ew_complex|envtopow" \
    -c 14 -o gpurun_out/prof_other python tools/ncu_other_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iqwaveform_b200 as iqw

n = 1 << 27
x = torch.randn(n, dtype=torch.complex64, device='cuda')
kw = dict(fs=1e8, window='hann', return_axis_arrays=False)
for nfft in (64, 256, 1024, 2048, 8192):
    iqw.spectrogram(x, nperseg=nfft, noverlap=nfft // 2, **kw)
iqw.spectrogram(x, nperseg=2048, noverlap=1024, dB=True, **kw)
iqw.stft(x[:n // 2], nperseg=2048, noverlap=1024, norm='power', **kw)
iqw.spectrogram(x, nperseg=65536, noverlap=32768, **kw)
iqw.iq_to_bin_power(x, 1 / 245.76e6, 1e-3, kind='mean', truncate=True)
iqw.iq_to_bin_power(x, 1 / 245.76e6, 1e-3, kind='peak', truncate=True)
p = torch.rand(n, device='cuda') + 1e-3
iqw.powtodB(p)
iqw.dBtopow(p)
iqw.envtopow(x)
torch.cuda.synchronize()
print('done')
