#!/usr/bin/env python
"""one launch of each kernel of the rows next to the hot path, for an ncu capture:
ncu --set full --clock-control none -k regex:"istft_kernel|ola_kernel|edge_count_kernel|bracket_collect_kernel|candidate_count_kernel" \
    -c 8 -o gpurun_out/prof_next python tools/ncu_next_rows.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _plan
from iqwaveform_b200 import distributed as D

n = 1 << 27
x = torch.randn(n, dtype=torch.complex64, device='cuda')
for nfft in (1024, 4096):
    y = iqw.stft(x, fs=1e6, window='hamming', nperseg=nfft, noverlap=nfft // 2, truncate=False, return_axis_arrays=False)
    iqw.istft(y, n, nfft=nfft, noverlap=nfft // 2)
    del y
    iqw.ola_filter(x, fs=1e6, nfft=nfft, window='hamming', passband=(-2e5, 2e5))
p = iqw.spectrogram(x, fs=1e6, window='hann', nperseg=4096, noverlap=2048, return_axis_arrays=False)   # (T, 4096)
iqw.sample_ccdf(p.reshape(-1), np.linspace(0, 8, 257), density=False)
T = p.shape[0]
reqs = _plan.stat_requests([0.1, 0.5, 0.9, 0.999], T)
sel = sorted(_plan.distinct_ranks(reqs, T))
half = T // 2
tg = D.ThreadGroup(2)
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(2) as ex:
    list(ex.map(lambda r: D.select_order_statistics(p[:half] if r == 0 else p[half:], sel, T, group=tg.member(r)), range(2)))
torch.cuda.synchronize()
print('done')
