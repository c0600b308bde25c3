#!/usr/bin/env python
"""cost of the per-launch profile events inside the headline step: python tools/profile_overhead_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib

n = 1_000_000_000
x = bench.device_capture(torch, n, 1234, torch.device('cuda', 0)).view(1, n)
kw = dict(fs=100e6, window='hann', resolution=100e6 / 4096, fractional_overlap=0.5, statistics=[0.1, 0.5, 0.9, 0.999], dB=True, axis=1)
for _ in range(3):
    iqw.persistence_spectrum(x, **kw)
for rep in range(2):
    for name, on, fine in (('off', False, False), ('level 1 (bench)', True, False), ('level 2 (every launch)', True, True)):
        _lib.profile(on, fine=fine)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            iqw.persistence_spectrum(x, **kw)
        e1.record(); torch.cuda.synchronize()
        print(f'profile events {name}: {e0.elapsed_time(e1) / 20:.4f} ms per step')
_lib.profile(False)
