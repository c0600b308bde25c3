#!/usr/bin/env python
"""per-kernel times of BASELINE configs[0] (15.36 MS/s x 1 s, nfft 1024, q = [0.5, 0.99]): a small,
launch-bound problem.  python tools/config0_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib

n = 15_360_000
x = bench.device_capture(torch, n, 1, torch.device('cuda', 0)).view(1, n)
kw = dict(fs=15.36e6, window='hann', resolution=15e3, fractional_overlap=0.5, statistics=[0.5, 0.99], dB=True, axis=1)
for it in range(3):
    _lib.profile(it == 2, fine=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = iqw.persistence_spectrum(x, **kw); e1.record(); torch.cuda.synchronize()
    print(f'call {it}: {e0.elapsed_time(e1):.3f} ms')
rep = _lib.profile_report()
tot = sum(ms for _, ms in rep.values())
for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f'  {k:24s} {c:3d} launches {ms * 1e3:8.1f} us')
print(f'  sum of kernels {tot * 1e3:.1f} us')
