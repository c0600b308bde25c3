#!/usr/bin/env python
"""time kernel 5 (edge counts): python tools/hist_probe.py [n] [n_edges,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from iqwaveform_b200.power_analysis import sample_ccdf
from iqwaveform_b200.util import histogram_last_axis

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
ne = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [16, 128, 1024]
PEAK, _ = bench.measured_peak()
p = torch.randn(n, device='cuda').square_()
for k in ne:
    edges = np.linspace(0, 8, k)
    for name, fn in (('sample_ccdf', lambda: sample_ccdf(p, edges, density=False)),
                     ('histogram_last_axis (1024 rows)', lambda: histogram_last_axis(p.view(1024, -1), edges))):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f'{name} {k} edges: {best:.3f} ms  {n * 4 / best / 1e6:.0f} GB/s ({n * 4 / best / 1e6 / PEAK:.2f} of the measured HBM peak)')
