"""probe: path counters and per-kernel times of time_statistics on bench-like data"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib
import bench
dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
x = bench.device_capture(torch, n, 1234, dev).view(1, n)
_, _, p = iqw.spectrogram(x, fs=100e6, window='hann', nperseg=4096, noverlap=2048, axis=1)
print('spectrogram', tuple(p.shape))
ALL = ([0.1, 0.5, 0.9, 0.999], [0.5], [0.5, 0.99, 'mean', 'max'])
sel = [ALL[int(a)] for a in sys.argv[2:] if not a.startswith('m=')] or ALL
for a in sys.argv[2:]:
    if a.startswith('m='):
        _lib.lib.iqw_debug_set_sample_margin(float(a[2:]), 2)
for stats in sel:
    cnt = []
    out = iqw.time_statistics(p, stats, dB=True, counters=cnt)
    torch.cuda.synchronize()
    _lib.profile(True, fine=True)
    for _ in range(3):
        iqw.time_statistics(p, stats, dB=True)
    torch.cuda.synchronize()
    rep = _lib.profile_report(); _lib.profile(False)
    print(stats, cnt, 'nan:', int(torch.isnan(out).sum()))
    for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]:
        print('   %-22s %3d  %.3f ms' % (k, c, ms / c))
