"""probe: BASELINE configs[0] (persistence spectrum of 15.36 M samples, nfft 1024, q = [0.5, 0.99]) --
plain call against the captured CUDA graph, with the per-kernel times of the plain call"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib
dev = torch.device('cuda:0')
n = 15_360_000
x = bench.device_capture(torch, n, 1, dev).view(1, n)
kw = dict(fs=15.36e6, window='hann', resolution=15e3, fractional_overlap=0.5, statistics=[0.5, 0.99], dB=True, axis=1)
plain = bench._timed(torch, lambda: iqw.persistence_spectrum(x, **kw), reps=10)
_lib.profile(True, fine=True)
ref = iqw.persistence_spectrum(x, **kw).clone()
torch.cuda.synchronize()
rep = _lib.profile_report(); _lib.profile(False)
print('plain call %.4f ms; kernels: %s' % (plain, ', '.join(f'{k} {c}x {ms:.4f}' for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]))))
print('sum of kernel times %.4f ms over %d launches' % (sum(ms for c, ms in rep.values()), sum(c for c, ms in rep.values())))
g = iqw.GraphedCall(iqw.persistence_spectrum, x, **kw)
graphed = bench._timed(torch, lambda: g(), reps=10)
with_copy = bench._timed(torch, lambda: g(x), reps=10)
print('graph replay %.4f ms (%.1f GS/s, %.3f of the measured HBM peak at 24 B/sample); with the input copy %.4f ms; identical: %s'
      % (graphed, n / graphed / 1e6, 24 * n / graphed / 1e6 / bench.measured_peak()[0], with_copy, bool(torch.equal(g(), ref))))

# sample size: rows / share
for share in (4, 8, 32):
    _lib.lib.iqw_debug_set_sample_min_rows(-share)
    pl = bench._timed(torch, lambda: iqw.persistence_spectrum(x, **kw), reps=10)
    same = bool(torch.equal(iqw.persistence_spectrum(x, **kw), ref))
    gs = iqw.GraphedCall(iqw.persistence_spectrum, x, **kw)
    print('sample = rows / %d: plain %.4f ms, graph %.4f ms, identical %s' % (share, pl, bench._timed(torch, lambda: gs(), reps=10), same))
_lib.lib.iqw_debug_set_sample_min_rows(0)

# the exact multi-pass pipeline instead of the sampled one-read path (the 123 MB matrix fits the L2)
_lib.lib.iqw_debug_set_sample_min_rows(1 << 30)
ex = bench._timed(torch, lambda: iqw.persistence_spectrum(x, **kw), reps=10)
_lib.profile(True, fine=True)
out2 = iqw.persistence_spectrum(x, **kw).clone()
torch.cuda.synchronize()
rep = _lib.profile_report(); _lib.profile(False)
print('exact pipeline: plain call %.4f ms, identical %s; kernels: %s' % (ex, bool(torch.equal(out2, ref)), ', '.join(f'{k} {c}x {ms:.4f}' for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]))))
g2 = iqw.GraphedCall(iqw.persistence_spectrum, x, **kw)
print('exact pipeline as a graph: %.4f ms' % bench._timed(torch, lambda: g2(), reps=10))
_lib.lib.iqw_debug_set_sample_min_rows(0)
