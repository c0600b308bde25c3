"""probe: fused reducible statistics (iqw_stft_reduce_c64 through persistence_spectrum) at config-3 size"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iqwaveform_b200 as iqw
import bench
dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000000
x = bench.device_capture(torch, n, 1234, dev).view(1, n)
for nfft in [int(a) for a in sys.argv[2:]] or (4096, 2048, 1024):
    for stats, dB in ((['mean', 'max'], True), (['max'], True), (['mean', 'max', 'min'], True), (['mean', 'max'], False)):
        f = lambda: iqw.persistence_spectrum(x, fs=100e6, window='hann', resolution=100e6 / nfft, fractional_overlap=0.5,
                                             statistics=stats, dB=dB, axis=1)
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            for _ in range(5): f()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 5)
        print(nfft, stats, 'dB' if dB else 'lin', '%.3f ms  %.1f GS/s' % (best, n / best / 1e6))
