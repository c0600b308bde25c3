"""probe: spectrogram throughput sweep (nfft x overlap), device-resident input, CUDA events"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib
dev = torch.device('cuda:0')
n = 1 << 28
x = torch.randn(n, dtype=torch.complex64, device=dev)
nffts = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [64, 256, 1024, 2048, 4096, 8192]
ovs = [float(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0.5, 0.75]
modes = sys.argv[3].split(',') if len(sys.argv) > 3 else ['power']
if len(sys.argv) > 4:
    _lib.lib.iqw_debug_set_stft_scratch_cap(int(float(sys.argv[4]) * (1 << 20)))
for nfft in nffts:
    for ov in ovs:
        for mode in modes:
            nov = int(nfft * ov)
            def run():
                if mode == 'complex':
                    return iqw.stft(x, fs=1e8, window='hann', nperseg=nfft, noverlap=nov, norm='power', return_axis_arrays=False)
                return iqw.spectrogram(x, fs=1e8, window='hann', nperseg=nfft, noverlap=nov, return_axis_arrays=False, dB=(mode == 'dB'))
            out = run(); del out
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            _lib.profile(True)
            for _ in range(3):
                e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1)); del out
            rep = _lib.profile_report(); _lib.profile(False)
            detail = ' '.join(f'{k}={ms / c:.3f}' for k, (c, ms) in rep.items()) if len(rep) > 1 else ''
            r = nfft / (nfft - nov)
            bps = 8 + (8 if mode == 'complex' else 4) * r
            print(f'{mode:8s} nfft={nfft:6d} ov={ov:4.2f}: {best:7.3f} ms {n / best / 1e6:7.1f} GS/s {n * bps / best / 1e6:7.0f} GB/s algorithmic ({n * bps / best / 1e6 / 6538.9 * 100:4.1f} % of measured peak) {detail}')
