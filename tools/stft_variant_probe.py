"""probe: three-pass (variant 1) against two-pass (variant 2) kernel 1 at nfft 1024 / 2048 / 4096:
spectrogram throughput (device-resident input, CUDA events, best of 5) and the largest difference
between the two outputs in units of the parity tolerance of tests/_tol.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib
dev = torch.device('cuda:0')
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
modes = sys.argv[2].split(',') if len(sys.argv) > 2 else ['power']
VARIANTS = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [1, 2, 3]
NFFTS = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else [1024, 2048, 4096]
x = torch.randn(n, dtype=torch.complex64, device=dev)
x += 3.0 * torch.exp(2j * torch.pi * 0.3256 * torch.arange(n, device=dev, dtype=torch.float64)).to(torch.complex64)
for nfft in NFFTS:
    for ov in (0.5, 0.75, 0.0):
        for mode in modes:
            nov = int(nfft * ov)
            def run():
                if mode == 'complex':
                    return iqw.stft(x[:n // 2], fs=1e8, window='hann', nperseg=nfft, noverlap=nov, norm='power', return_axis_arrays=False)
                return iqw.spectrogram(x, fs=1e8, window='hann', nperseg=nfft, noverlap=nov, return_axis_arrays=False, dB=(mode == 'dB'))
            res = {}
            outs = {}
            if nfft > 4096 and len(VARIANTS) > 2:
                pass
            for variant in VARIANTS:
                _lib.check(_lib.lib.iqw_debug_set_stft_variant(variant))
                out = run(); del out
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best = 1e9
                for _ in range(5):
                    e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                    if _ < 4: del out
                res[variant] = best
                outs[variant] = out[:20000].clone() if mode != 'complex' else torch.view_as_real(out[:20000]).clone()
                del out
            _lib.check(_lib.lib.iqw_debug_set_stft_variant(0))
            a, b = outs[VARIANTS[0]].double(), outs[VARIANTS[-1]].double()
            if mode == 'dB':
                err = (a - b).abs().max().item()
                unit = 'dB'
            else:
                mx = a.abs().amax(dim=-1 if mode != 'complex' else (-2, -1), keepdim=True)
                err = ((a - b).abs() / (1e-5 * a.abs() + 2e-7 * mx)).max().item()
                unit = 'tol units'
            m = n if mode != 'complex' else n // 2
            r = nfft / (nfft - nov)
            bps = 8 + (8 if mode == 'complex' else 4) * r
            f = lambda ms: f'{ms:7.3f} ms {m * bps / ms / 1e6 / 6538.9 * 100:5.1f} %'
            print(f'{mode:8s} nfft={nfft:5d} ov={ov:4.2f}: ' + '  '.join(f'v{k} {f(res[k])}' for k in VARIANTS) + f'   max diff first/last {err:.3g} {unit}', flush=True)
