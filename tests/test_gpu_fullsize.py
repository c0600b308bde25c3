"""size-independent properties at BASELINE.json's full sizes (no oracle: it would take minutes and
tens of GB on the host).  Each test states the property and why it pins the result."""
import math

import numpy as np
import pytest
import torch

import iqwaveform_b200 as iqw

pytestmark = pytest.mark.gpu


def device_capture(n, seed, device, tones=((0.0651, 0.5), (-0.2148, 0.05), (0.3256, 3.0))):
    """unit-variance complex noise + off-bin tones, generated on the device in chunks"""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty(n, dtype=torch.complex64, device=device)
    xr = torch.view_as_real(x)
    chunk = 1 << 26
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        xr[s:s + m].normal_(0.0, math.sqrt(0.5), generator=g)
        k = torch.arange(s, s + m, device=device, dtype=torch.float64)
        for f, a in tones:
            ph = (2 * math.pi) * torch.remainder(f * k, 1.0)
            xr[s:s + m, 0] += (a * torch.cos(ph)).float()
            xr[s:s + m, 1] += (a * torch.sin(ph)).float()
    return x


def test_config3_shard_order_statistics_are_exact(cuda_device):
    """config 3, one channel (10 s at 100 MS/s, nfft 4096, 50 % overlap, T = 488 280).
    Property: v is the r-th order statistic of a column iff #{x < v} <= r < #{x <= v}.  Counting
    with torch on the materialised spectrogram is independent of the selection kernel."""
    n, nfft = 1_000_000_000, 4096
    x = device_capture(n, 1234, cuda_device)
    stats = [0.1, 0.5, 0.9, 0.999]
    out = iqw.persistence_spectrum(x.view(1, -1), fs=100e6, window='hann', resolution=100e6 / nfft,
                                   fractional_overlap=0.5, statistics=stats + ['max', 'min'],
                                   dB=False, axis=1)[0]
    p = iqw.spectrogram(x, fs=100e6, window='hann', nperseg=nfft, noverlap=nfft // 2,
                        return_axis_arrays=False)
    del x
    T = p.shape[0]
    assert T == 488280 and tuple(out.shape) == (6, nfft)
    assert torch.equal(out[4], p.max(dim=0).values) and torch.equal(out[5], p.min(dim=0).values)
    from iqwaveform_b200._plan import quantile_plan
    sel = iqw.time_statistics(p.view(1, T, nfft), stats, dB=False)[0]
    assert torch.equal(sel, out[:4])            # fused path == standalone statistics kernel
    cols = torch.arange(0, nfft, 37, device=cuda_device)       # 111 columns, every region of the band
    pc = p[:, cols].contiguous()
    for i, q in enumerate(stats):
        lo, hi, g = quantile_plan(T, q)
        v = out[i, cols]
        # the lerp result lies between the two neighbouring order statistics
        below = (pc < v).sum(dim=0)
        below_eq = (pc <= v).sum(dim=0)
        assert torch.all(below <= hi) and torch.all(below_eq > lo), q
    # monotone in q, bounded by min / max
    assert torch.all(out[5] <= out[0]) and torch.all(out[0] <= out[1])
    assert torch.all(out[1] <= out[2]) and torch.all(out[2] <= out[3]) and torch.all(out[3] <= out[4])


def test_config2_spectrogram_parseval_and_tone(cuda_device):
    """config 2 (10 s at 100 MS/s, nfft 2048 Blackman-Harris, 50 % overlap, dB out, T = 976 561).
    Properties: (i) sum_k P[t,k] = (1/nfft) sum_n |c[n] x[n]|^2 * nfft  (Parseval, checked in float64
    on 4000 sampled frames incl. first and last => frame indexing); (ii) dB == 10 log10(power);
    (iii) the strong tone peaks in the bin its frequency maps to, in every sampled frame."""
    n, nfft, hop = 1_000_000_000, 2048, 1024
    x = device_capture(n, 99, cuda_device)
    p = iqw.spectrogram(x, fs=100e6, window='blackmanharris', nperseg=nfft, noverlap=hop,
                        return_axis_arrays=False)
    T = p.shape[0]
    assert T == 976561
    idx = torch.cat([torch.tensor([0, 1, T - 2, T - 1]),
                     torch.randint(0, T, (3996,), generator=torch.Generator().manual_seed(0))]).to(cuda_device)
    from iqwaveform_b200._plan import stft_coefficients
    c = torch.from_numpy(stft_coefficients('blackmanharris', nfft, 0, 'power', hop)).to(cuda_device).double()
    offs = idx[:, None] * hop + torch.arange(nfft, device=cuda_device)[None, :]
    frames = x[offs]
    energy = ((frames.real.double() * c) ** 2 + (frames.imag.double() * c) ** 2).sum(dim=1) * nfft
    total = p[idx].double().sum(dim=1)
    assert torch.allclose(total, energy, rtol=2e-6)
    peak_bin = round(0.3256 * nfft) + nfft // 2
    assert torch.all((p[idx].argmax(dim=1) - peak_bin).abs() <= 1)
    sub = p[idx].clone()
    del p
    d = iqw.spectrogram(x, fs=100e6, window='blackmanharris', nperseg=nfft, noverlap=hop,
                        return_axis_arrays=False, dB=True)
    assert torch.allclose(d[idx], 10 * torch.log10(sub), atol=5e-5)


def test_config4_bin_power_against_torch(cuda_device):
    """config 4 shape (1 ms bins at 245.76 MS/s = 245 760 samples per bin), 4 s of capture.
    Property: mean/max/min of |x|^2 per bin equal a float64 torch evaluation of the same samples."""
    nb, bins = 245760, 4000
    x = device_capture(nb * bins, 7, cuda_device)
    mean = iqw.iq_to_bin_power(x, 1 / 245.76e6, 1e-3, kind='mean')
    peak = iqw.iq_to_bin_power(x, 1 / 245.76e6, 1e-3, kind='peak')
    low = iqw.iq_to_bin_power(x, 1 / 245.76e6, 1e-3, kind='min')
    assert tuple(mean.shape) == (bins,)
    xr = torch.view_as_real(x).view(bins, nb, 2)
    for b0 in range(0, bins, 500):
        blk = xr[b0:b0 + 500]
        pw = blk[..., 0] * blk[..., 0] + blk[..., 1] * blk[..., 1]
        assert torch.allclose(mean[b0:b0 + 500].double(), pw.double().mean(dim=1), rtol=2e-6)
        assert torch.allclose(peak[b0:b0 + 500], pw.max(dim=1).values, rtol=1e-6)
        assert torch.allclose(low[b0:b0 + 500], pw.min(dim=1).values, rtol=1e-6, atol=1e-12)


def test_frame_permutation_invariance(cuda_device):
    """without overlap, permuting whole frames of the capture permutes spectrogram rows, so every
    order statistic must be BITWISE unchanged (exactness of selection, independent of row order)."""
    nfft, T = 1024, 200_000
    x = device_capture(nfft * T, 5, cuda_device)
    perm = torch.randperm(T, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(3))
    xp = x.view(T, nfft)[perm].reshape(-1)
    kw = dict(fs=1e8, window='hann', resolution=1e8 / nfft, fractional_overlap=0,
              statistics=[0.01, 0.5, 0.99, 'max', 'min', 'median'], axis=1)
    a = iqw.persistence_spectrum(x.view(1, -1), **kw)
    b = iqw.persistence_spectrum(xp.view(1, -1), **kw)
    assert torch.equal(a, b)


def test_channels_pipelined_on_two_streams_equal_one_by_one():
    """(C, N) input large enough for the two-stream pipeline over channels: same bits as channel by channel"""
    import iqwaveform_b200.fourier as F
    g = torch.Generator('cuda').manual_seed(77)
    x = torch.randn((3, 1 << 25), dtype=torch.complex64, device='cuda', generator=g)
    kw = dict(fs=1e8, window='hann', resolution=1e8 / 2048, fractional_overlap=0.5,
              statistics=[0.1, 0.5, 'mean', 0.999, 'max'], dB=True, axis=1)
    assert (x.shape[1] // 1024) * 2048 * 4 >= F.PIPELINE_MIN_BYTES
    together = iqw.persistence_spectrum(x, **kw)
    single = torch.stack([iqw.persistence_spectrum(x[c], **dict(kw, axis=0)) for c in range(3)])
    assert torch.equal(together, single)
    # the caller's stream sees finished results: an immediate consumer on it is correct
    assert torch.equal(iqw.persistence_spectrum(x, **kw) + 0, single)
