"""kernel 5 (edge counts) behind `sample_ccdf` and `histogram_last_axis`: integer results, so the
comparison with the oracle and the committed reference outputs is exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
import iqwaveform_b200 as iqw
from iqwaveform_b200.power_analysis import sample_ccdf
from iqwaveform_b200.util import histogram_last_axis
from oracle import iqw_oracle as orc

pytestmark = pytest.mark.gpu


def test_golden():
    _, a = load_golden('ccdf_power_61')
    d = sample_ccdf(a['p'], a['edges'], density=True)
    assert isinstance(d, np.ndarray) and d.dtype == np.float64 and np.array_equal(d, a['density'])
    c = sample_ccdf(a['p'], a['edges'], density=False)
    assert c.dtype == np.int64 and np.array_equal(c, a['counts'])
    p, a = load_golden('hist_db_50')
    h, e = histogram_last_axis(a['x'], p['bins'], tuple(p['range']))
    assert np.array_equal(h, a['hist']) and np.array_equal(e, a['edges'])


@pytest.mark.parametrize('n', [1, 3, 1000, 100003, 5_000_000])
@pytest.mark.parametrize('edges', [np.linspace(0, 4, 41), np.array([1.0]), np.array([0.5, 1.0, 1.0, 2.0], dtype=np.float32),
                                   np.linspace(-3, 9, 4096)])
def test_sample_ccdf(n, edges):
    rng = np.random.default_rng(n)
    p = (rng.standard_normal(n) ** 2).astype(np.float32)
    p[::97] = 1.0
    p[::1013] = np.inf
    if n > 10:
        p[5] = np.nan
    for density in (True, False):
        assert np.array_equal(sample_ccdf(p, edges, density=density), orc.sample_ccdf(p, edges, density=density))
    got = sample_ccdf(torch.from_numpy(p).cuda(), torch.from_numpy(edges), density=False)
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), orc.sample_ccdf(p, edges, density=False))


@pytest.mark.parametrize('shape', [(4000,), (3, 5, 4001), (1000, 17), (2, 1 << 20)])
@pytest.mark.parametrize('bins,rg', [(40, (-2.0, 2.0)), (7, None), (np.array([-1.0, 0.0, 0.25, 3.0]), None), (4095, (-5.0, 5.0))])
def test_histogram_last_axis(shape, bins, rg):
    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape).astype(np.float32)
    x.reshape(-1)[:10] = 2.0          # exactly on the last edge of the first case: not counted
    h, e = histogram_last_axis(x, bins, rg)
    h2, e2 = orc.histogram_last_axis(x, bins, rg)
    assert h.shape == h2.shape and np.array_equal(h, h2) and np.array_equal(e, e2)


def test_counts_add_up_at_full_size():
    """2^28 dB values of a spectrogram: every sample lands in exactly one of the n_edges + 1 classes"""
    n = 1 << 28
    x = torch.randn(n, dtype=torch.complex64, device='cuda')
    p = iqw.spectrogram(x, fs=1e6, window='hann', nperseg=1024, noverlap=512, dB=True, return_axis_arrays=False)
    edges = np.linspace(-80, 20, 201)
    c = sample_ccdf(p.reshape(-1), edges, density=False)
    assert c.shape == (201,) and bool((c[:-1] >= c[1:]).all()) and int(c[0]) <= p.numel()
    h, _ = histogram_last_axis(p, edges)
    assert h.shape == (p.shape[0], 200)
    below = int((p < -80).sum())
    assert int(h.sum()) + below + int((p >= 20).sum()) == p.numel()


def test_errors():
    with pytest.raises(ValueError):
        sample_ccdf(np.zeros((3, 3), np.float32), np.array([0.0]))
    with pytest.raises(NotImplementedError):
        sample_ccdf(np.zeros(3, np.float64), np.array([0.0]))
    with pytest.raises(NotImplementedError):
        sample_ccdf(np.zeros(3, np.float32), np.linspace(0, 1, 5000))


@pytest.mark.parametrize('overlap,bw', [(True, None), (True, 0.5e6), (False, 0.75e6), (True, 1e6)])
def test_iq_to_stft_spectrogram(overlap, bw):
    import _tol
    from oracle.make_golden import synth
    x = synth(12, (20000,))
    got = iqw.iq_to_stft_spectrogram(x, 'hann', 256, 1e-6, overlap=overlap, analysis_bandwidth=bw)
    want = orc.iq_to_stft_spectrogram(x, 'hann', 256, 1e-6, overlap=overlap, analysis_bandwidth=bw)
    assert got.shape == want.shape and (want.size == 0 or got.values.dtype == np.float32)
    assert np.array_equal(got.columns.values, want.columns.values) and np.array_equal(got.index.values, want.index.values)
    if want.size:
        full = orc.iq_to_stft_spectrogram(x, 'hann', 256, 1e-6, overlap=overlap).values.astype(np.float64)
        tol = _tol.POWER_RTOL * np.abs(want.values) + _tol.POWER_FLOOR * full.max(axis=1, keepdims=True)
        assert np.all(np.abs(got.values - want.values) <= tol)
    with pytest.raises(ValueError):
        iqw.iq_to_stft_spectrogram(x, 'hann', 256, 1e-6, analysis_bandwidth=0.3333e6)


@pytest.mark.parametrize('cc,ov,bins', [(4, 0, 48), (4, 32, 48), (1, 0, 48), (8, 0, 62), (2, 0, 64), (1, 0, 64)])
def test_channelize_power(cc, ov, bins):
    from oracle.make_golden import synth
    x = synth(13, (30000,))
    kw = dict(analysis_bins_per_channel=bins, window='hann', channel_count=cc, fft_overlap_per_channel=ov)
    if bins == 64 and cc > 1:       # X[:, 0:-0] is empty and freqs[0] of nothing raises, in the oracle too
        for f in (iqw.channelize_power, orc.channelize_power):
            with pytest.raises(IndexError):
                f(x, 1e-6, 64, **kw)
        return
    got = iqw.channelize_power(x, 1e-6, 64, **kw)
    want = orc.channelize_power(x, 1e-6, 64, **kw)
    assert len(got) == len(want)
    for g, w in zip(got[:-1], want[:-1]):
        assert np.array_equal(g, w)
    g, w = got[-1], want[-1]
    assert g.shape == w.shape and g.dtype == np.float32
    if w.size:      # (bins == fft size: the reference's X[:, 0:-0] is empty)
        assert np.all(np.abs(g - w) <= 2e-5 * np.abs(w) + 2e-6 * w.max())
    with pytest.raises(ValueError):
        iqw.channelize_power(x, 1e-6, 64, analysis_bins_per_channel=65, window='hann')
    with pytest.raises(ValueError):
        iqw.channelize_power(x, 1e-6, 64, analysis_bins_per_channel=63, window='hann', channel_count=1)
    with pytest.raises(NotImplementedError):
        iqw.channelize_power(x, 1e-6, 64, analysis_bins_per_channel=48, window='hann', axis=1)


def test_sigmf_capture_feeds_the_persistence_spectrum(tmp_path, cuda_device):
    """a SigMF (npy) capture file -> read_sigmf (memory mapped) -> persistence spectrum of
    every capture segment, equal to the same call on the arrays"""
    import json
    import numpy as np
    import iqwaveform_b200 as iqw
    from oracle import iqw_oracle as orc
    from oracle.make_golden import synth
    n_seg, seg_len = 2, 1 << 16
    x = synth(9, (n_seg * seg_len,))
    meta = {'global': {'core:sample_rate': 1e6},
            'captures': [{'core:sample_start': i * seg_len, 'core:frequency': 1e9 + 1e6 * i,
                          'core:datetime': f'2024-01-01T00:00:0{i}Z'} for i in range(n_seg)], 'annotations': []}
    path = tmp_path / 'c.sigmf-meta'
    path.write_text(json.dumps(meta))
    np.save(tmp_path / 'c.sigmf-data.npy', x)
    kw = dict(window='hann', resolution=1e6 / 1024, fractional_overlap=0.5, statistics=[0.5, 'max'], dB=True)
    freqs, got = iqw.persistence_spectrum_from_sigmf(path, **kw)
    assert np.array_equal(freqs, [1e9, 1e9 + 1e6]) and got.shape == (2, 2, 1024)
    want = iqw.persistence_spectrum(x.reshape(n_seg, seg_len), fs=1e6, axis=1, **kw)
    assert np.array_equal(got, want)
    ref = orc.persistence_spectrum(x.reshape(n_seg, seg_len), fs=1e6, axis=1, **kw)
    assert np.abs(got - ref).max() < 1e-3


def test_graphed_call_replays_the_same_bits(cuda_device):
    """small problems: the whole persistence-spectrum call captured once as a CUDA graph and replayed on
    new data gives the bits of the plain call"""
    import torch
    import iqwaveform_b200 as iqw
    from oracle.make_golden import synth
    kw = dict(fs=15.36e6, window='hann', resolution=15e3, fractional_overlap=0.5, statistics=[0.5, 0.99, 'max'],
              dB=True, axis=1)
    a = torch.from_numpy(synth(1, (1, 1 << 21))).to(cuda_device)
    b = torch.from_numpy(synth(2, (1, 1 << 21))).to(cuda_device)
    g = iqw.GraphedCall(iqw.persistence_spectrum, a, **kw)
    assert torch.equal(g(), iqw.persistence_spectrum(a, **kw))
    assert torch.equal(g(b), iqw.persistence_spectrum(b, **kw))
    assert torch.equal(g(a), iqw.persistence_spectrum(a, **kw))
    with pytest.raises(TypeError):
        iqw.GraphedCall(iqw.persistence_spectrum, a.cpu(), **kw)
