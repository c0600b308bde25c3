"""GPU parity: the CUDA path (through the C-ABI) against the numpy oracle on identical seeded
inputs and against the committed golden fixtures (outputs of the unmodified reference).
Tolerances are defined in tests/_tol.py."""
import numpy as np
import pytest
import torch

import _tol
from conftest import load_golden
import iqwaveform_b200 as iqw
from oracle import iqw_oracle as orc
from oracle.make_golden import synth

pytestmark = pytest.mark.gpu

NFFTS = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]


def dev_of(x, cuda_device):
    return torch.from_numpy(x).to(cuda_device)


# ---------------------------------------------------------------------------------------------
# stft
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('nfft', NFFTS)
@pytest.mark.parametrize('overlap', [0.0, 0.5, 0.75])
def test_stft_vs_oracle(cuda_device, nfft, overlap):
    nov = int(nfft * overlap)
    x = synth(nfft + 1, (2, nfft * 23 + 17))             # ragged tail on purpose
    for norm in ('power', None):
        f, t, y = orc.stft(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm=norm)
        f2, t2, y2 = iqw.stft(dev_of(x, cuda_device), fs=1e6, window='hann', nperseg=nfft,
                              noverlap=nov, axis=1, norm=norm)
        assert isinstance(f2, np.ndarray) and np.array_equal(f, f2) and np.array_equal(t, t2)
        assert y2.dtype == torch.complex64 and tuple(y2.shape) == y.shape
        err = np.abs(y2.cpu().numpy() - y) / _tol.complex_tol(y)
        assert err.max() <= 1.0, (norm, err.max())


@pytest.mark.parametrize('name', ['stft_hann_256_128_power', 'stft_hann_256_128_cola',
                                  'stft_bh_512_0_power', 'stft_kaiser_64_48_power',
                                  'stft_hann_1024_512_nzero'])
def test_stft_vs_golden(cuda_device, name):
    p, a = load_golden(name)
    f, t, y = iqw.stft(dev_of(a['x'], cuda_device), **p)
    assert np.array_equal(f, a['freqs']) and np.array_equal(t, a['times'])
    assert tuple(y.shape) == a['y'].shape
    assert (np.abs(y.cpu().numpy() - a['y']) / _tol.complex_tol(a['y'])).max() <= 1.0


def test_stft_layouts(cuda_device):
    """1-D axis 0, (C,N) axis 1, (N,C) axis 0 with and without overlap, negative axis"""
    x = synth(3, (3, 5000))
    _, _, want = orc.stft(x, fs=1.0, window='hann', nperseg=128, noverlap=64, axis=1, norm='power')
    xd = dev_of(x, cuda_device)
    got1 = iqw.stft(xd[0], fs=1.0, window='hann', nperseg=128, noverlap=64, axis=0, norm='power',
                    return_axis_arrays=False)
    assert tuple(got1.shape) == want.shape[1:]
    np.testing.assert_allclose(got1.cpu().numpy(), want[0], atol=2e-7 * np.abs(want).max())
    got2 = iqw.stft(xd, fs=1.0, window='hann', nperseg=128, noverlap=64, axis=-1, norm='power',
                    return_axis_arrays=False)
    np.testing.assert_allclose(got2.cpu().numpy(), want, atol=2e-7 * np.abs(want).max())
    # time axis first: the reference (noverlap=0) returns (T, nfft, C)
    _, _, w0 = orc.stft(x, fs=1.0, window='hann', nperseg=128, noverlap=0, axis=1, norm='power')
    got3 = iqw.stft(xd.T.contiguous(), fs=1.0, window='hann', nperseg=128, noverlap=0, axis=0,
                    norm='power', return_axis_arrays=False)
    assert tuple(got3.shape) == (w0.shape[1], 128, 3)
    np.testing.assert_allclose(got3.cpu().numpy(), np.moveaxis(w0, 0, -1), atol=2e-7 * np.abs(w0).max())


def test_known_answers_on_gpu(cuda_device):
    nfft, hop, n = 64, 16, 1000
    x = np.arange(n).astype(np.complex64)
    _, t, y = iqw.stft(dev_of(x, cuda_device), fs=1.0, window='rect', nperseg=nfft,
                       noverlap=nfft - hop, norm='power')
    T = (n - nfft) // hop + 1
    m = np.arange(T)
    assert tuple(y.shape) == (T, nfft) and np.array_equal(t, m * float(hop))
    np.testing.assert_allclose(y[:, nfft // 2].real.cpu().numpy(), m * hop + (nfft - 1) / 2, rtol=1e-6)
    s = 333
    x = np.zeros(n, np.complex64); x[s] = 1
    p = iqw.spectrogram(dev_of(x, cuda_device), fs=1.0, window='rect', nperseg=nfft,
                        noverlap=nfft - hop, return_axis_arrays=False).cpu().numpy()
    hit = (m * hop <= s) & (s < m * hop + nfft)
    assert np.all(p[hit].min(axis=1) > 0) and np.all(p[~hit] == 0)
    k, A = 5, 2.0
    x = (A * np.exp(2j * np.pi * k * np.arange(n) / nfft)).astype(np.complex64)
    p = iqw.spectrogram(dev_of(x, cuda_device), fs=1.0, window='rect', nperseg=nfft, noverlap=0,
                        return_axis_arrays=False).cpu().numpy()
    assert np.all(p.argmax(axis=1) == k + nfft // 2)
    np.testing.assert_allclose(p.max(axis=1), A * A, rtol=1e-5)


# ---------------------------------------------------------------------------------------------
# spectrogram (power, dB)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('nfft,overlap,window', [(64, 0.75, 'hann'), (256, 0.5, 'hann'),
                                                 (1024, 0.5, 'hann'), (2048, 0.5, 'blackmanharris'),
                                                 (4096, 0.5, 'hann'), (8192, 0.75, ('kaiser', 10.0)),
                                                 (16384, 0.5, 'hann'), (65536, 0.75, 'blackmanharris'),
                                                 (4096, 0.0, 'rect')])
def test_spectrogram_vs_oracle_and_truth(cuda_device, nfft, overlap, window):
    nov = int(nfft * overlap)
    x = synth(7 * nfft, (2, nfft * 30 + 5))
    _, _, ref = orc.spectrogram(x, fs=1e8, window=window, nperseg=nfft, noverlap=nov, axis=1)
    truth = orc.stft_power_f64(x, window=window, nperseg=nfft, noverlap=nov)
    got = iqw.spectrogram(dev_of(x, cuda_device), fs=1e8, window=window, nperseg=nfft,
                          noverlap=nov, axis=1, return_axis_arrays=False)
    assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape
    got = got.cpu().numpy()
    assert _tol.power_err_units(got, ref) <= 1.0
    # at least as close to the float64 evaluation as the reference is
    e_gpu, e_ref = _tol.power_err_units(got, truth), _tol.power_err_units(ref, truth)
    assert e_gpu <= max(1.25 * e_ref, 0.1), (e_gpu, e_ref)
    # fused dB output == powtodB(spectrogram) of the reference
    dref = orc.powtodB(ref.copy())
    dgot = iqw.spectrogram(dev_of(x, cuda_device), fs=1e8, window=window, nperseg=nfft,
                           noverlap=nov, axis=1, return_axis_arrays=False, dB=True).cpu().numpy()
    finite = np.isfinite(dref)
    tol = _tol.db_tol(dref, ref.max(axis=-1, keepdims=True))
    assert np.all(np.abs(dgot - dref)[finite] <= tol[finite])
    assert np.array_equal(np.isneginf(dgot), np.isneginf(dref)) or not np.any(~finite)


@pytest.mark.parametrize('name', ['spg_bh_2048_1024', 'spg_hann_1024_768', 'spg_rect_4096_0'])
def test_spectrogram_vs_golden(cuda_device, name):
    p, a = load_golden(name)
    f, t, got = iqw.spectrogram(dev_of(a['x'], cuda_device), **p)
    assert np.array_equal(f, a['freqs']) and np.array_equal(t, a['times'])
    assert _tol.power_err_units(got.cpu().numpy(), a['power']) <= 1.0
    _, _, d = iqw.spectrogram(dev_of(a['x'], cuda_device), dB=True, **p)
    tol = _tol.db_tol(a['dB'], a['power'].max(axis=-1, keepdims=True))
    assert np.all(np.abs(d.cpu().numpy() - a['dB']) <= tol)


def test_zero_input_gives_minus_inf_dB(cuda_device):
    x = torch.zeros(4096, dtype=torch.complex64, device=cuda_device)
    d = iqw.spectrogram(x, fs=1.0, window='hann', nperseg=256, return_axis_arrays=False, dB=True)
    assert torch.all(torch.isneginf(d))           # reference: log10(0) = -inf, silently
    d = iqw.spectrogram(x, fs=1.0, window='hann', nperseg=256, return_axis_arrays=False, dB=True, eps=1e-25)
    np.testing.assert_allclose(d.cpu().numpy(), -250.0, atol=1e-4)


def test_host_arrays_round_trip(cuda_device):
    """numpy in -> numpy out, CPU torch in -> CPU torch out (host<->device copies inside)"""
    x = synth(9, (2, 40000))
    _, _, ref = orc.spectrogram(x, fs=1e6, window='hann', nperseg=1024, noverlap=512, axis=1)
    f, t, got = iqw.spectrogram(x, fs=1e6, window='hann', nperseg=1024, noverlap=512, axis=1)
    assert isinstance(got, np.ndarray) and _tol.power_err_units(got, ref) <= 1.0
    got = iqw.spectrogram(torch.from_numpy(x), fs=1e6, window='hann', nperseg=1024, noverlap=512,
                          axis=1, return_axis_arrays=False)
    assert isinstance(got, torch.Tensor) and not got.is_cuda
    assert _tol.power_err_units(got.numpy(), ref) <= 1.0


def test_unsupported_sizes_fail_loudly(cuda_device):
    x = torch.zeros(1 << 18, dtype=torch.complex64, device=cuda_device)
    with pytest.raises(NotImplementedError):
        iqw.spectrogram(x, fs=1.0, window='hann', nperseg=1001)         # odd: complex phase-ramp window
    with pytest.raises(NotImplementedError):
        iqw.spectrogram(x, fs=1.0, window='hann', nperseg=40000)        # not a power of two and > 32768
    with pytest.raises(NotImplementedError):
        iqw.spectrogram(x, fs=1.0, window="hann", nperseg=131072)
    with pytest.raises(NotImplementedError):
        iqw.spectrogram(x.to(torch.complex128), fs=1.0, window='hann', nperseg=64)


# ---------------------------------------------------------------------------------------------
# persistence spectrum
# ---------------------------------------------------------------------------------------------
def _check_persistence(got, ref, stats, x, nfft, dB, peak):
    isq = orc.find_float_inds(stats)
    for i, s in enumerate(stats):
        g, r = got[:, i].astype(np.float64), ref[:, i].astype(np.float64)
        if dB:
            tol = _tol.db_tol(r, peak)
            if s in ('mean', 'rms'):
                tol = np.maximum(tol, 1e-3)       # reference's fp32 running sum
            assert np.all(np.abs(g - r) <= tol), (s, np.max(np.abs(g - r) / tol))
        else:
            tol = _tol.POWER_RTOL * np.abs(r) + _tol.POWER_FLOOR * peak
            if s in ('mean', 'rms'):
                tol = tol + 2e-5 * np.abs(r)
            assert np.all(np.abs(g - r) <= tol), (s, np.max(np.abs(g - r) / tol))


@pytest.mark.parametrize('n,nfft,ovl,stats,kw', [
    (1 << 18, 1024, 0.5, [0.5, 0.99, 'mean', 'max'], {}),
    (1 << 19, 1024, 0.5, [0.1, 0.5, 0.9, 0.999, 'min', 'median', 'rms', 'peak'], {}),
    (1 << 19, 4096, 0.5, [0.1, 0.5, 0.9, 0.999], dict(bandwidth=0.5e6)),
    (1 << 16, 256, 0.75, ['mean', 0.25, 'max', 1.0, 0.0], dict(dB=False)),
    (1 << 17, 512, 0.5, ['0.5', 'max'], dict(fractional_window=0.75)),
    (1 << 17, 2048, 0.0, [i / 10 for i in range(1, 10)], {}),          # 18 ranks -> split calls
])
def test_persistence_vs_oracle(cuda_device, n, nfft, ovl, stats, kw):
    x = synth(n % 89, (2, n))
    args = dict(fs=1e6, window='hann', resolution=1e6 / nfft, fractional_overlap=ovl,
                statistics=stats, axis=1, **kw)
    ref = orc.persistence_spectrum(x, **args)
    got = iqw.persistence_spectrum(dev_of(x, cuda_device), **args)
    assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape
    nz = round((1 - kw.get('fractional_window', 1)) * nfft)
    _, _, p = orc.spectrogram(x, fs=1e6, window='hann', nperseg=nfft, noverlap=round(ovl * nfft),
                              nzero=nz, axis=1)
    peak = p.max(axis=(1, 2))[:, None]
    _check_persistence(got.cpu().numpy(), ref, stats, x, nfft, kw.get('dB', True), peak)


@pytest.mark.parametrize('name', ['psd_hann_1024_half', 'psd_hann_4096_trim', 'psd_bh_256_linear'])
def test_persistence_vs_golden(cuda_device, name):
    p, a = load_golden(name)
    got = iqw.power_spectral_density(dev_of(a['x'], cuda_device), **p).cpu().numpy()
    isq = a['is_quantile']
    ref = np.empty_like(got)
    ref[:, isq] = a['quantile_rows']
    ref[:, ~isq] = a['named_rows']
    nfft = round(p['fs'] / p['resolution'])
    _, _, pw = orc.spectrogram(a['x'], fs=p['fs'], window=p['window'], nperseg=nfft,
                               noverlap=round(p['fractional_overlap'] * nfft), axis=1)
    _check_persistence(got, ref, p['statistics'], a['x'], nfft, p.get('dB', True),
                       pw.max(axis=(1, 2))[:, None])


def test_persistence_config0_full_size_vs_oracle(cuda_device):
    """BASELINE configs[0] at FULL size against the oracle (fourier.py:1236-1327): 1 s of complex64
    noise + tones at 15.36 MS/s, nfft 1024 Hann, 50 % overlap, q = [0.5, 0.99], dB.  T = 29 999 frames:
    the long-column path of kernel 2 (row sample -> brackets -> one read of the matrix) on real STFT
    data.  About 2 s of CPU for the oracle."""
    n = 15_360_000
    x = synth(77, (1, n))
    args = dict(fs=15.36e6, window='hann', resolution=15e3, fractional_overlap=0.5,
                statistics=[0.5, 0.99], dB=True, axis=1)
    ref = orc.persistence_spectrum(x, **args)
    got = iqw.persistence_spectrum(dev_of(x, cuda_device), **args)
    assert tuple(got.shape) == ref.shape == (1, 2, 1024)
    _, _, p = orc.spectrogram(x, fs=15.36e6, window='hann', nperseg=1024, noverlap=512, axis=1)
    assert p.shape[1] == 29999
    _check_persistence(got.cpu().numpy(), ref, [0.5, 0.99], x, 1024, True, p.max(axis=(1, 2))[:, None])
    # the selection is exact: on the device's own spectrogram the rows are bitwise numpy's quantiles
    pd = iqw.spectrogram(dev_of(x, cuda_device), fs=15.36e6, window='hann', nperseg=1024, noverlap=512,
                         axis=1, return_axis_arrays=False)
    cnt = []
    sel = iqw.time_statistics(pd, [0.5, 0.99], dB=False, counters=cnt).cpu().numpy()
    want = np.quantile(pd.cpu().numpy(), np.array([0.5, 0.99], dtype=np.float32), axis=1)
    assert np.array_equal(sel, np.moveaxis(want, 0, 1))
    assert cnt[0]['inconsistent'] == 0


def test_persistence_long_path_two_channels_vs_oracle(cuda_device):
    """(2, 2^23) samples, nfft 1024, 50 % overlap: T = 16 383 -> the long-column path, with mean, max and
    four quantiles, two channels (the two-stream channel pipeline is not taken at this size)"""
    n, nfft = 1 << 23, 1024
    x = synth(5, (2, n))
    stats = ['mean', 'max', 0.1, 0.5, 0.9, 0.999]
    args = dict(fs=1e6, window='hann', resolution=1e6 / nfft, fractional_overlap=0.5, statistics=stats,
                dB=True, axis=1)
    ref = orc.persistence_spectrum(x, **args)
    got = iqw.persistence_spectrum(dev_of(x, cuda_device), **args)
    assert tuple(got.shape) == ref.shape == (2, 6, nfft)
    _, _, p = orc.spectrogram(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nfft // 2, axis=1)
    assert p.shape[1] == 16383
    _check_persistence(got.cpu().numpy(), ref, stats, x, nfft, True, p.max(axis=(1, 2))[:, None])


def test_persistence_1d_and_config1_shape(cuda_device):
    """BASELINE config 1 at a tenth of its length, 1-D input (the full size is the test above)"""
    x = synth(21, (1, 1536000))
    args = dict(fs=15.36e6, window='hann', resolution=15e3, fractional_overlap=0.5,
                statistics=[0.5, 0.99], dB=True)
    ref = orc.persistence_spectrum(x, axis=1, **args)
    got = iqw.persistence_spectrum(dev_of(x[0], cuda_device), axis=0, **args)
    assert tuple(got.shape) == (2, 1024)
    _, _, p = orc.spectrogram(x, fs=15.36e6, window='hann', nperseg=1024, noverlap=512, axis=1)
    _check_persistence(got.cpu().numpy()[None], ref, [0.5, 0.99], x, 1024, True, p.max())


# ---------------------------------------------------------------------------------------------
# exact order statistics (no FFT involved): bitwise against numpy
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('T,nb', [(1, 5), (2, 130), (3, 1), (77, 64), (1025, 33), (5000, 300),
                                  (70001, 257), (300000, 128)])
def test_time_statistics_bitwise(cuda_device, T, nb):
    rng = np.random.default_rng(T)
    a = rng.standard_normal((2, T, nb)).astype(np.float32)
    a[0, :, 0] = 1.5                                    # constant column
    if nb > 1:
        a[0, :, 1] = np.round(a[0, :, 1])               # heavy ties
    if nb > 2 and T > 10:
        a[1, 3, 2], a[1, 4, 2] = 1e30, -1e30            # outliers far outside the sampled range
        a[1, :, 3 % nb] = np.where(np.arange(T) % 2 == 0, 0.0, -0.0)   # signed zeros
    if nb > 4:
        a[1, :, 4] = np.exp(6 * a[1, :, 4])             # 50 dB of dynamic range
        a[0, :, 4] = np.arange(T, dtype=np.float32)     # sorted ramp
    qs = [0.0, 0.1, 0.5, 0.999, 1.0]
    got = iqw.time_statistics(dev_of(a, cuda_device), qs + ['median', 'min', 'max', 'mean'],
                              dB=False).cpu().numpy()
    want = np.quantile(a, np.array(qs, dtype=np.float32), axis=1)
    for i in range(len(qs)):
        assert np.array_equal(got[:, i], want[i]), qs[i]
    assert np.array_equal(got[:, 5], np.median(a, axis=1))
    assert np.array_equal(got[:, 6], a.min(axis=1)) and np.array_equal(got[:, 7], a.max(axis=1))
    truth = a.astype(np.float64).mean(axis=1)
    np.testing.assert_allclose(got[:, 8], truth, rtol=1e-6, atol=1e-7 * np.abs(a).max())


@pytest.mark.parametrize('sigmas,extra', [(0.0, 0), (0.5, 0), (5.0, 2)])
def test_sampled_path_is_exact_even_when_brackets_miss(cuda_device, sigmas, extra):
    """long columns take the row-sample -> bracket path; with a zero margin about half of the
    brackets miss their rank and the result must still be bitwise equal to numpy"""
    from iqwaveform_b200 import _lib
    rng = np.random.default_rng(11)
    T, nb = 120000, 200
    a = rng.standard_normal((1, T, nb)).astype(np.float32)
    a[0, :, 0] = np.sort(a[0, :, 0])                       # trend: early rows small, late rows large
    a[0, :, 1] = np.where(np.arange(T) % 7 == 3, 100.0, a[0, :, 1])      # periodic bursts
    a[0, :, 2] = np.round(a[0, :, 2] * 2) / 2              # few distinct values, heavy ties
    a[0, :, 3] = 0.25                                      # constant
    a[0, T // 2:, 4] += 50.0                               # level shift half way (bimodal)
    a[0, :, 5] = np.exp(8 * a[0, :, 5])                    # 70 dB of dynamic range
    qs = [0.001, 0.1, 0.5, 0.999]
    try:
        _lib.lib.iqw_debug_set_sample_margin(sigmas, extra)
        got = iqw.time_statistics(dev_of(a, cuda_device), qs + ['min', 'max'], dB=False).cpu().numpy()
    finally:
        _lib.lib.iqw_debug_set_sample_margin(5.0, 2)
    want = np.quantile(a, np.array(qs, dtype=np.float32), axis=1)
    for i in range(len(qs)):
        assert np.array_equal(got[:, i], want[i]), (qs[i], np.argwhere(got[:, i] != want[i])[:5])
    assert np.array_equal(got[:, 4], a.min(axis=1)) and np.array_equal(got[:, 5], a.max(axis=1))


@pytest.mark.parametrize('sigmas,extra', [(0.0, 0), (0.5, 0), (5.0, 2)])
def test_sampled_raw_mode_is_exact_when_brackets_miss(cuda_device, sigmas, extra):
    """the same on data whose brackets are all non-negative: the bracket pass compares raw float bits (difference
    form, negative values clamped), select counts the keys inside every bracket and moves ranks beyond them on to
    the gap that follows.  A few negative values, signed zeros and +inf sit below / above every bracket."""
    from iqwaveform_b200 import _lib
    rng = np.random.default_rng(12)
    T, nb = 120000, 136
    a = np.abs(rng.standard_normal((1, T, nb))).astype(np.float32) + np.float32(0.01)
    a[0, :, 0] = np.sort(a[0, :, 0])                       # trend: the row sample is biased
    a[0, ::7919, 1] = -3.0                                 # 16 negative values: below every (positive) bracket
    a[0, 5::9973, 1] = -0.0
    a[0, :, 2] = np.round(a[0, :, 2] * 4) / 4 + 0.25       # few distinct values, heavy ties
    a[0, :, 3] = 0.25                                      # constant
    a[0, T // 2:, 4] += 50.0                               # level shift half way (bimodal)
    a[0, :, 5] = np.exp(8 * a[0, :, 5])                    # 70 dB of dynamic range
    a[0, 100, 6] = np.inf
    qs = [0.01, 0.1, 0.5, 0.999]
    cnt = []
    try:
        _lib.lib.iqw_debug_set_sample_margin(sigmas, extra)
        got = iqw.time_statistics(dev_of(a, cuda_device), qs + ['min', 'max'], dB=False, counters=cnt).cpu().numpy()
    finally:
        _lib.lib.iqw_debug_set_sample_margin(5.0, 2)
    want = np.quantile(a, np.array(qs, dtype=np.float32), axis=1)
    for i in range(len(qs)):
        assert np.array_equal(got[:, i], want[i]), (qs[i], np.argwhere(got[:, i] != want[i])[:5])
    assert np.array_equal(got[:, 4], a.min(axis=1)) and np.array_equal(got[:, 5], a.max(axis=1))
    assert cnt[0]['key_mode'] == 0 and cnt[0]['inconsistent'] == 0
    if sigmas == 0.0:
        assert cnt[0]['missed_ranks'] > nb // 4            # the fallback through the gap intervals really ran


def test_long_path_with_many_brackets(cuda_device):
    """8 well separated single-rank statistics on a long matrix through the C-ABI: 8 brackets -> the
    M = 8 instantiations of the bracket pass (generic C++ body) and of select; negative data -> key
    mode instead of raw bits.  Then 3 quantiles of non-negative data: raw-bit mode."""
    import ctypes
    from iqwaveform_b200 import _lib
    rng = np.random.default_rng(5)
    T, nb = 50001, 130
    a = rng.standard_normal((1, T, nb)).astype(np.float32)
    a[0, :, 1] = np.abs(a[0, :, 1])
    ranks = [500, 6000, 12000, 20000, 27000, 34000, 42000, 49500]
    reqs = (_lib.iqw_stat * 8)()
    for r, k in zip(reqs, ranks):
        r.kind, r.rank_lo, r.rank_hi, r.gamma = _lib.STAT_QUANTILE, k, k, 0.0
    ad = dev_of(a, cuda_device)
    out = torch.empty((1, 8, nb), dtype=torch.float32, device=cuda_device)
    nws = _lib.lib.iqw_time_stats_workspace_bytes(1, T, nb, 8)
    ws = torch.empty(nws, dtype=torch.uint8, device=cuda_device)
    _lib.check(_lib.lib.iqw_time_stats_f32(ctypes.c_void_p(ad.data_ptr()), 1, T, nb, T * nb, reqs, 8, 0, 0.0,
                                           ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws.data_ptr()), nws,
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    c16 = (ctypes.c_uint32 * 16)()
    _lib.check(_lib.lib.iqw_debug_time_stats_counters(ctypes.c_void_p(ws.data_ptr()), nb, c16))
    srt = np.sort(a[0], axis=0)
    assert np.array_equal(out.cpu().numpy()[0], srt[ranks])
    assert c16[7] >= 8 * nb - 16 and c16[6] == 0 and c16[8] == 1        # selected, inconsistent, key mode
    # non-negative data, 3 brackets: raw-bit mode, everything settled from the candidate lists
    p = np.abs(a)
    cnt = []
    got = iqw.time_statistics(dev_of(p, cuda_device), [0.1, 0.5, 0.9], dB=False, counters=cnt).cpu().numpy()
    want = np.quantile(p, np.array([0.1, 0.5, 0.9], dtype=np.float32), axis=1)
    assert np.array_equal(got, np.moveaxis(want, 0, 1))
    assert cnt[0]['key_mode'] == 0 and cnt[0]['refine'] == 0 and cnt[0]['inconsistent'] == 0


def test_time_statistics_dB_is_monotone_image(cuda_device):
    rng = np.random.default_rng(2)
    p = rng.exponential(1e-4, (1, 20000, 96)).astype(np.float32)
    qs = [0.1, 0.5, 0.9]
    got = iqw.time_statistics(dev_of(p, cuda_device), qs + ['max', 'min', 'mean'], dB=True,
                              eps=1e-25).cpu().numpy()
    d = orc.powtodB(p.copy(), eps=1e-25)
    want = np.quantile(d, np.array(qs, dtype=np.float32), axis=1)
    np.testing.assert_allclose(got[:, :3], np.moveaxis(want, 0, 1), atol=5e-5)
    np.testing.assert_allclose(got[:, 3], d.max(axis=1), atol=5e-5)
    np.testing.assert_allclose(got[:, 4], d.min(axis=1), atol=5e-5)
    np.testing.assert_allclose(got[:, 5], d.astype(np.float64).mean(axis=1), atol=5e-5)


# ---------------------------------------------------------------------------------------------
# bin power
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('kind', ['mean', 'rms', 'max', 'peak', 'min', 'median', 0.25])
@pytest.mark.parametrize('nbin', [1, 7, 100, 1536, 50000, 150000])
def test_bin_power_vs_oracle(cuda_device, kind, nbin):
    x = synth(3, (3, 300001))
    ref = orc.iq_to_bin_power(x, 1.0, float(nbin), kind=kind, axis=1, truncate=True)
    got = iqw.iq_to_bin_power(dev_of(x, cuda_device), 1.0, float(nbin), kind=kind, axis=1, truncate=True)
    assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape
    rtol = 2e-6 if kind in ('mean', 'rms') else 1e-6
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=rtol)


def test_bin_power_golden_and_layouts(cuda_device):
    p, a = load_golden('binpower_1536')
    xd = dev_of(a['x'], cuda_device)
    for kind in ('mean', 'max', 'min', 'median', 'rms', 'peak'):
        got = iqw.iq_to_bin_power(xd, kind=kind, **p).cpu().numpy()
        np.testing.assert_allclose(got, a[kind], rtol=2e-6)
    np.testing.assert_allclose(iqw.iq_to_bin_power(xd, kind=0.25, **p).cpu().numpy(), a['q25'], rtol=1e-6)
    # 1-D and time-axis-first layouts
    q = dict(p, axis=0)
    got = iqw.iq_to_bin_power(xd[1], kind='mean', **q).cpu().numpy()
    np.testing.assert_allclose(got, a['mean'][1], rtol=2e-6)
    got = iqw.iq_to_bin_power(xd.T.contiguous(), kind='max', **q).cpu().numpy()
    np.testing.assert_allclose(got, a['max'].T, rtol=1e-6)
    # unaligned start (odd element offset -> scalar peel path)
    got = iqw.iq_to_bin_power(xd[0, 1:], 1.0, 1537.0, kind='mean', truncate=True).cpu().numpy()
    ref = orc.iq_to_bin_power(a['x'][0, 1:], 1.0, 1537.0, kind='mean', truncate=True)
    np.testing.assert_allclose(got, ref, rtol=2e-6)


def test_persistence_streams_host_captures(cuda_device, monkeypatch):
    """host captures above STREAM_MIN_BYTES are copied in chunks with the STFT following each chunk:
    same frames, same kernels -> bitwise equal to the device-resident call, for any chunking"""
    from iqwaveform_b200 import fourier
    x = synth(31, (3, 300007))
    kw = dict(fs=1e6, window='hann', resolution=1e6 / 1024, fractional_overlap=0.75,
              statistics=[0.1, 0.5, 0.999, 'max', 'mean'], axis=1)
    want = iqw.persistence_spectrum(dev_of(x, cuda_device), **kw).cpu().numpy()
    monkeypatch.setattr(fourier, 'STREAM_MIN_BYTES', 1)
    for chunks in (1, 5, 16, 1000):
        monkeypatch.setattr(fourier, 'STREAM_CHUNKS', chunks)
        got = iqw.persistence_spectrum(x, **kw)                       # numpy in -> numpy out
        assert isinstance(got, np.ndarray) and np.array_equal(got, want), chunks
    pinned = torch.from_numpy(x[0]).pin_memory()
    got = iqw.persistence_spectrum(pinned, **dict(kw, axis=0))        # 1-D pinned torch in -> CPU torch out
    assert isinstance(got, torch.Tensor) and not got.is_cuda and np.array_equal(got.numpy(), want[0])


# ---------------------------------------------------------------------------------------------
# kernel-1 geometries at nfft 1024 / 2048 / 4096: three-pass, two-pass with global loads, two-pass
# with bulk-copy (TMA) staging -- same transform, every variant against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('variant', [1, 2, 3])
@pytest.mark.parametrize('nfft,nov', [(1024, 512), (2048, 1536), (4096, 2048), (4096, 0), (1024, 511), (2048, 1)])
def test_stft_kernel_variants_vs_oracle(cuda_device, variant, nfft, nov):
    """(1024, 511) and (2048, 1) have an odd hop: frame starts are not 16-byte aligned, so the staged
    kernel must hand over to the global-load one (variant 3 is a request, not a guarantee)"""
    from iqwaveform_b200 import _lib
    x = synth(nfft + nov, (3, nfft * 9 + 13))
    _, _, ref = orc.stft(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm='power')
    _, _, pref = orc.spectrogram(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1)
    try:
        _lib.check(_lib.lib.iqw_debug_set_stft_variant(variant))
        y = iqw.stft(dev_of(x, cuda_device), fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1,
                     norm='power', return_axis_arrays=False).cpu().numpy()
        p = iqw.spectrogram(dev_of(x, cuda_device), fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1,
                            return_axis_arrays=False).cpu().numpy()
        # odd channel stride (a view of a wider buffer) and a band trim
        wide = torch.zeros((3, x.shape[1] + 1), dtype=torch.complex64, device=cuda_device)
        wide[:, :-1] = dev_of(x, cuda_device)
        pb = iqw.persistence_spectrum(wide[:, :-1], fs=1e6, window='hann', resolution=1e6 / nfft,
                                      fractional_overlap=nov / nfft, statistics=['max'], dB=False, axis=1,
                                      bandwidth=0.5e6) if nov * 2 == nfft else None
    finally:
        _lib.check(_lib.lib.iqw_debug_set_stft_variant(0))
    assert (np.abs(y - ref) / _tol.complex_tol(ref)).max() <= 1.0
    assert _tol.power_err_units(p, pref) <= 1.0
    if pb is not None:
        lo, hi = nfft // 4, nfft - nfft // 4
        want = pref.max(axis=1)[:, lo:hi]
        got = pb.cpu().numpy()[:, 0]
        assert got.shape == want.shape
        assert np.all(np.abs(got - want) <= _tol.POWER_RTOL * want + _tol.POWER_FLOOR * pref.max())


# ---------------------------------------------------------------------------------------------
# reducible statistics fused into kernel 1 (no spectrogram in memory): fourier.py:1322-1325
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,nfft,ovl,stats,kw', [
    (1 << 18, 1024, 0.5, ['mean', 'max'], {}),
    (1 << 19, 4096, 0.5, ['max', 'min', 'mean', 'rms', 'peak'], dict(bandwidth=0.5e6)),
    (1 << 16, 256, 0.75, ['mean', 'max', 'min'], dict(dB=False)),
    (1 << 17, 64, 0.0, ['min'], {}),
    (1 << 18, 8192, 0.5, ['mean', 'peak'], {}),
    (1 << 17, 512, 0.5, ['max'], dict(fractional_window=0.75)),
])
def test_fused_reducible_statistics_vs_oracle(cuda_device, n, nfft, ovl, stats, kw):
    from iqwaveform_b200 import fourier as F
    assert F._reducible(stats, nfft)
    x = synth(n % 83, (2, n))
    args = dict(fs=1e6, window='hann', resolution=1e6 / nfft, fractional_overlap=ovl,
                statistics=stats, axis=1, **kw)
    ref = orc.persistence_spectrum(x, **args)
    from iqwaveform_b200 import _lib
    _lib.profile(True)
    got = iqw.persistence_spectrum(dev_of(x, cuda_device), **args)
    torch.cuda.synchronize()
    names = set(_lib.profile_report())
    _lib.profile(False)
    assert 'stft_reduce_kernel' in names and 'stft_kernel' not in names      # nothing was materialised
    assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape
    nz = round((1 - kw.get('fractional_window', 1)) * nfft)
    _, _, p = orc.spectrogram(x, fs=1e6, window='hann', nperseg=nfft, noverlap=round(ovl * nfft), nzero=nz, axis=1)
    _check_persistence(got.cpu().numpy(), ref, stats, x, nfft, kw.get('dB', True), p.max(axis=(1, 2))[:, None])
    # same rows as the materialising path (statistics kernel 2) up to the mean's summation order, when both run on
    # the same FFT geometry (variant 1: the fused epilogue of the three-pass kernel, bitwise max / min).  The default
    # fused path at nfft 1024-4096 is the warp-specialised two-pass kernel: within the dB tolerance of the same rows.
    try:
        _lib.check(_lib.lib.iqw_debug_set_stft_variant(1))
        fused1 = iqw.persistence_spectrum(dev_of(x, cuda_device), **args)
        mixed = iqw.persistence_spectrum(dev_of(x, cuda_device), **dict(args, statistics=stats + [0.5]))[:, :len(stats)]
    finally:
        _lib.check(_lib.lib.iqw_debug_set_stft_variant(0))
    for i, s in enumerate(stats):
        if s in ('mean', 'rms'):
            assert torch.allclose(fused1[:, i], mixed[:, i], rtol=1e-5, atol=1e-4)
        else:
            assert torch.equal(fused1[:, i], mixed[:, i]), s
    _check_persistence(fused1.cpu().numpy(), ref, stats, x, nfft, kw.get('dB', True), p.max(axis=(1, 2))[:, None])


def test_fused_reducible_statistics_host_capture_and_1d(cuda_device):
    """a host capture large enough for the chunked copy path: every chunk is reduced as it lands and the
    partial rows are combined; equals the device-resident call"""
    n, nfft = 1 << 24, 2048         # 128 MB >= STREAM_MIN_BYTES
    x = synth(12, (n,))
    args = dict(fs=1e6, window='blackmanharris', resolution=1e6 / nfft, fractional_overlap=0.5,
                statistics=['max', 'mean', 'min'], dB=True, axis=0)
    host = iqw.persistence_spectrum(x, **args)
    assert isinstance(host, np.ndarray) and host.shape == (3, nfft)
    devr = iqw.persistence_spectrum(dev_of(x, cuda_device), **args).cpu().numpy()
    assert np.array_equal(host[0], devr[0]) and np.array_equal(host[2], devr[2])
    np.testing.assert_allclose(host[1], devr[1], atol=2e-4)


# ---------------------------------------------------------------------------------------------
# frame lengths that are not a power of two (fourier.py:1250-1255: nfft = round(fs / resolution))
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('nfft', [6, 100, 1000, 1536, 3000, 4094, 10000])
@pytest.mark.parametrize('overlap', [0.0, 0.5])
def test_any_even_frame_length_vs_oracle(cuda_device, nfft, overlap):
    nov = int(nfft * overlap)
    x = synth(nfft + 3, (2, nfft * 11 + 7))
    for norm in ('power', None):
        f, t, y = orc.stft(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm=norm)
        f2, t2, y2 = iqw.stft(dev_of(x, cuda_device), fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm=norm)
        assert np.array_equal(f, f2) and np.array_equal(t, t2) and tuple(y2.shape) == y.shape
        err = np.abs(y2.cpu().numpy() - y) / _tol.complex_tol(y)
        assert err.max() <= 1.0, (norm, err.max())
    _, _, ref = orc.spectrogram(x, fs=1e6, window='blackmanharris', nperseg=nfft, noverlap=nov, axis=1)
    got = iqw.spectrogram(dev_of(x, cuda_device), fs=1e6, window='blackmanharris', nperseg=nfft, noverlap=nov, axis=1,
                          return_axis_arrays=False).cpu().numpy()
    assert _tol.power_err_units(got, ref) <= 1.0
    truth = orc.stft_power_f64(x, window='blackmanharris', nperseg=nfft, noverlap=nov)
    e_gpu, e_ref = _tol.power_err_units(got, truth), _tol.power_err_units(ref, truth)
    assert e_gpu <= max(2.0 * e_ref, 0.5), (e_gpu, e_ref)
    dgot = iqw.spectrogram(dev_of(x, cuda_device), fs=1e6, window='blackmanharris', nperseg=nfft, noverlap=nov, axis=1,
                           return_axis_arrays=False, dB=True).cpu().numpy()
    dref = orc.powtodB(ref.copy())
    assert np.all(np.abs(dgot - dref) <= _tol.db_tol(dref, ref.max(axis=-1, keepdims=True)))


@pytest.mark.parametrize('fs,resolution,stats,kw', [
    (1e6, 1e3, [0.1, 0.5, 0.99, 'mean', 'max'], {}),                      # nfft 1000
    (100e6, 10e3, [0.5, 'max'], dict(bandwidth=40e6)),                      # nfft 10000, band trim
    (15.36e6, 10e3, ['median', 'min', 0.9], dict(dB=False)),                # nfft 1536
])
def test_persistence_any_frame_length_vs_oracle(cuda_device, fs, resolution, stats, kw):
    nfft = round(fs / resolution)
    x = synth(nfft % 71, (2, nfft * 150))
    args = dict(fs=fs, window='hann', resolution=resolution, fractional_overlap=0.5, statistics=stats, axis=1, **kw)
    ref = orc.persistence_spectrum(x, **args)
    got = iqw.persistence_spectrum(dev_of(x, cuda_device), **args)
    assert tuple(got.shape) == ref.shape
    _, _, p = orc.spectrogram(x, fs=fs, window='hann', nperseg=nfft, noverlap=nfft // 2, axis=1)
    _check_persistence(got.cpu().numpy(), ref, stats, x, nfft, kw.get('dB', True), p.max(axis=(1, 2))[:, None])


@pytest.mark.parametrize('nfft,nov', [(32768, 16384), (65536, 49152), (8192, 4096), (16384, 0)])
def test_cluster_kernel_on_request_vs_oracle(cuda_device, nfft, nov):
    """variant 3 at nfft 32768 / 65536: one frame per thread-block cluster of 2 / 4 CTAs, samples staged by bulk
    copies, exchanges through an L2-resident scratch ordered by barrier.cluster; variant 1 at 8192 / 16384: the
    older kernels those sizes used before.  Same transform as the default path."""
    from iqwaveform_b200 import _lib
    x = synth(nfft % 97, (2, nfft * 7 + 11))
    _, _, ref = orc.stft(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm='power')
    _, _, pref = orc.spectrogram(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1)
    try:
        _lib.check(_lib.lib.iqw_debug_set_stft_variant(3 if nfft >= 32768 else 1))
        y = iqw.stft(dev_of(x, cuda_device), fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1,
                     norm='power', return_axis_arrays=False).cpu().numpy()
        p = iqw.spectrogram(dev_of(x, cuda_device), fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1,
                            return_axis_arrays=False, dB=True).cpu().numpy()
    finally:
        _lib.check(_lib.lib.iqw_debug_set_stft_variant(0))
    assert (np.abs(y - ref) / _tol.complex_tol(ref)).max() <= 1.0
    dref = orc.powtodB(pref.copy())
    assert np.all(np.abs(p - dref) <= _tol.db_tol(dref, pref.max(axis=-1, keepdims=True)))


@pytest.mark.parametrize('nfft', [1024, 2048, 4096, 8192, 16384])
def test_few_frames_and_many_channels(cuda_device, nfft):
    """edge shapes of the slot-per-frame kernels: a single frame, fewer frames than frame slots, more channels than
    frames, a frame range that ends in the middle of a CTA's slots, band trim + dB -- against the oracle"""
    for C, T, ovl in [(1, 1, 0.5), (5, 1, 0.0), (3, 2, 0.5), (2, 7, 0.75), (7, 3, 0.5)]:
        nov = int(nfft * ovl)
        hop = nfft - nov
        n = (T - 1) * hop + nfft + (hop // 3 if T > 1 else 0)       # ragged tail shorter than a hop
        x = synth(C * 31 + T, (C, n))
        _, _, ref = orc.stft(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm='power')
        assert ref.shape == (C, T, nfft)
        y = iqw.stft(dev_of(x, cuda_device), fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm='power',
                     return_axis_arrays=False).cpu().numpy()
        assert (np.abs(y - ref) / _tol.complex_tol(ref)).max() <= 1.0, (C, T, ovl)
        kw = dict(fs=1e6, window='hann', resolution=1e6 / nfft, fractional_overlap=ovl, statistics=['max', 'min', 0.5],
                  dB=True, axis=1, bandwidth=0.25e6)
        got = iqw.persistence_spectrum(dev_of(x, cuda_device), **kw).cpu().numpy()
        want = orc.persistence_spectrum(x, **kw)
        _, _, p = orc.spectrogram(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1)
        assert got.shape == want.shape
        _check_persistence(got, want, ['max', 'min', 0.5], x, nfft, True, p.max(axis=(1, 2))[:, None])


def test_time_axis_first_layouts_use_the_library_transpose(cuda_device):
    """(N, C) axis=0 and (B, N, C) axis=1 captures: the tiled transpose kernel brings them to (channels, time);
    same results as the (C, N) axis=1 call"""
    from iqwaveform_b200 import _arrays
    x = synth(4, (3, 40000))
    xd = dev_of(x, cuda_device)
    t2, lead, trail = _arrays.as_channels(xd.T.contiguous(), 0)
    assert torch.equal(t2, xd) and lead == () and trail == (3,)
    x3 = torch.stack([xd.T.contiguous(), 2 * xd.T.contiguous()])          # (2, N, 3)
    t3, lead, trail = _arrays.as_channels(x3, 1)
    assert t3.shape == (6, 40000) and torch.equal(t3[:3], xd) and torch.equal(t3[3:], 2 * xd)
    want = iqw.iq_to_bin_power(xd, 1e-6, 1e-4, kind='mean', axis=1)
    got = iqw.iq_to_bin_power(xd.T.contiguous(), 1e-6, 1e-4, kind='mean', axis=0)
    assert torch.equal(got, want.T)
    _, _, w0 = orc.stft(x, fs=1.0, window='hann', nperseg=256, noverlap=0, axis=1, norm='power')
    y = iqw.stft(xd.T.contiguous(), fs=1.0, window='hann', nperseg=256, noverlap=0, axis=0, norm='power',
                 return_axis_arrays=False)
    np.testing.assert_allclose(y.cpu().numpy(), np.moveaxis(w0, 0, -1), atol=2e-7 * np.abs(w0).max())
