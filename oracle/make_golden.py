"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden

Every fixture stores the exact input bits, the call parameters (as a JSON string) and what the
reference returned.  For ``power_spectral_density`` the quantile rows the reference returns are
uninitialised memory, so the fixture stores the named-statistic rows from the
reference's return value and, separately, the quantile rows evaluated with the reference's own
building blocks: ``np.quantile(powtodB(spectrogram(x)[band], eps=1e-25), float32(q), axis=1)``.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def synth(seed: int, shape, tones=((0.0651, 0.5), (-0.2148, 0.05), (0.3256, 3.0))) -> np.ndarray:
    """unit-variance complex noise plus off-bin tones, complex64"""
    rng = np.random.default_rng(seed)
    n = shape[-1]
    x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)
    k = np.arange(n)
    for f, a in tones:
        x = x + a * np.exp(2j * np.pi * f * k)
    return x.astype(np.complex64)


def _save(name, params, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + '.npz')
    np.savez_compressed(path, params=json.dumps(params), **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB')


def main():
    ref = ref_shim.load()
    if ref is None:
        raise SystemExit('reference not present; golden fixtures can only be made in the build container')
    fourier, pa = ref.fourier, ref.power_analysis

    # --- stft, complex output ------------------------------------------------------------
    x = synth(11, (2, 6000))
    for tag, kw in (
        ('stft_hann_256_128_power', dict(window='hann', nperseg=256, noverlap=128, norm='power')),
        ('stft_hann_256_128_cola', dict(window='hann', nperseg=256, noverlap=128, norm=None)),
        ('stft_bh_512_0_power', dict(window='blackmanharris', nperseg=512, noverlap=0, norm='power')),
        ('stft_kaiser_64_48_power', dict(window=('kaiser', 8.0), nperseg=64, noverlap=48, norm='power')),
        ('stft_hann_1024_512_nzero', dict(window='hann', nperseg=1024, noverlap=512, nzero=256, norm='power')),
    ):
        f, t, y = fourier.stft(x.copy(), fs=15.36e6, axis=1, **kw)
        p = dict(kw, fs=15.36e6, axis=1)
        if isinstance(p['window'], tuple):
            p['window'] = list(p['window'])
        _save(tag, p, x=x, freqs=f, times=t, y=y)

    # --- spectrogram ---------------------------------------------------------------------
    x = synth(12, (1, 20000))
    for tag, kw in (
        ('spg_bh_2048_1024', dict(window='blackmanharris', nperseg=2048, noverlap=1024)),
        ('spg_hann_1024_768', dict(window='hann', nperseg=1024, noverlap=768)),
        ('spg_rect_4096_0', dict(window='rect', nperseg=4096, noverlap=0)),
    ):
        f, t, p = fourier.spectrogram(x.copy(), fs=100e6, axis=1, **kw)
        _save(tag, dict(kw, fs=100e6, axis=1), x=x, freqs=f, times=t, power=p,
              dB=pa.powtodB(p.copy()))

    # --- persistence spectrum ------------------------------------------------------------
    x = synth(13, (2, 1 << 16))
    _save('psd_input', dict(seed=13), x=x)
    for tag, kw in (
        ('psd_hann_1024_half', dict(fs=15.36e6, window='hann', resolution=15e3,
                                    fractional_overlap=0.5, statistics=[0.5, 0.99, 'mean', 'max'])),
        ('psd_hann_4096_trim', dict(fs=100e6, window='hann', resolution=100e6 / 4096,
                                    fractional_overlap=0.5, bandwidth=50e6,
                                    statistics=[0.1, 'min', 0.5, 0.9, 'median', 0.999, 'peak'])),
        ('psd_bh_256_linear', dict(fs=1e6, window='blackmanharris', resolution=1e6 / 256,
                                   fractional_overlap=0.75, dB=False,
                                   statistics=['mean', 0.25, 'max', 1.0, 0.0])),
    ):
        ret = fourier.power_spectral_density(x.copy(), axis=1, **kw)
        nfft = round(kw['fs'] / kw['resolution'])
        nov = round(kw['fractional_overlap'] * nfft)
        _, _, spg = fourier.spectrogram(x.copy(), fs=kw['fs'], window=kw['window'], nperseg=nfft,
                                        noverlap=nov, axis=1)
        bw = kw.get('bandwidth', float('inf'))
        if bw != float('inf'):
            ilo, ihi = fourier._freq_band_edges(nfft, 1.0 / kw['fs'], -bw / 2, bw / 2)
            spg = spg[..., ilo:ihi]
        if kw.get('dB', True):
            spg = pa.powtodB(spg, eps=1e-25, out=spg)
        isq = [not isinstance(s, str) for s in kw['statistics']]
        q = np.array([s for s in kw['statistics'] if not isinstance(s, str)], dtype=np.float32)
        qrows = np.quantile(spg, q, axis=1)                       # (nq, C, nbins)
        named = ret[:, [i for i, f in enumerate(isq) if not f], :]
        p = dict(kw, axis=1)
        if p.get('bandwidth') == float('inf'):
            p.pop('bandwidth')
        _save(tag, dict(p, input='psd_input'), named_rows=named, quantile_rows=np.moveaxis(qrows, 0, 1),
              is_quantile=np.array(isq))

    # --- bin power -----------------------------------------------------------------------
    x = synth(14, (3, 30000))
    out = {}
    for kind in ('mean', 'max', 'min', 'median', 'rms', 'peak'):
        out[kind] = pa.iq_to_bin_power(x, 1 / 15.36e6, 1536 / 15.36e6, kind=kind, axis=1,
                                       truncate=True)
    out['q25'] = pa.iq_to_bin_power(x, 1 / 15.36e6, 1536 / 15.36e6, kind=0.25, axis=1,
                                    truncate=True)
    _save('binpower_1536', dict(Ts=1 / 15.36e6, Tbin=1536 / 15.36e6, axis=1, truncate=True),
          x=x, **out)


def make_inverse():
    """istft / ola_filter fixtures; `python -m oracle.make_golden inverse`"""
    ref = ref_shim.load()
    if ref is None:
        raise SystemExit('reference not present; golden fixtures can only be made in the build container')
    fourier = ref.fourier
    x = synth(15, (2, 8192))
    for tag, kw, size in (
        ('istft_hamming_256_128', dict(window='hamming', nperseg=256, noverlap=128), 8192),
        ('istft_bh_1024_768', dict(window='blackmanharris', nperseg=1024, noverlap=768), None),
        ('istft_rect_64_0', dict(window='rect', nperseg=64, noverlap=0), 8001),
    ):
        _, _, y = fourier.stft(x.copy(), fs=1e6, axis=1, truncate=False, **kw)
        xr = fourier.istft(y.copy(), size, nfft=kw['nperseg'], noverlap=kw['noverlap'], axis=1)
        _save(tag, dict(kw, fs=1e6, axis=1, size=size), y=y, x=np.ascontiguousarray(xr))
    x = synth(16, (2, 16384))
    for tag, kw in (
        ('ola_hamming_512_all', dict(fs=1e6, nfft=512, window='hamming', passband=[-2e5, 2e5])),
        # the reference's passband arithmetic only reacts to tiny values (oracle.ola_passband_bins)
        ('ola_hamming_512_band', dict(fs=1e6, nfft=512, window='hamming',
                                      passband=[-1.3647444248199463 - 2e-6, 1.3647444248199463 + 1e-6])),   # bins 129..316
    ):
        out = fourier.ola_filter(x.copy(), axis=1, **dict(kw, passband=tuple(kw['passband'])))
        _save(tag, dict(kw, axis=1), x=x, out=np.ascontiguousarray(out))


def make_resample():
    """oaresample fixtures; `python -m oracle.make_golden resample`"""
    ref = ref_shim.load()
    if ref is None:
        raise SystemExit('reference not present; golden fixtures can only be made in the build container')
    x = synth(18, (2, 8192))
    for tag, kw in (
        ('oares_down_1024_512', dict(up=512, down=1024)),
        ('oares_up_512_1024_fir', dict(up=1024, down=512, filter_bandwidth=0.4e6, transition_bandwidth=100e3, scale=0.5)),
        ('oares_shift_1024_256', dict(up=256, down=1024, frequency_shift=1e6 / 1024 * 100)),
    ):
        out = ref.fourier.oaresample(x.copy(), fs=1e6, axis=1, window='hamming', **kw)
        _save(tag, dict(kw, fs=1e6, axis=1, window='hamming'), x=x, out=np.ascontiguousarray(out))


def make_consumers():
    """sample_ccdf / histogram_last_axis fixtures; `python -m oracle.make_golden consumers`"""
    ref = ref_shim.load()
    if ref is None:
        raise SystemExit('reference not present; golden fixtures can only be made in the build container')
    x = synth(17, (30000,))
    p = ref.power_analysis.envtopow(x)
    edges = np.linspace(0.0, 30.0, 61)
    _save('ccdf_power_61', dict(), p=p, edges=edges,
          density=ref.power_analysis.sample_ccdf(p, edges, density=True),
          counts=ref.power_analysis.sample_ccdf(p, edges, density=False))
    pdb = ref.power_analysis.powtodB(p.reshape(6, 5000).copy())
    h, e = ref.util.histogram_last_axis(pdb, 50, (-30.0, 20.0))
    _save('hist_db_50', dict(bins=50, range=[-30.0, 20.0]), x=pdb, hist=h, edges=e)


if __name__ == '__main__':
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == 'resample':
        make_resample()
        raise SystemExit
    if len(sys.argv) > 1 and sys.argv[1] == 'consumers':
        make_consumers()
        raise SystemExit
    if len(sys.argv) > 1 and sys.argv[1] == 'inverse':
        make_inverse()
    else:
        main()
        make_inverse()
        make_resample()
        make_consumers()
