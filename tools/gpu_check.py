"""developer script: quick parity + timing sweep on a GPU box (not part of the test-suite)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iqwaveform_b200 as iqw
from oracle import iqw_oracle as orc
from oracle.make_golden import synth

torch.cuda.init()
dev = torch.device('cuda:0')
print(torch.cuda.get_device_name(0), 'SMs', torch.cuda.get_device_properties(0).multi_processor_count)


def relerr_power(p, ref):
    fmax = ref.max(axis=-1, keepdims=True)
    return np.max(np.abs(p - ref) / (1e-5 * np.abs(ref) + 2e-7 * fmax))


ok = True
for nfft in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
    for ov in (0, 0.5, 0.75):
        nov = int(nfft * ov)
        x = synth(nfft, (2, nfft * 40 + 17))
        f, t, y = orc.stft(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm='power')
        f2, t2, y2 = iqw.stft(torch.from_numpy(x).to(dev), fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1, norm='power')
        y2 = y2.cpu().numpy()
        assert y2.shape == y.shape, (y2.shape, y.shape)
        scale = np.abs(y).max()
        e = np.abs(y2 - y).max() / scale
        p64 = orc.stft_power_f64(x, window='hann', nperseg=nfft, noverlap=nov)
        _, _, p = orc.spectrogram(x, fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1)
        _, _, p2 = iqw.spectrogram(torch.from_numpy(x).to(dev), fs=1e6, window='hann', nperseg=nfft, noverlap=nov, axis=1)
        p2 = p2.cpu().numpy()
        e_ref = relerr_power(p, p64); e_gpu = relerr_power(p2, p64); e_gr = relerr_power(p2, p)
        flag = '' if (e < 2e-6 and e_gr < 1) else '  <<<<<< FAIL'
        if flag: ok = False
        print(f'nfft={nfft:5d} ov={ov:4.2f} stft max|d|/max={e:.2e}  power tol-units: ref-vs-f64={e_ref:.2f} gpu-vs-f64={e_gpu:.2f} gpu-vs-ref={e_gr:.2f}{flag}')

# dB
x = synth(5, (1, 200000))
_, _, p = orc.spectrogram(x, fs=1e6, window='blackmanharris', nperseg=2048, noverlap=1024, axis=1)
d = orc.powtodB(p.copy())
_, _, d2 = iqw.spectrogram(torch.from_numpy(x).to(dev), fs=1e6, window='blackmanharris', nperseg=2048, noverlap=1024, axis=1, dB=True)
print('dB max abs diff', np.abs(d2.cpu().numpy() - d).max())

# persistence
for (n, nfft, ovf, stats, kw) in [
    (1 << 18, 1024, 0.5, [0.5, 0.99, 'mean', 'max'], {}),
    (1 << 20, 1024, 0.5, [0.1, 0.5, 0.9, 0.999, 'min', 'median'], {}),
    (1 << 19, 4096, 0.5, [0.1, 0.5, 0.9, 0.999], dict(bandwidth=0.5e6)),
    (1 << 16, 256, 0.75, ['mean', 0.25, 'max', 1.0, 0.0], dict(dB=False)),
]:
    x = synth(n % 97, (2, n))
    ref = orc.persistence_spectrum(x, fs=1e6, window='hann', resolution=1e6 / nfft, fractional_overlap=ovf, statistics=stats, axis=1, **kw)
    got = iqw.persistence_spectrum(torch.from_numpy(x).to(dev), fs=1e6, window='hann', resolution=1e6 / nfft, fractional_overlap=ovf, statistics=stats, axis=1, **kw).cpu().numpy()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if kw.get('dB', True):
        err = np.abs(got - ref).max(axis=(0, 2))
    else:
        err = (np.abs(got - ref) / np.abs(ref)).max(axis=(0, 2))
    print(f'psd n={n} nfft={nfft} stats={stats}: per-stat max err', np.array2string(err, precision=2))

# exact order statistics on a synthetic matrix (no FFT): compare to np.quantile bitwise
rng = np.random.default_rng(0)
for (T, nb) in [(1, 5), (2, 130), (77, 64), (5000, 300), (70001, 257), (300000, 128)]:
    a = rng.standard_normal((2, T, nb)).astype(np.float32)
    a[0, :, 0] = 1.5          # constant column
    a[0, :, 1] = np.round(a[0, :, 1])   # heavy ties
    if T > 10: a[1, 3, 2] = 1e30; a[1, 4, 2] = -1e30   # outliers
    qs = [0.0, 0.1, 0.5, 0.999, 1.0]
    stats = qs + ['median', 'min', 'max', 'mean']
    got = iqw.time_statistics(torch.from_numpy(a).to(dev), stats, dB=False).cpu().numpy()
    ref_q = np.quantile(a, np.array(qs, dtype=np.float32), axis=1)
    exact = all(np.array_equal(got[:, i], ref_q[i]) for i in range(len(qs)))
    med = np.array_equal(got[:, 5], np.median(a, axis=1))
    mn = np.array_equal(got[:, 6], a.min(axis=1)); mx = np.array_equal(got[:, 7], a.max(axis=1))
    me = np.abs(got[:, 8] - a.astype(np.float64).mean(axis=1)).max()
    flag = '' if (exact and med and mn and mx) else '  <<<<<< FAIL'
    if flag: ok = False
    print(f'time_stats T={T} nb={nb}: quantiles bitwise={exact} median={med} min={mn} max={mx} mean err={me:.2e}{flag}')

# bin power
x = synth(3, (3, 300000))
for kind in ('mean', 'max', 'min', 'median', 0.25):
    for nbin in (100, 1536, 50000, 150000):
        ref = orc.iq_to_bin_power(x, 1.0, float(nbin), kind=kind, axis=1, truncate=True)
        got = iqw.iq_to_bin_power(torch.from_numpy(x).to(dev), 1.0, float(nbin), kind=kind, axis=1, truncate=True).cpu().numpy()
        assert got.shape == ref.shape
        e = np.abs(got - ref).max() / np.abs(ref).max()
        flag = '' if e < 2e-6 else '  <<<<<< FAIL'
        if flag: ok = False
        print(f'bin_power kind={kind} bin={nbin}: rel err {e:.2e}{flag}')
print('ALL OK' if ok else 'SOME FAILED')

# ---- timing ----
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]

N = 1 << 28
g = torch.Generator(device=dev).manual_seed(1)
xr = torch.randn(2 * N, generator=g, device=dev, dtype=torch.float32)
xb = torch.view_as_complex(xr.view(N, 2)).view(1, N)
for nfft in (64, 256, 1024, 2048, 4096, 8192):
    for ov in (0.5, 0.75):
        best, med = timeit(lambda: iqw.spectrogram(xb, fs=1e8, window='hann', nperseg=nfft, noverlap=int(nfft * ov), axis=1, return_axis_arrays=False))
        r = 1 / (1 - ov)
        print(f'spectrogram nfft={nfft} ov={ov}: {best:.2f} ms  {N / best / 1e6:.1f} GS/s  {(8 + 4 * r) * N / best / 1e6:.0f} GB/s algorithmic')
best, med = timeit(lambda: iqw.iq_to_bin_power(xb, 1.0, 245760.0, kind='mean', axis=1, truncate=True))
print(f'bin_power mean: {best:.2f} ms {N / best / 1e6:.1f} GS/s {8 * N / best / 1e6:.0f} GB/s')
best, med = timeit(lambda: iqw.persistence_spectrum(xb, fs=1e8, window='hann', resolution=1e8 / 4096, fractional_overlap=0.5, statistics=[0.1, 0.5, 0.9, 0.999], axis=1), n=3)
print(f'persistence nfft=4096 4 quantiles: {best:.2f} ms {N / best / 1e6:.1f} GS/s')
p = iqw.spectrogram(xb, fs=1e8, window='hann', nperseg=4096, noverlap=2048, axis=1, return_axis_arrays=False)
best, med = timeit(lambda: iqw.time_statistics(p, [0.1, 0.5, 0.9, 0.999], dB=True), n=3)
print(f'time_statistics alone on {tuple(p.shape)}: {best:.2f} ms')
