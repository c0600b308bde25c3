"""GPU parity of the elementwise power transforms (power_analysis.py:168-338) against the numpy
oracle (the reference's generic branch) on seeded inputs."""
import numpy as np
import pytest
import torch

import iqwaveform_b200 as iqw
from oracle import iqw_oracle as orc

pytestmark = pytest.mark.gpu


def _dev(x, d):
    return torch.from_numpy(x).to(d)


@pytest.mark.parametrize('n', [1, 3, 1000, 4097, 1 << 20])
def test_powtodB_dBtopow_envtopow_envtodB(cuda_device, n):
    rng = np.random.default_rng(n)
    p = rng.exponential(1e-6, n).astype(np.float32)
    p[::7] = 0.0                                   # log of zero -> -inf, as in the reference
    got = iqw.powtodB(_dev(p, cuda_device)).cpu().numpy()
    want = orc.powtodB(p.copy())
    assert np.array_equal(np.isneginf(got), np.isneginf(want))
    f = np.isfinite(want)
    assert np.max(np.abs(got[f] - want[f]), initial=0) <= 5e-5          # dB
    got = iqw.powtodB(_dev(p, cuda_device), eps=1e-25).cpu().numpy()
    assert np.max(np.abs(got - orc.powtodB(p.copy(), eps=1e-25))) <= 5e-5
    # abs=False: negative arguments are NaN
    s = (p - 5e-7).astype(np.float32)
    got = iqw.powtodB(_dev(s, cuda_device), abs=False).cpu().numpy()
    want = orc.powtodB_noabs(s)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    f = np.isfinite(want)
    assert np.max(np.abs(got[f] - want[f]), initial=0) <= 5e-5

    d = rng.uniform(-150, 30, n).astype(np.float32)
    got = iqw.dBtopow(_dev(d, cuda_device)).cpu().numpy()
    np.testing.assert_allclose(got, orc.dBtopow(d), rtol=3e-6)

    z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    got = iqw.envtopow(_dev(z, cuda_device)).cpu().numpy()
    np.testing.assert_allclose(got, orc.envtopow(z), rtol=4e-7)         # re^2+im^2 vs hypot^2
    got = iqw.envtopow(_dev(d, cuda_device)).cpu().numpy()
    np.testing.assert_allclose(got, orc.envtopow(d), rtol=1e-7)
    got = iqw.envtodB(_dev(z, cuda_device)).cpu().numpy()
    assert np.max(np.abs(got - orc.envtodB(z))) <= 5e-5
    got = iqw.envtodB(_dev(z, cuda_device), eps=1e-3).cpu().numpy()
    assert np.max(np.abs(got - orc.envtodB(z, eps=1e-3))) <= 5e-5
    got = iqw.envtodB(_dev(np.abs(d) + 1, cuda_device), abs=False).cpu().numpy()
    assert np.max(np.abs(got - orc.envtodB(np.abs(d) + 1, abs=False))) <= 5e-5


def test_elementwise_kinds_and_scalars(cuda_device):
    p = np.array([1.0, 10.0, 100.0], np.float32)
    out = iqw.powtodB(p)                                   # numpy in -> numpy out
    assert isinstance(out, np.ndarray) and np.allclose(out, [0, 10, 20], atol=5e-5)
    assert iqw.powtodB(1) == 0 and iqw.powtodB(100.0) == 20.0          # reference tests/test_transforms.py
    assert iqw.dBtopow(20.0) == pytest.approx(100.0)
    assert iqw.envtodB(10.0) == pytest.approx(20.0)
    with pytest.raises(NotImplementedError):
        iqw.powtodB(torch.zeros(4, dtype=torch.float64, device=cuda_device))
    with pytest.raises(TypeError):
        iqw.powtodB(torch.zeros(4, dtype=torch.complex64, device=cuda_device))
    buf = torch.empty(3, dtype=torch.float32, device=cuda_device)
    r = iqw.envtopow(_dev(p, cuda_device), out=buf)
    assert r is buf and np.allclose(buf.cpu().numpy(), p * p)


def test_dBlinmean_dBlinsum(cuda_device):
    rng = np.random.default_rng(3)
    d = rng.uniform(-120, -60, (5000, 96)).astype(np.float32)
    lin = 10.0 ** (d.astype(np.float64) / 10)
    got = iqw.dBlinmean(_dev(d, cuda_device), axis=0).cpu().numpy()
    np.testing.assert_allclose(got, 10 * np.log10(lin.mean(axis=0)), atol=1e-4)
    got = iqw.dBlinsum(_dev(d, cuda_device), axis=0).cpu().numpy()
    np.testing.assert_allclose(got, 10 * np.log10(lin.sum(axis=0)), atol=1e-4)
    d3 = d.reshape(2, 2500, 96)
    got = iqw.dBlinmean(_dev(d3, cuda_device), axis=1).cpu().numpy()
    np.testing.assert_allclose(got, 10 * np.log10((10.0 ** (d3.astype(np.float64) / 10)).mean(axis=1)), atol=1e-4)
    got = iqw.dBlinmean(_dev(d[:, 0].copy(), cuda_device))
    assert got.shape == () and float(got) == pytest.approx(10 * np.log10(lin[:, 0].mean()), abs=1e-4)
    with pytest.raises(NotImplementedError):
        iqw.dBlinmean(_dev(d, cuda_device), axis=1)


def test_iq_to_cyclic_power(cuda_device):
    from oracle.make_golden import synth
    x = synth(12, (3, 60000))
    kw = dict(Ts=1e-6, detector_period=1e-5, cyclic_period=1e-3)
    want = orc.iq_to_cyclic_power(x, axis=1, **kw)
    got = iqw.iq_to_cyclic_power(_dev(x, cuda_device), axis=1, **kw)
    for d in ('rms', 'peak'):
        for k in ('min', 'mean', 'max'):
            g = got[d][k].cpu().numpy()
            assert g.shape == want[d][k].shape == (3, 100)
            np.testing.assert_allclose(g, want[d][k], rtol=3e-6)
    got1 = iqw.iq_to_cyclic_power(x[0], **kw)                 # 1-D numpy in -> numpy out
    assert isinstance(got1['rms']['mean'], np.ndarray)
    np.testing.assert_allclose(got1['peak']['max'], want['peak']['max'][0], rtol=3e-6)
    with pytest.raises(ValueError):
        iqw.iq_to_cyclic_power(_dev(x, cuda_device), axis=1, Ts=1e-6, detector_period=1e-5, cyclic_period=1.05e-4)
    with pytest.raises(ValueError):
        iqw.iq_to_cyclic_power(_dev(x, cuda_device), axis=1, Ts=1e-6, detector_period=1e-5, cyclic_period=7e-4)


def test_input_domains(cuda_device):
    """set_input_domain('frequency') / ('time_binned_power') branches"""
    from oracle.make_golden import synth
    x = synth(8, (2, 40000))
    _, _, X = orc.stft(x, fs=1e6, window='hann', nperseg=256, noverlap=128, axis=1, norm='power')
    stats = [0.1, 0.5, 'mean', 'max']
    for dB in (True, False):
        for bw in (float('inf'), 0.5e6):
            want = orc.persistence_spectrum_from_stft(X, fs=1e6, bandwidth=bw, resolution=1e6 / 256,
                                                      fractional_overlap=0.5, statistics=stats, dB=dB, axis=1)
            with iqw.set_input_domain('frequency'):
                got = iqw.power_spectral_density(_dev(X, cuda_device), fs=1e6, bandwidth=bw, window='hann',
                                                 resolution=1e6 / 256, fractional_overlap=0.5,
                                                 statistics=stats, dB=dB, axis=1).cpu().numpy()
            assert got.shape == want.shape
            if dB:
                assert np.max(np.abs(got - want)) <= 1e-4
            else:
                np.testing.assert_allclose(got, want, rtol=5e-6)
    assert iqw.get_input_domain() == iqw.Domain.TIME
    # binned power in, cyclic statistics out
    kw = dict(Ts=1e-6, detector_period=1e-5, cyclic_period=1e-3)
    binned = {d: orc.iq_to_bin_power(x, 1e-6, 1e-5, kind=d, axis=1) for d in ('rms', 'peak')}
    want = orc.iq_to_cyclic_power(x, axis=1, **kw)
    with iqw.set_input_domain(iqw.Domain.TIME_BINNED_POWER):
        got = iqw.iq_to_cyclic_power({d: _dev(v, cuda_device) for d, v in binned.items()}, axis=1,
                                     detectors=None, **kw)
        with pytest.raises(TypeError):
            iqw.iq_to_cyclic_power(_dev(x, cuda_device), axis=1, **kw)
    for d in ('rms', 'peak'):
        for k in ('min', 'mean', 'max'):
            np.testing.assert_allclose(got[d][k].cpu().numpy(), want[d][k], rtol=3e-6)


# ---------------------------------------------------------------------------------------------
# labelled containers (power_analysis.py:104-165) -- includes the reference's own tests
# (/root/reference/tests/test_transforms.py:1-17) run against the product
# ---------------------------------------------------------------------------------------------
def test_reference_test_transforms(cuda_device):
    import pandas as pd
    from iqwaveform_b200 import powtodB
    assert powtodB(1) == 0                      # test_transform_int
    assert powtodB(1.0) == 0                    # test_transform_float
    s = pd.Series([1, 10, 100])                 # test_transform_series
    ret = powtodB(s)
    assert isinstance(ret, pd.Series) and ret.index.equals(s.index)
    assert np.allclose(pd.Series([0, 10, 20]).values, ret.values)


def test_pandas_in_pandas_out(cuda_device):
    import pandas as pd
    from iqwaveform_b200 import dBtopow, envtodB, envtopow, powtodB
    rng = np.random.default_rng(0)
    idx = pd.Index(np.arange(50) * 0.25, name='t')
    df = pd.DataFrame(rng.random((50, 3)).astype(np.float32) + 0.1, index=idx, columns=['a', 'b', 'c'])
    out = powtodB(df)
    assert isinstance(out, pd.DataFrame) and out.index.equals(idx) and list(out.columns) == ['a', 'b', 'c']
    np.testing.assert_allclose(out.values, orc.powtodB(df.values.copy()), atol=5e-5)
    back = dBtopow(out)
    assert isinstance(back, pd.DataFrame)
    np.testing.assert_allclose(back.values, df.values, rtol=2e-5)
    z = pd.Series((rng.standard_normal(40) + 1j * rng.standard_normal(40)).astype(np.complex64))
    p = envtopow(z)
    assert isinstance(p, pd.Series) and p.dtype == np.float32
    np.testing.assert_allclose(p.values, np.abs(z.values) ** 2, rtol=1e-6)
    d = envtodB(z)
    np.testing.assert_allclose(d.values, 20 * np.log10(np.abs(z.values)), atol=5e-5)
    # float64 / integer host arrays: computed in float32, returned in the reference's dtype
    y = powtodB(np.array([1.0, 100.0]))
    assert y.dtype == np.float64 and np.allclose(y, [0, 20], atol=1e-5)


def test_unit_transforms_match_the_reference_strings():
    from iqwaveform_b200 import power_analysis as P
    assert P.unit_linear_to_dB('mW') == 'dBm' and P.unit_dB_to_linear('dBW/Hz') == 'W/Hz'
    assert P.unit_wave_to_dB('√mW') == 'dBm' and P.unit_wave_to_linear('√W') == 'W'
    assert P.unit_dB_to_wave('dBm') == '√mW'
