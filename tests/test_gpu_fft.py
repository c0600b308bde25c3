"""row A4's public pair `fft` / `ifft` (fourier.py:200-246) and `zero_stft_by_freq` (707-720) on the
device: kernel 1 with an all-ones window / kernel 4 without the shift sign, against the oracle
(= the reference's own scipy.fft calls)."""
import numpy as np
import pytest
import torch

import _tol
import iqwaveform_b200 as iqw
from oracle import iqw_oracle as orc
from oracle.make_golden import synth

pytestmark = pytest.mark.gpu


def _close(got, want, tol):
    got = np.asarray(got)
    assert got.shape == want.shape and got.dtype == np.complex64
    assert np.all(np.abs(got - want) <= tol), float(np.max(np.abs(got - want) / tol))


@pytest.mark.parametrize('n', [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_fft_matches_oracle(n):
    rows = 37 if n <= 8192 else 5
    x = synth(n, (rows, n))
    want = orc.fft(x, axis=1)
    got = iqw.fft(torch.from_numpy(x).cuda(), axis=1)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    _close(got.cpu().numpy(), want, _tol.complex_tol(want))


@pytest.mark.parametrize('n', [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_ifft_matches_oracle_and_inverts_fft(n):
    x = synth(n + 1, (29, n))
    want = orc.ifft(x, axis=1)
    xd = torch.from_numpy(x).cuda()
    got = iqw.ifft(xd, axis=1)
    _close(got.cpu().numpy(), want, _tol.waveform_tol(want))
    back = iqw.ifft(iqw.fft(xd))
    _close(back.cpu().numpy(), x, 2 * _tol.waveform_tol(x))


def test_fft_is_the_unshifted_stft_of_a_rect_window():
    """bin k of `fft` is bin k + nfft/2 (mod nfft) of `stft(window='rect', norm='power')` times nfft:
    the two public faces of kernel 1 agree (to rounding: the (-1)^n folded into the stft window
    makes it a different butterfly sequence for the same bin)"""
    n = 1024
    x = synth(3, (8 * n,))
    xd = torch.from_numpy(x).cuda()
    y = iqw.stft(xd, fs=1.0, window='rect', nperseg=n, noverlap=0, norm='power', return_axis_arrays=False)
    f = iqw.fft(xd.reshape(8, n), axis=1)
    want = (torch.roll(f, n // 2, dims=1) / n).cpu().numpy()
    _close(y.cpu().numpy(), want, _tol.complex_tol(want))


@pytest.mark.parametrize('shape,axis', [((64,), 0), ((64,), -1), ((3, 128), 1), ((256, 5), 0), ((2, 512, 3), 1),
                                        ((2, 3, 32), -1)])
def test_layouts_numpy_in_numpy_out(shape, axis):
    x = synth(11, shape)
    got = iqw.fft(x, axis=axis)
    assert isinstance(got, np.ndarray)
    want = orc.fft(x, axis=axis)
    _close(got, want, np.moveaxis(_tol.complex_tol(np.moveaxis(want, axis, -1)), -1, axis))
    goti = iqw.ifft(x, axis=axis)
    wanti = orc.ifft(x, axis=axis)
    _close(goti, wanti, np.moveaxis(_tol.waveform_tol(np.moveaxis(wanti, axis, -1)), -1, axis))


def test_out_argument():
    x = torch.from_numpy(synth(13, (6, 256))).cuda()
    out = torch.empty_like(x)
    r = iqw.fft(x, axis=1, out=out)
    assert r is out and torch.equal(out, iqw.fft(x, axis=1))
    r = iqw.ifft(x, axis=1, out=out)
    assert r is out and torch.equal(out, iqw.ifft(x, axis=1))


def test_unsupported_sizes_and_types_raise():
    with pytest.raises(NotImplementedError):
        iqw.fft(torch.zeros(4, 100, dtype=torch.complex64, device='cuda'))
    with pytest.raises(NotImplementedError):
        iqw.ifft(torch.zeros(2, 16384, dtype=torch.complex64, device='cuda'))
    with pytest.raises(NotImplementedError):
        iqw.fft(torch.zeros(4, 64, dtype=torch.complex128, device='cuda'))
    with pytest.raises(TypeError):
        iqw.fft([1, 2, 3])


@pytest.mark.parametrize('passband', [(-0.2e6, 0.1e6), (-3.0, 2.0), (-1e9, 1e9), (None, 2.0), (-3.0, None)])
def test_zero_stft_by_freq_bit_exact(passband):
    x = synth(5, (2, 9000))
    f, _, y = orc.stft(x, fs=1e6, window='hamming', nperseg=256, noverlap=128, axis=1)
    want = orc.zero_stft_by_freq(f, y.copy(), passband=passband, axis=1)
    yd = torch.from_numpy(y).cuda()
    got = iqw.zero_stft_by_freq(f, yd, passband=passband, axis=1)
    assert got is yd
    assert np.array_equal(got.cpu().numpy().view(np.float32), want.view(np.float32))
    with pytest.raises(TypeError):
        iqw.zero_stft_by_freq(f, y, passband=passband, axis=1)


@pytest.mark.parametrize('n,rows', [(16, 1), (64, 7), (1024, 3), (4096, 5), (8192, 2)])
def test_no_write_outside_the_output(n, rows):
    """compute-sanitizer is not available on the GPU pool: the outputs sit between guard bands"""
    x = torch.from_numpy(synth(n + rows, (rows, n))).cuda()
    guard = 4096
    for fn in (iqw.fft, iqw.ifft):
        buf = torch.full((rows * n + 2 * guard,), complex(123.0, -321.0), dtype=torch.complex64, device='cuda')
        out = buf[guard:guard + rows * n].view(rows, n)
        assert out.is_contiguous()
        fn(x, axis=1, out=out)
        torch.cuda.synchronize()
        assert torch.all(buf[:guard] == complex(123.0, -321.0)) and torch.all(buf[-guard:] == complex(123.0, -321.0))
        assert torch.equal(out, fn(x, axis=1))
