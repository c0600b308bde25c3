"""B200-native ``iq_to_bin_power`` with the reference signature
(/root/reference/src/iqwaveform/power_analysis.py:341-385)."""
from __future__ import annotations

import ctypes
from numbers import Number

import torch

from . import _arrays, _lib, _plan
from .fourier import _stream_ptr, time_statistics

__all__ = ['iq_to_bin_power']

_DIRECT = {'mean': 'mean', 'rms': 'mean', 'max': 'max', 'peak': 'max', 'min': 'min'}


def iq_to_bin_power(iq, Ts: float, Tbin: float, randomize: bool = False, kind='mean',
                    truncate=False, axis=0):
    """power of `iq` along `axis` detected on contiguous bins of duration Tbin.

    kind: 'mean' | 'rms' | 'max' | 'peak' | 'min' (one streaming kernel), 'median' or a float
    quantile (|x|^2 transposed, then the exact order-statistic kernel).  Returns float32 with the
    time axis replaced by the bin axis."""
    if truncate or _plan.isroundmod(Tbin, Ts):
        nb = round(Tbin / Ts)
    else:
        raise ValueError(
            f'bin period ({Tbin} s) must be multiple of waveform sample period ({Ts})')
    if randomize:
        raise NotImplementedError('randomize=True is not built (random gather, reference axis 0 only)')
    if callable(kind):
        raise NotImplementedError('callable detectors are not built (arbitrary python)')
    if isinstance(kind, str):
        if kind not in _DIRECT and kind != 'median':
            raise ValueError(f"kind argument must be one of {list(_DIRECT) + ['median']}")
    elif not isinstance(kind, Number):
        raise ValueError(f'invalid statistic ufunc "{kind}"')

    shape = getattr(iq, 'shape', None)
    if shape is None:
        raise TypeError('unrecognized object type')
    ax = axis + len(shape) if axis < 0 else axis
    if not 0 <= ax < len(shape):
        raise ValueError(f'axis {axis} exceeds the number of dimensions')
    if 0 in tuple(shape):
        raise IndexError('cannot form blocks on arrays of size 0')
    if nb < 1:
        raise ValueError('bin period shorter than one sample')
    if shape[ax] % nb and not truncate:
        raise ValueError(f'axis 0 size {shape[ax]} is not a factor of block size {nb}')

    xd, res = _arrays.to_device(iq)
    if xd.dtype != torch.complex64:
        raise NotImplementedError(f'only complex64 waveforms are built (got {xd.dtype})')
    x2, lead, trail = _arrays.as_channels(xd, axis)
    C, N = x2.shape
    n_bins = N // nb
    dev = x2.device
    ch_stride = x2.stride(0) if C > 1 else N
    out = torch.empty((C, n_bins), dtype=torch.float32, device=dev)
    if n_bins == 0:
        return res.give_back(_arrays.restore_layout(out, lead, trail, 1))

    if isinstance(kind, str) and kind in _DIRECT:
        ws_bytes = _lib.lib.iqw_bin_power_workspace_bytes(C, nb, n_bins)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ptr = ctypes.c_void_p(out.data_ptr())
        which = _DIRECT[kind]
        _lib.check(_lib.lib.iqw_bin_power_c64(
            ctypes.c_void_p(x2.data_ptr()), C, ch_stride, nb, n_bins,
            ptr if which == 'mean' else None, ptr if which == 'max' else None,
            ptr if which == 'min' else None, ctypes.c_void_p(ws.data_ptr()), ws_bytes,
            _stream_ptr(dev)))
    else:
        # order statistics inside each bin: power written as (C, nb, n_bins), statistics over axis 1
        pt = torch.empty((C, nb, n_bins), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib.iqw_envtopow_transposed_c64(
            ctypes.c_void_p(x2.data_ptr()), C, ch_stride, nb, n_bins,
            ctypes.c_void_p(pt.data_ptr()), _stream_ptr(dev)))
        time_statistics(pt, [kind], dB=False, out=out.view(C, 1, n_bins))
    return res.give_back(_arrays.restore_layout(out, lead, trail, 1))
