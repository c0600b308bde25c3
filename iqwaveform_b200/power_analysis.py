"""B200-native ``iq_to_bin_power`` with the reference signature
(/root/reference/src/iqwaveform/power_analysis.py:341-385) and the elementwise power transforms
``powtodB`` / ``dBtopow`` / ``envtopow`` / ``envtodB`` / ``dBlinmean`` / ``dBlinsum``
(power_analysis.py:168-338), each one streaming CUDA kernel behind ``iqw_elementwise_*``."""
from __future__ import annotations

import ctypes
import math
from numbers import Number

import torch

from . import _arrays, _lib, _plan
from .fourier import _stream_ptr, time_statistics
from .util import Domain, get_input_domain

__all__ = ['iq_to_bin_power', 'iq_to_cyclic_power', 'powtodB', 'dBtopow', 'envtopow', 'envtodB', 'dBlinmean', 'dBlinsum',
           'sample_ccdf']

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


_CYCLE_STATS = ('min', 'max', 'peak', 'mean', 'rms', 'median')


def iq_to_cyclic_power(x, Ts: float, detector_period: float, cyclic_period: float, truncate=False,
                       detectors=('rms', 'peak'), cycle_stats=('min', 'mean', 'max'), axis=0):
    """time series of periodic frame power statistics (power_analysis.py:388-493): detector power
    on bins of `detector_period` (the bin-power kernel), folded on `cyclic_period`, then a statistic
    over the cycles for every bin of the cycle (the time-statistics kernel on the
    (channels, cycles, bins-per-cycle) view -- no data is moved in between).

    Supported layouts: (channels, time) with axis=1 -- the one the reference supports (it tests
    ``power_shape[1]``, line 454) -- and 1-D captures.  Returns {detector: {statistic: array}}."""
    domain = get_input_domain()
    if domain == Domain.TIME_BINNED_POWER:
        # precalculated binned power: dict keyed by detector (power_analysis.py:438-448)
        if not isinstance(x, dict):
            raise TypeError('in time-binned power domain, expected dict input keyed by detector')
        if detectors is None:
            detectors = tuple(x.keys())
        elif set(x.keys()) != set(detectors):
            raise ValueError('input data keys do not match supplied ')
    elif domain != Domain.TIME:
        raise ValueError(f'unsupported cyclic power domain "{domain}"')
    elif detectors is None:
        raise ValueError('supply detectors argument to evaluate binned power from time domain IQ')
    if _plan.isroundmod(cyclic_period, detector_period, atol=1e-6):
        nbins = round(cyclic_period / detector_period)
    else:
        raise ValueError('cyclic period must be positive integer multiple of the detector period')
    for k in cycle_stats:
        if k not in _CYCLE_STATS:
            raise ValueError(f'kind argument must be one of {list(_CYCLE_STATS)}')
    first = x[detectors[0]] if domain == Domain.TIME_BINNED_POWER else x
    shape = getattr(first, 'shape', None)
    if shape is None:
        raise TypeError('unrecognized object type')
    ax = axis + len(shape) if axis < 0 else axis
    if not ((len(shape) == 1 and ax == 0) or (len(shape) == 2 and ax == 1)):
        raise NotImplementedError('iq_to_cyclic_power is built for (channels, time) axis=1 and 1-D captures')
    if domain == Domain.TIME:
        xd, res = _arrays.to_device(x)
    ret = {}
    for d in detectors:
        if domain == Domain.TIME:
            p = iq_to_bin_power(xd, Ts, detector_period, kind=d, truncate=truncate, axis=ax)  # (C, n_bins) | (n_bins,)
        else:
            p, res = _arrays.to_device(x[d])
            if p.dtype != torch.float32:
                raise NotImplementedError(f'only float32 binned power is built (got {p.dtype})')
            p = p.contiguous()
        n_det = p.shape[-1]
        if nbins < 1 or n_det % nbins != 0:
            raise ValueError('pass truncate=True to allow truncation to align with cyclic windows')
        p3 = p.reshape(-1, n_det // nbins, nbins)
        stats = time_statistics(p3, list(cycle_stats), dB=False)                            # (C, nstat, nbins)
        ret[d] = {k: res.give_back(stats[:, i].reshape(p.shape[:-1] + (nbins,)))
                  for i, k in enumerate(cycle_stats)}
    return ret


# ---------------------------------------------------------------------------------------------
# elementwise transforms (power_analysis.py:168-338)
# ---------------------------------------------------------------------------------------------
# unit bookkeeping of xarray results (power_analysis.py:39-70): regex substitutions on attrs['units']
_DB_UNIT_MAPPING = {'dBm': 'mW', 'dBW': 'W', 'dB': 'unitless'}


def _unit_sub(prefix_from, prefix_to):
    import re

    def transform(s: str) -> str:
        for db_unit, lin_unit in _DB_UNIT_MAPPING.items():
            s, _ = re.subn('^' + prefix_from(db_unit, lin_unit), prefix_to(db_unit, lin_unit), s, count=1)
        return s
    return transform


unit_dB_to_linear = _unit_sub(lambda d, l: d, lambda d, l: l)
unit_linear_to_dB = _unit_sub(lambda d, l: l, lambda d, l: d)
unit_dB_to_wave = _unit_sub(lambda d, l: d, lambda d, l: '\u221a' + l)
unit_wave_to_dB = _unit_sub(lambda d, l: '\u221a' + l, lambda d, l: d)
unit_wave_to_linear = _unit_sub(lambda d, l: '\u221a' + l, lambda d, l: l)
_UNIT_TRANSFORM = {_lib.EW_POWTODB: unit_linear_to_dB, _lib.EW_DBTOPOW: unit_dB_to_linear,
                   _lib.EW_ENVTOPOW: unit_wave_to_linear, _lib.EW_ENVTODB: unit_wave_to_dB}


def _is_labelled(x) -> bool:
    """pandas.Series / pandas.DataFrame / xarray.DataArray: objects that wrap an array in `.values`
    (power_analysis.py:113-118), without importing either package"""
    return hasattr(x, 'values') and not isinstance(x, (torch.Tensor, dict)) and not hasattr(x, '__cuda_array_interface__') \
        and type(x).__module__.split('.')[0] in ('pandas', 'xarray')


def _repackage_arraylike(values, obj, unit_transform=None):
    """package `values` into a data type matching `obj` (power_analysis.py:139-165)"""
    mod = type(obj).__module__.split('.')[0]
    if mod == 'pandas':
        import pandas as pd
        if isinstance(obj, pd.Series):
            return pd.Series(values, index=obj.index)
        if isinstance(obj, pd.DataFrame):
            return pd.DataFrame(values, index=obj.index, columns=obj.columns)
    elif mod == 'xarray':
        ret = obj.copy(deep=False, data=values)
        units = ret.attrs.get('units', None)
        if units is not None and unit_transform is not None:
            ret.attrs['units'] = unit_transform(units)
        return ret
    raise TypeError(f'unrecognized input type {type(obj)}')


def _elementwise(x, op: int, *, use_abs: bool = True, eps: float = 0.0, out=None):
    """float32 / complex64 array-like -> float32 of the same shape, in the caller's kind.  Labelled
    containers (pandas Series / DataFrame, xarray DataArray) are unwrapped, transformed on the device
    and re-wrapped with their index / columns / coordinates, as the reference does
    (power_analysis.py:104-165); host arrays of another real dtype (integers, float64) are computed
    in float32 on the device and returned in the dtype the reference would return (float64)."""
    if _is_labelled(x):
        import numpy as np
        values = np.asarray(x.values)
        if out is not None and hasattr(out, 'values'):
            out = out.values
        res = _elementwise(values, op, use_abs=use_abs, eps=eps, out=None)
        if out is not None:
            out[...] = res
            res = out
        return _repackage_arraylike(res, x, _UNIT_TRANSFORM.get(op))
    if not isinstance(x, (Number, torch.Tensor)) and hasattr(x, 'dtype') and hasattr(x, 'astype'):
        import numpy as np
        if isinstance(x, np.ndarray) and x.dtype not in (np.float32, np.complex64):
            if np.iscomplexobj(x):
                return _elementwise(x.astype(np.complex64), op, use_abs=use_abs, eps=eps, out=out).astype(np.float64)
            if x.dtype.kind in 'iubf':
                wide = np.float64 if x.dtype.itemsize >= 8 or x.dtype.kind in 'iub' else np.float32
                return _elementwise(x.astype(np.float32), op, use_abs=use_abs, eps=eps, out=out).astype(wide)
    if isinstance(x, Number):            # scalars never touch the device (reference: numexpr)
        if op == _lib.EW_DBTOPOW:
            return 10.0 ** (x / 10.0)
        if op == _lib.EW_ENVTOPOW:
            return float(abs(x)) ** 2
        v = (abs(x) if use_abs else x) + eps
        scale = 10.0 if op == _lib.EW_POWTODB else 20.0
        if v == 0:
            return -math.inf
        return scale * math.log10(v) if v > 0 else math.nan
    xd, res = _arrays.to_device(x)
    if xd.dtype == torch.complex64:
        if op not in (_lib.EW_ENVTOPOW, _lib.EW_ENVTODB):
            raise TypeError('powtodB / dBtopow expect real input')
        if not use_abs:
            raise NotImplementedError('abs=False on complex input is not built')
    elif xd.dtype != torch.float32:
        raise NotImplementedError(f'only float32 and complex64 input is built (got {xd.dtype})')
    xc = xd.contiguous()
    if out is not None:
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.float32 and
                out.shape == xc.shape and out.is_contiguous()):
            raise ValueError('out must be a contiguous float32 CUDA tensor of the input shape')
        y = out
    else:
        y = torch.empty(xc.shape, dtype=torch.float32, device=xc.device)
    n = xc.numel()
    if xc.dtype == torch.complex64:
        _lib.check(_lib.lib.iqw_elementwise_c64(op, ctypes.c_void_p(xc.data_ptr()), ctypes.c_void_p(y.data_ptr()),
                                                n, float(eps), _stream_ptr(xc.device)))
    else:
        _lib.check(_lib.lib.iqw_elementwise_f32(op, ctypes.c_void_p(xc.data_ptr()), ctypes.c_void_p(y.data_ptr()),
                                                n, int(bool(use_abs)), float(eps), _stream_ptr(xc.device)))
    return y if out is not None else res.give_back(y)


def powtodB(x, abs: bool = True, eps: float = 0, out=None):
    """``10*log10(abs(x) + eps)`` (or without abs), power_analysis.py:168-206"""
    return _elementwise(x, _lib.EW_POWTODB, use_abs=abs, eps=eps, out=out)


def dBtopow(x, out=None):
    """``10**(x/10)``, power_analysis.py:209-231"""
    return _elementwise(x, _lib.EW_DBTOPOW, out=out)


def envtopow(x, out=None):
    """``abs(x)**2`` of a real or complex waveform, power_analysis.py:234-257"""
    return _elementwise(x, _lib.EW_ENVTOPOW, out=out)


def envtodB(x, abs: bool = True, eps: float = 0, out=None):
    """``20*log10(abs(x) + eps)`` (or without abs), power_analysis.py:260-298"""
    return _elementwise(x, _lib.EW_ENVTODB, use_abs=abs, eps=eps, out=out)


def _dBlin(x_dB, axis, reduce: str):
    """powtodB(dBtopow(x).<mean|sum>(axis)) (power_analysis.py:301-338).  dBtopow and powtodB run in
    the elementwise kernel, the reduction of the linear power in the time-statistics kernel, whose
    reduction axis is the row axis of a (channels, rows, columns) layout: built for axis 0 of a
    2-D array, axis 1 (-2) of a 3-D array, and for 1-D arrays / axis=None up to 2**20 elements."""
    xd, res = _arrays.to_device(x_dB)
    lin = _elementwise(xd, _lib.EW_DBTOPOW)
    nd = lin.ndim
    ax = None if axis is None else (axis + nd if axis < 0 else axis)
    if ax is None or nd == 1:
        if lin.numel() > (1 << 20):
            raise NotImplementedError('reduce large arrays along axis 0 of a 2-D (rows, columns) layout')
        lin3, out_shape = lin.reshape(1, -1, 1), ()
    elif nd == 2 and ax == 0:
        lin3, out_shape = lin.reshape(1, lin.shape[0], lin.shape[1]), (lin.shape[1],)
    elif nd == 3 and ax == 1:
        lin3, out_shape = lin, (lin.shape[0], lin.shape[2])
    else:
        raise NotImplementedError('dBlinmean / dBlinsum are built for axis 0 of 2-D and axis 1 of 3-D arrays')
    n = lin3.shape[1]
    mean = time_statistics(lin3, ['mean'], dB=False)              # (C, 1, cols) float32
    y = _elementwise(mean, _lib.EW_POWTODB)                          # 10*log10(mean)
    if reduce == 'sum':                                               # log10(n * mean) = log10(mean) + log10(n)
        y.add_(10.0 * math.log10(n))     # (C, cols) scalars-per-column epilogue on the reduced result
    return res.give_back(y.reshape(out_shape))


def dBlinmean(x_dB, axis=None, overwrite_x=False):
    """mean in linear power of values in dB, power_analysis.py:301-318"""
    return _dBlin(x_dB, axis, 'mean')


def dBlinsum(x_dB, axis=None, overwrite_x=False):
    """sum in linear power of values in dB, power_analysis.py:321-338"""
    return _dBlin(x_dB, axis, 'sum')


def _edge_counts(a2: torch.Tensor, edges, side_right: bool) -> torch.Tensor:
    """(rows, n) float32 on the device, ascending edges -> (rows, n_edges + 1) int64 counts of
    searchsorted(edges, a, side) (csrc/iqw_histogram.cu)"""
    if a2.dtype != torch.float32:
        raise NotImplementedError(f'only float32 samples are built (got {a2.dtype})')
    if not a2.is_contiguous():
        a2 = a2.contiguous()
    e = torch.as_tensor(edges)
    if e.ndim != 1 or e.numel() < 1:
        raise ValueError('edges must be a non-empty vector')
    e = e.to(device=a2.device, dtype=torch.float64).contiguous()
    counts = torch.empty((a2.shape[0], e.numel() + 1), dtype=torch.int64, device=a2.device)
    _lib.check(_lib.lib.iqw_edge_counts_f32(
        ctypes.c_void_p(a2.data_ptr()), a2.shape[0], a2.shape[1], ctypes.c_void_p(e.data_ptr()), e.numel(),
        int(side_right), ctypes.c_void_p(counts.data_ptr()), _stream_ptr(a2.device)))
    return counts


def sample_ccdf(a, edges, density: bool = True):
    """fraction (or number) of the samples of the vector `a` that exceed each edge; same arguments as
    the reference (power_analysis.py:552-583): searchsorted(edges, a, 'left') -> bincount ->
    ``a.size - cumsum``, int64 counts or float64 fractions"""
    ad, res = _arrays.to_device(a)
    if ad.ndim != 1:
        raise ValueError('object too deep for desired array')      # numpy.bincount's message for non-1-D input
    if ad.numel() == 0:
        raise ValueError('sample_ccdf of an empty vector')
    counts = _edge_counts(ad.reshape(1, -1), edges, side_right=False)[0]
    ccdf = (ad.shape[0] - counts.cumsum(0))[:-1]
    if density:
        # a tensor divisor: torch turns division by a python scalar into a multiplication by 1/n
        ccdf = ccdf.to(torch.float64) / torch.tensor(ad.shape[0], dtype=torch.float64, device=ccdf.device)
    return res.give_back(ccdf)
