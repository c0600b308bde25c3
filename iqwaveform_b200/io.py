"""SigMF (npy flavour) capture reader: the feeder of the hot path.

Mirrors /root/reference/src/iqwaveform/io.py:13-96 (`extract_ntia_calibration_metadata`,
`read_sigmf_metadata`, `read_sigmf`) and 109-146 (`waveform_to_frame`): same arguments, same return
tuples.  Host work only -- JSON metadata and an ``.npy`` file.  Two additive arguments of `read_sigmf`:

* ``mmap=True`` maps the data file instead of reading it, so that a capture larger than host memory can be
  handed to ``persistence_spectrum`` / ``spectrogram`` / ``iq_to_bin_power`` piece by piece;
* ``pin=True`` returns page-locked torch tensors, which the chunked host path of ``persistence_spectrum``
  (fourier.py: `_psd_from_host`) copies to the device at PCIe speed while earlier chunks are transformed.

`persistence_spectrum_from_sigmf` is the two calls in one: every capture segment of the file becomes one
channel of the persistence spectrum.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

__all__ = ['extract_ntia_calibration_metadata', 'read_sigmf_metadata', 'read_sigmf', 'waveform_to_frame',
           'persistence_spectrum_from_sigmf']


def extract_ntia_calibration_metadata(metadata: dict) -> dict:
    """io.py:13-34: sensor temperature, noise figure and preselector gain of the first
    CalibrationAnnotation (all None when there is none)"""
    temp_K = noise_fig_dB = gain_dB = None
    for a in metadata['annotations']:
        if a['ntia-core:annotation_type'] == 'CalibrationAnnotation':
            temp_K = a['ntia-sensor:temperature'] + 273.15
            noise_fig_dB = a['ntia-sensor:noise_figure_sensor']
            gain_dB = a['ntia-sensor:gain_preselector']
            break
    return {'ambient temperature (K)': temp_K, 'noise figure (dB)': noise_fig_dB, 'gain (dB)': gain_dB}


def read_sigmf_metadata(metadata_fn, ntia=False):
    """io.py:37-55 -> ({sample_start: frequency}, {sample_start: datetime}, sample_rate, calibration)"""
    with open(metadata_fn, 'r') as fd:
        metadata = json.load(fd)
    captures = [{k.replace('core:', ''): v for k, v in c.items()} for c in metadata['captures']]
    cal = extract_ntia_calibration_metadata(metadata) if ntia else {}
    return ({c['sample_start']: c['frequency'] for c in captures},
            {c['sample_start']: c['datetime'] for c in captures},
            metadata['global']['core:sample_rate'], cal)


def read_sigmf(metadata_path: str, force_sample_rate: float = None, sigmf_data_ext='.npy', stack=False,
               ntia_extensions=False, z0=50, *, mmap: bool = False, pin: bool = False):
    """io.py:58-96 -> (list of capture segments (or an (N, M) array with stack=True), centre
    frequencies, sample period, calibration)"""
    metadata_path = Path(metadata_path)
    center_freqs, timestamps, sample_rate, cal = read_sigmf_metadata(metadata_path, ntia=ntia_extensions)
    if force_sample_rate is not None:
        sample_rate = force_sample_rate
    if sigmf_data_ext != '.npy':
        raise TypeError(f'SIGMF data extension {sigmf_data_ext} not supported')
    data_fn = metadata_path.with_suffix('.sigmf-data.npy')
    x = np.load(data_fn, mmap_mode='r' if mmap else None)
    x_split = np.array_split(x, list(center_freqs.keys())[1:])
    if stack:
        x_split = np.vstack(x_split).T
    if cal.get('gain (dB)', None) is not None:
        print('gain dB: ', cal['gain (dB)'])
        gain = 10 ** (cal['gain (dB)'] / 10.0)
        scale = np.sqrt(gain * 2 / z0)
        x_split = x_split / scale if stack else [s / scale for s in x_split]
    elif ntia_extensions:
        raise LookupError('no calibration data is available in NTIA extensions')
    if pin:
        import torch

        def pinned(a):
            t = torch.empty(a.shape, dtype=torch.complex64, pin_memory=True)
            t.numpy()[...] = a
            return t
        x_split = pinned(x_split) if stack else [pinned(s) for s in x_split]
    return (x_split, np.array(list(center_freqs.values())), 1.0 / sample_rate, cal)


def waveform_to_frame(waveform, Ts: float, columns=None, column_name=None):
    """io.py:109-146: pandas Series (1-D) or DataFrame (2-D, one column per waveform) with a time index"""
    import pandas as pd

    if waveform.ndim == 2:
        if columns is None:
            columns = np.arange(waveform.shape[1])
        obj = pd.DataFrame(waveform, columns=columns)
        if column_name is not None:
            obj.columns.name = column_name
    elif waveform.ndim == 1:
        obj = pd.Series(waveform)
    else:
        raise TypeError('iq must have 1 or 2 dimensions')
    obj.index = pd.Index(np.linspace(0, Ts * waveform.shape[0], waveform.shape[0], endpoint=False),
                         name='Time elapsed (s)')
    return obj


def persistence_spectrum_from_sigmf(metadata_path: str, *, window, resolution: float, statistics,
                                    fractional_overlap=0, dB=True, force_sample_rate: float = None, **kw):
    """read a SigMF capture file and return (centre frequencies, persistence spectrum of every capture
    segment as (segments, nstat, nbins)): `read_sigmf` (memory-mapped) feeding `persistence_spectrum`,
    whose host path copies each segment in chunks and transforms them while they arrive.  Segments must
    have equal length to be stacked; ragged files return a list of (nstat, nbins) arrays."""
    from . import fourier

    segments, freqs, Ts, _ = read_sigmf(metadata_path, force_sample_rate=force_sample_rate, mmap=True)
    fs = 1.0 / Ts
    rows = []
    for seg in segments:
        seg = np.array(seg, dtype=np.complex64)       # out of the (read-only) file mapping
        rows.append(fourier.persistence_spectrum(seg, fs=fs, window=window, resolution=resolution,
                                                 fractional_overlap=fractional_overlap, statistics=statistics,
                                                 dB=dB, axis=0, **kw))
    if len({r.shape for r in rows}) == 1:
        return freqs, np.stack(rows)
    return freqs, rows
