"""input-domain context of the reference (/root/reference/src/iqwaveform/util.py:118, 144-166):
callers declare that the arrays they pass are already in another domain (an STFT, binned power)."""
from __future__ import annotations

from contextlib import contextmanager
from enum import Enum

__all__ = ['Domain', 'set_input_domain', 'get_input_domain', 'histogram_last_axis']

_input_domain: list = []


class Domain(Enum):
    TIME = 'time'
    FREQUENCY = 'frequency'
    TIME_BINNED_POWER = 'time_binned_power'


@contextmanager
def set_input_domain(domain):
    """set the current domain from input arrays of DSP calls (util.py:150-156)"""
    i = len(_input_domain)
    _input_domain.append(Domain(domain))
    try:
        yield
    finally:
        del _input_domain[i]


def get_input_domain(default=Domain.TIME):
    Domain(default)
    return _input_domain[-1] if _input_domain else default


def histogram_last_axis(x, bins, range=None):
    """histogram along the last axis; same arguments and return value as the reference
    (util.py:497-543): ``(counts of shape x.shape[:-1] + (n_bins,), bin edges)``.  Like the reference,
    bin i holds edges[i] <= x < edges[i+1] and a value equal to the last edge is NOT counted."""
    import numpy as np
    import torch

    from . import _arrays
    from .power_analysis import _edge_counts

    xd, res = _arrays.to_device(x)
    if isinstance(bins, int):
        if range is None:       # not a hot path: the reference's only caller passes bounds
            lo, hi = torch.aminmax(xd)
            range = np.float32(lo.item()), np.float32(hi.item())    # float32 scalars: float32 edges, like numpy
        edges = np.linspace(range[0], range[1], bins + 1)
    else:
        edges = np.asarray(bins.cpu() if isinstance(bins, torch.Tensor) else bins)
    counts = _edge_counts(xd.reshape(-1, xd.shape[-1]), edges, side_right=True)
    hist = counts[:, 1:edges.size].reshape(tuple(xd.shape[:-1]) + (edges.size - 1,))
    if res.kind == 'numpy':
        edges_out = np.asarray(edges)
    else:
        edges_out = torch.as_tensor(edges, device=xd.device if res.kind == 'torch_cuda' else 'cpu')
    return res.give_back(hist), edges_out
