"""host-side helpers of the reference's util module (/root/reference/src/iqwaveform/util.py) that sit
on the boundary of the accelerated path: the input-domain context (util.py:118,
144-166: callers declare that the arrays they pass are already an STFT or binned power), the
parameter checks and dtype rules the public functions apply, the axis helpers, and
`histogram_last_axis`.  Everything here except the histogram is plain host logic on array metadata;
the helpers accept numpy arrays and torch tensors alike."""
from __future__ import annotations

from contextlib import contextmanager
from enum import Enum

__all__ = ['Domain', 'set_input_domain', 'get_input_domain', 'histogram_last_axis', 'isroundmod', 'find_float_inds',
           'float_dtype_like', 'dtype_change_float', 'axis_index', 'axis_slice', 'to_blocks']

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


# ---------------------------------------------------------------------------------------------
# parameter checks, dtype rules and axis helpers (util.py:121-141, 365-397, 400-442, 466-494, 545-568)
# ---------------------------------------------------------------------------------------------
def isroundmod(value, div, atol: float = 1e-6):
    """is value/div within atol of an integer?  (arrays: elementwise)"""
    import math
    import numpy as np
    ratio = value / div
    if isinstance(ratio, (int, float)):
        return abs(math.remainder(ratio, 1)) <= atol
    return np.abs(np.rint(ratio) - ratio) <= atol


def find_float_inds(seq) -> list:
    """flags the entries of a statistics list that read as floats (quantiles)"""
    from ._plan import find_float_inds as impl
    return impl(seq)


_FLOAT_OF = {'float16': 'float16', 'float32': 'float32', 'float64': 'float64', 'complex64': 'float32',
             'complex128': 'float64', 'bfloat16': 'float32'}


def _dtype_name(x) -> str:
    d = getattr(x, 'dtype', x)
    return str(d).replace('torch.', '')


def float_dtype_like(x, min_dtype=None):
    """numpy float dtype of the real and imaginary parts of x (an array, tensor or number); other
    kinds of data count as float32; never smaller than `min_dtype` when that is given"""
    import numpy as np
    from numbers import Number
    name = _dtype_name(np.asarray(x)) if isinstance(x, Number) else _dtype_name(x)
    dtype = np.dtype(_FLOAT_OF.get(name, 'float32'))
    if min_dtype is not None and np.dtype(min_dtype).itemsize > dtype.itemsize:
        dtype = np.dtype(min_dtype)
    return dtype


def dtype_change_float(dtype, float_basis_dtype):
    """the dtype of the same kind as `dtype` (real or complex) built on the float type of
    `float_basis_dtype`: (complex128, float32) -> complex64, (float64, float32) -> float32"""
    import numpy as np
    kind = np.dtype(dtype).type
    basis = np.finfo(np.dtype(float_basis_dtype)).dtype.type
    if kind in (np.complex64, np.complex128):
        if basis is np.float32:
            return np.complex64
        if basis is np.float64:
            return np.complex128
    elif kind in (np.float16, np.float32, np.float64):
        return basis
    raise ValueError(f'unable to identify output dtype similar to {dtype} matching floating point {float_basis_dtype}')


def _index_on_axis(ndim: int, axis: int, item) -> tuple:
    if not -ndim <= axis < ndim:
        raise ValueError(f'axis {axis} exceeds the number of dimensions')
    axis %= ndim
    return (slice(None),) * axis + (item,) + (slice(None),) * (ndim - axis - 1)


def axis_index(a, index, axis: int = -1):
    """a[..., index, ...] with `index` (an integer array or boolean mask) applied on `axis`"""
    return a[_index_on_axis(a.ndim, axis, index)]


def axis_slice(a, start, stop=None, step=None, axis: int = -1):
    """a[..., start:stop:step, ...] on `axis`"""
    return a[_index_on_axis(a.ndim, axis, slice(start, stop, step))]


def to_blocks(y, size: int, truncate: bool = False, axis: int = 0):
    """view of y with `axis` split into (blocks, size); a ragged tail is an error unless `truncate`"""
    if not isinstance(size, int):
        raise TypeError('block size must be integer')
    n_total = y.numel() if hasattr(y, 'numel') else y.size
    if n_total == 0:
        raise IndexError('cannot form blocks on arrays of size 0')
    n = y.shape[axis]
    if n % size:
        if not truncate:
            raise ValueError(f'axis 0 size {n} is not a factor of block size {size}')
        y = axis_slice(y, None, size * (n // size), axis=axis)
    shape = tuple(y.shape)
    after = () if axis == -1 else shape[axis + 1:]
    return y.reshape(shape[:axis] + (n // size, size) + after)
