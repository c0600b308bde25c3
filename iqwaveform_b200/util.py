"""input-domain context of the reference (/root/reference/src/iqwaveform/util.py:118, 144-166):
callers declare that the arrays they pass are already in another domain (an STFT, binned power)."""
from __future__ import annotations

from contextlib import contextmanager
from enum import Enum

__all__ = ['Domain', 'set_input_domain', 'get_input_domain']

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
