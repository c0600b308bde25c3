"""TEST INFRASTRUCTURE ONLY -- import the unmodified reference (dgkuester/iqwaveform 0.52.0).

The reference cannot be imported as-is in this image: five third-party
modules are absent and one private SciPy symbol has disappeared.  This module installs the
smallest set of stand-ins that lets ``import iqwaveform`` run WITHOUT editing reference code:

* ``array_api_compat``      -> the real vendored copy in ``sklearn.externals`` (authentic dispatch)
* ``numexpr``               -> stub whose ``evaluate`` raises (ndarrays never reach it, fact 0.5)
* ``xarray``/``methodtools``-> empty stand-ins (never touched on the hot path)
* ``scipy.signal.windows._windows._win_equiv`` -> ``{}`` when missing (``windows.py:119``)

``/root/reference`` exists only in the build container.  ``__graft_entry__.build()`` therefore copies the
reference's package directory, byte for byte, to the git-ignored ``baseline/_ref/iqwaveform`` (the
contract's install target; ``pip install --target`` itself fails here because the reference's build
backend, hatchling, is not in the image -- the package is pure Python, so the copy IS the install).
That copy travels to the GPU box with the repo snapshot and is what ``bench.py --impl reference``
times there.  ``load()`` prefers ``/root/reference/src`` and falls back to ``baseline/_ref``; it
returns ``None`` when neither exists.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_SRC = '/root/reference/src'
INSTALLED_SRC = os.path.join(_ROOT, 'baseline', '_ref')


def source_dir() -> str | None:
    """directory holding the unmodified reference package: the read-only checkout in the build
    container, else the copy `build()` installed under baseline/_ref"""
    for d in (REFERENCE_SRC, INSTALLED_SRC):
        if os.path.isfile(os.path.join(d, 'iqwaveform', 'fourier.py')):
            return d
    return None


def available() -> bool:
    return source_dir() is not None


def install() -> str | None:
    """copy the reference package to baseline/_ref (build container only); returns the target or None"""
    import shutil
    src = os.path.join(REFERENCE_SRC, 'iqwaveform')
    if not os.path.isdir(src):
        return None
    dst = os.path.join(INSTALLED_SRC, 'iqwaveform')
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(INSTALLED_SRC, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__'))
    return dst


def _stub(name: str, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    # util.lazy_import() calls importlib.util.find_spec, which needs a __spec__
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m


def load():
    """return the reference package (module ``iqwaveform``) or None when it is not present"""
    if not available():
        return None
    if 'iqwaveform' in sys.modules:
        return sys.modules['iqwaveform']

    try:
        import array_api_compat  # noqa: F401
    except ImportError:
        import sklearn.externals.array_api_compat as aac
        import sklearn.externals.array_api_compat.numpy as aacnp

        sys.modules['array_api_compat'] = aac
        sys.modules['array_api_compat.numpy'] = aacnp

    def _never(*a, **k):
        raise RuntimeError('numexpr stub reached (only python scalars get here)')

    for name, attrs in (
        ('numexpr', dict(evaluate=_never)),
        ('xarray', dict(DataArray=type('DataArray', (), {}))),
        ('methodtools', dict(lru_cache=lambda *a, **k: (lambda f: f))),
    ):
        try:
            importlib.import_module(name)
        except ImportError:
            _stub(name, **attrs)

    import scipy.signal.windows._windows as _w

    if not hasattr(_w, '_win_equiv'):
        _w._win_equiv = {}

    src = source_dir()
    if src not in sys.path:
        sys.path.insert(0, src)
    return importlib.import_module('iqwaveform')
