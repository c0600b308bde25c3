"""the C-ABI library loads on a CPU-only machine and exports every function include/iqw_b200.h
declares; the ctypes table in _lib.py covers exactly that set.  No compute calls."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'iqw_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return set(re.findall(r'\b(iqw_[a-z0-9_]+)\s*\(', text))


def test_every_declared_symbol_is_exported_and_bound():
    from iqwaveform_b200 import _lib

    names = _declared()
    assert len(names) >= 8
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f'{n} declared in the header but not exported'
    assert names == set(_lib.SIGNATURES), 'ctypes table and header disagree'
    assert _lib.lib.iqw_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    from iqwaveform_b200 import _lib

    assert ctypes.sizeof(_lib.iqw_stat) == 32        # int32 + pad, int64, int64, float + pad
    assert _lib.iqw_stat.rank_lo.offset == 8 and _lib.iqw_stat.gamma.offset == 24


def test_workspace_queries_need_no_gpu():
    from iqwaveform_b200 import _lib

    assert _lib.lib.iqw_time_stats_workspace_bytes(1, 1000, 4096, 4) > 4096 * 8 * 1024 * 4
    assert _lib.lib.iqw_bin_power_workspace_bytes(1, 245760, 10) >= 256


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib.util
    import shutil

    src = os.path.join(ROOT, 'iqwaveform_b200', '_lib.py')
    dst = tmp_path / '_lib_copy.py'
    shutil.copy(src, dst)
    spec = importlib.util.spec_from_file_location('_lib_copy', dst)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except ImportError as e:
        assert 'no CPU fallback' in str(e)
    else:
        raise AssertionError('loading without libiqw_b200.so must raise ImportError')
