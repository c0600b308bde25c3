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


def test_generated_bracket_visit_is_in_sync(tmp_path):
    """csrc/bracket_visit.inc is the output of tools/gen_bracket_visit.py: regenerate it and compare (a hand edit of
    the inline PTX, or a generator change without a regeneration, fails here)"""
    import subprocess
    import sys

    inc = os.path.join(ROOT, 'iqwaveform_b200', 'csrc', 'bracket_visit.inc')
    before = open(inc).read()
    try:
        subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'gen_bracket_visit.py')], check=True,
                       capture_output=True)
        after = open(inc).read()
    finally:
        open(inc, 'w').write(before)
    assert after == before


def test_bracket_visit_difference_form_on_the_host():
    """the arithmetic of the raw-bit bracket test (csrc/bracket_visit.inc, bracket_visit_diff): for a key v and a
    bracket [lo, hi] of raw float bits, d = max(v, 0) - lo as uint32 has its sign bit set exactly when the key lies
    below lo, and d < hi - lo + 1 exactly when it lies inside -- checked against float comparisons for every pair of a
    value set that holds negative values, signed zeros, denormals, infinities and NaN patterns"""
    import numpy as np

    vals = np.array([-np.inf, -3.0, -1e-40, -0.0, 0.0, 1e-45, 1e-40, 1.17549435e-38, 0.5, 1.0, 1.0000001, 3.0, 1e30,
                     np.inf], dtype=np.float32)
    bits = vals.view(np.uint32).astype(np.int64)
    bits = np.concatenate([bits, [0x7FC00000, 0x7FFFFFFF, 0xFFC00000]])      # +NaN patterns, -NaN
    signed = np.where(bits >= 2 ** 31, bits - 2 ** 32, bits)
    key = np.where(bits >= 2 ** 31, (~bits) & 0xFFFFFFFF, bits | 0x80000000)  # order-preserving keys (float_to_key)
    nonneg = [b for b in bits if b < 2 ** 31]
    for lo_bits in [0] + [b for b in nonneg if b > 0]:                        # 0 = open low end; closed bounds are > +0
        for hi_bits in [b for b in nonneg if b >= lo_bits]:
            klo = 0 if lo_bits == 0 else (lo_bits | 0x80000000)
            khi = hi_bits | 0x80000000
            wd = (hi_bits - lo_bits + 1) & 0xFFFFFFFF
            vc = np.maximum(signed, 0)
            d = (vc - lo_bits) & 0xFFFFFFFF
            below = (d >> 31).astype(bool)
            inside = d < wd
            assert np.array_equal(below, key < klo), (hex(lo_bits), hex(hi_bits))
            assert np.array_equal(inside, (key >= klo) & (key <= khi)), (hex(lo_bits), hex(hi_bits))
