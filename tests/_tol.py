"""parity tolerances (BASELINE.json names them, BASELINE.md section 4.7 / SURVEY.md 8d give the numbers).

* frame/bin indexing, shapes, axis arrays: exact
* stft / power:  |d| <= 1e-5*|ref| + 2e-7*max_k(power of the frame)          (POWER_RTOL, POWER_FLOOR)
  (two correct fp32 FFTs differ by rounding noise proportional to the frame's total energy, so a
  purely relative bound is not met even by the reference against float64 truth -- SURVEY.md H3)
* dB values: the dB image of the bound above plus DB_ATOL = 5e-5 dB for the logarithm itself
* order statistics are 1-Lipschitz in the sup norm, so quantile / max / min rows get the element
  bound evaluated at the row's own level; selection itself is exact (checked bitwise elsewhere)
* mean rows: 1e-3 dB against the reference (its fp32 running sum drifts) and the element bound
  against float64 truth
"""
import numpy as np

POWER_RTOL = 1e-5
POWER_FLOOR = 2e-7
DB_ATOL = 5e-5
DB_PER_REL = 10.0 / np.log(10.0)


def power_tol(ref_power):
    """elementwise tolerance for a (..., T, nbins) power array"""
    frame_max = np.max(ref_power, axis=-1, keepdims=True)
    return POWER_RTOL * np.abs(ref_power) + POWER_FLOOR * frame_max


def complex_tol(ref_y):
    """amplitude version: |dy| <= 0.5e-5*|y| + 1e-7*max_k|y|  (half the power bound, first order)"""
    mag = np.abs(ref_y)
    return 0.5 * POWER_RTOL * mag + 2e-7 * np.max(mag, axis=-1, keepdims=True)


def power_err_units(got, ref_power):
    """max error in units of the tolerance (<= 1 passes)"""
    return float(np.max(np.abs(got.astype(np.float64) - ref_power) / power_tol(ref_power.astype(np.float64))))


def db_tol(ref_db, peak_power, eps=0.0):
    """tolerance for dB values whose frames peak at `peak_power` (scalar or broadcastable)"""
    p = 10.0 ** (np.asarray(ref_db, dtype=np.float64) / 10.0)
    p = np.maximum(p, 1e-300)
    return DB_ATOL + DB_PER_REL * (POWER_RTOL + POWER_FLOOR * peak_power / p)


WAVE_FLOOR = 4e-7


def waveform_tol(ref_x):
    """inverse STFT / overlap-add output: |dx| <= 0.5e-5*|x| + 4e-7*max_n|x| (two float32 inverse
    FFTs differ by rounding noise proportional to the frame energy; 2..4 frames are summed)"""
    mag = np.abs(ref_x)
    return 0.5 * POWER_RTOL * mag + WAVE_FLOOR * np.max(mag, axis=-1, keepdims=True)
