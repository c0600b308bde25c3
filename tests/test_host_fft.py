"""the device FFT building blocks (fft_core.cuh is __host__ __device__) run thread-by-thread on
the CPU against a float64 DFT, for every supported nfft.  No GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft_core_host_emulation(tmp_path):
    exe = tmp_path / 'fft_emulate'
    subprocess.run(['g++', '-O2', '-std=c++17', '-o', str(exe),
                    os.path.join(ROOT, 'tests', 'host', 'fft_emulate.cpp')], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith('N=')]
    assert len(lines) == 10
    for l in lines:
        assert float(l.split('rel_err=')[1]) < 5e-7, l
