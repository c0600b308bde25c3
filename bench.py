#!/usr/bin/env python
"""bench.py -- headline benchmark: complex IQ GS/s of persistence_spectrum on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): persistence
spectrum of 100 MS/s complex64 captures, 10 s per channel (1e9 samples), nfft 4096 Hann, 50 %
overlap, quantiles [0.1, 0.5, 0.9, 0.999], dB.  Channels are independent, so the job shards by
channel with no data-path collective: ONE channel per GPU ("weak" scaling; 8 GPUs = the full
8-channel config), plus the NCCL all_gather of the (4, 4096) result rows the north-star names.

A step = one full pass of the hot path over one channel per GPU:
    kernel 1 (STFT -> |X|^2, 8 GB in / 8 GB out)  ->  kernel 2 (exact per-bin order statistics).
`value`  : device-resident input, CUDA events on the launching stream, max over ranks.
`e2e`    : the same call with the capture in pinned HOST memory (host->device copy of 8 GB and
           device->host read of the result inside the timed region).
`roofline`: dominant kernel's algorithmic bytes / its measured duration (library profile events
           recorded inside the timed region) against MEASURED_PEAKS.json.
`cpu_baseline`: the numpy/scipy oracle port of the reference (same library calls as the
           reference: scipy.fft with cpu_count//2 workers, np.quantile) on a bounded sample.
--impl reference times that CPU port alone (rank 0), as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 100e6
NFFT = 4096
OVERLAP = 0.5
STATS = [0.1, 0.5, 0.9, 0.999]
WINDOW = 'hann'
SAMPLES_PER_CHANNEL = 1_000_000_000
CPU_SAMPLE = 1 << 25          # bounded sample for the CPU arms (0.34 s of one channel)
METRIC = 'complex IQ GS/s for persistence_spectrum'
UNIT = 'GS/s'


def workload_config(n_gpus, samples):
    return {
        'workload': 'BASELINE configs[2] persistence_spectrum, sharded by channel: 1 channel per GPU',
        'channels': n_gpus, 'channels_per_gpu': 1, 'samples_per_channel': samples,
        'sample_rate_hz': FS, 'nfft': NFFT, 'window': WINDOW, 'overlap': OVERLAP,
        'statistics': STATS, 'dB': True,
        'frames_per_channel': (samples - NFFT) // (NFFT // 2) + 1,
        'l2': 'no flush: per-step input (8 B/sample) and spectrogram (8 B/sample) exceed the 126 MB L2',
        'collective': 'nccl all_gather of the (4, 4096) fp32 result per channel' if n_gpus > 1 else 'none',
    }


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference
# ------------------------------------------------------------------------------------------------
def cpu_capture(n, seed=1234):
    import numpy as np

    rng = np.random.default_rng(seed)
    x = np.empty(n, dtype=np.complex64)
    chunk = 1 << 22
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        k = np.arange(s, s + m, dtype=np.float64)
        z = (rng.standard_normal(m) + 1j * rng.standard_normal(m)) * math.sqrt(0.5)
        for f, a in ((0.0651, 0.5), (-0.2148, 0.05), (0.3256, 3.0)):
            z += a * np.exp(2j * np.pi * ((f * k) % 1.0))
        x[s:s + m] = z
    return x


def cpu_step(x):
    from oracle import iqw_oracle as orc

    return orc.persistence_spectrum(x[None, :], fs=FS, window=WINDOW, resolution=FS / NFFT,
                                    fractional_overlap=OVERLAP, statistics=STATS, dB=True, axis=1)


def cpu_cores():
    return max((os.cpu_count() or 1) // 2, 1)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    x = cpu_capture(CPU_SAMPLE)
    for _ in range(max(args.warmup, 1)):
        cpu_step(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(x)
    dt = time.perf_counter() - t0
    value = CPU_SAMPLE * args.steps / dt / 1e9
    sample = (f'1 channel x {CPU_SAMPLE} samples per step ({CPU_SAMPLE / FS:.3f} s of the 10 s capture), same '
              f'nfft/overlap/statistics; numpy/scipy oracle port of the reference, scipy.fft workers = '
              f'cpu_count//2 = {cpu_cores()} of {os.cpu_count()} cores (the reference policy, fourier.py:214), '
              f'numpy stages single-threaded as in the reference')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': max(args.warmup, 1), 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': workload_config(args.gpus, SAMPLES_PER_CHANNEL),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cpu_cores(), 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    return line


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '50',
                 '-i', str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def device_capture(torch, n, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty(n, dtype=torch.complex64, device=device)
    xr = torch.view_as_real(x)
    chunk = 1 << 26
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        xr[s:s + m].normal_(0.0, math.sqrt(0.5), generator=g)
        k = torch.arange(s, s + m, device=device, dtype=torch.float64)
        for f, a in ((0.0651, 0.5), (-0.2148, 0.05), (0.3256, 3.0)):
            ph = (2 * math.pi) * torch.remainder(f * k, 1.0)
            xr[s:s + m, 0] += (a * torch.cos(ph)).float()
            xr[s:s + m, 1] += (a * torch.sin(ph)).float()
    return x


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (of measured)'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'B200_PROFILING.md fallback 6.65 TB/s (of fallback)'


def ncu_traffic(kernel):
    """dram bytes per launch from the committed ncu --set full capture of this workload, if any"""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            return json.load(f).get(kernel)
    except (OSError, ValueError):
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    import iqwaveform_b200 as iqw
    from iqwaveform_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa_cpus = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # one process per GPU: keep each rank's pinned host buffer on the socket its GPU hangs off
        numa_cpus = iqw.distributed.bind_to_gpu_numa_node(local)
        dist.init_process_group('nccl', device_id=dev)
    n = args.samples
    warmup = max(args.warmup, 3)

    x = device_capture(torch, n, 1234 + 1000 * rank, dev).view(1, n)
    kw = dict(fs=FS, window=WINDOW, resolution=FS / NFFT, fractional_overlap=OVERLAP,
              statistics=STATS, dB=True, axis=1)

    def step(inp):
        out = iqw.persistence_spectrum(inp, **kw)
        if world > 1:       # the (4, 4096) rows of every channel on every rank: one NCCL all_gather
            iqw.distributed.gather_rows(out if out.is_cuda else out.to(dev))
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None     # clocks under load: warm-up + timed region
    for _ in range(warmup):
        step(x)
    barrier()

    # ---- device-resident timed region --------------------------------------------------------
    _lib.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(x)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    prof = _lib.profile_report()
    _lib.profile(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * args.steps / (ms * 1e-3) / 1e9

    # ---- end to end: pinned host input, result read back ---------------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty((1, n), dtype=torch.complex64, pin_memory=True)
        host.copy_(x)
        torch.cuda.synchronize()
        res = step(host)          # warm-up (allocator, pinned result buffer)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step(host)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {'value': world * n * args.steps / dt / 1e9, 'unit': UNIT,
               'h2d_bytes_per_step': world * n * 8,
               'd2h_bytes_per_step': world * res.numel() * 4,
               'ms_per_step': dt / args.steps * 1e3,
               'api': 'iqwaveform_b200.persistence_spectrum(pinned CPU torch tensor) -> CPU tensor',
               'host_cpus_rank0': numa_cpus}
        del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    # ---- roofline of the dominant kernel -------------------------------------------------------
    T = (n - NFFT) // (NFFT // 2) + 1
    alg_bytes = {                       # algorithmic bytes per launch
        'stft_kernel': 8 * n + 4 * T * NFFT,            # read each sample once, write |X|^2 once
        'stats_bracket_pass': 4 * T * NFFT,             # read |X|^2 once (candidate lists are overhead)
    }
    peak, peak_src = measured_peak()
    stages, launches = [], 0
    for name, (cnt, tot) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        launches += cnt
        avg = tot / cnt
        b = alg_bytes.get(name)
        gbs = b / (avg * 1e-3) / 1e9 if (b and avg > 0.01) else None
        stages.append({'kernel': name, 'launches': cnt, 'avg_ms': round(avg, 4),
                       'share': round(tot / ms, 4),
                       'algorithmic_GBps': round(gbs, 1) if gbs else None,
                       'frac': round(gbs / peak, 4) if gbs else None})
    top = next(s for s in stages if s['algorithmic_GBps'])
    roofline = {'bound': 'hbm', 'kernel': top['kernel'], 'achieved': top['algorithmic_GBps'],
                'peak': peak, 'unit': 'GB/s', 'frac': top['frac'], 'peak_source': peak_src,
                'traffic': ncu_traffic(top['kernel']),
                'algorithmic_bytes_per_launch': alg_bytes[top['kernel']],
                'avg_launch_ms': top['avg_ms'], 'share_of_step': top['share'],
                # whole pipeline against the two lower bounds of SURVEY.md 8d
                'pipeline_materialise_once_frac': round(24 * value / peak, 4),
                'pipeline_compulsory_frac': round(8 * value / peak, 4),
                'stages': stages}

    # ---- CPU baseline on a bounded sample ------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        xc = cpu_capture(CPU_SAMPLE)
        cpu_step(xc)
        best = float('inf')
        for _ in range(2):
            t0 = time.perf_counter(); cpu_step(xc); best = min(best, time.perf_counter() - t0)
        cpu = {'value': CPU_SAMPLE / best / 1e9, 'unit': UNIT, 'cores': cpu_cores(), 'kind': 'port',
               'sample': f'1 channel x {CPU_SAMPLE} samples ({CPU_SAMPLE / FS:.3f} s of capture), same '
                         f'nfft/overlap/statistics, best of 2 after 1 warm-up; oracle port = the reference\'s '
                         f'own scipy.fft (workers=cpu_count//2={cpu_cores()} of {os.cpu_count()}) + np.quantile calls'}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(world, n), 'clocks': clocks, 'e2e': e2e,
        'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'impl': 'b200',
    }
    if world > 1:
        dist.destroy_process_group()
    return line


class QuietStdout:
    """everything libraries print to fd 1 while the benchmark runs (NCCL's version banner, ...)
    goes to stderr, so that stdout carries exactly one JSON line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--samples', type=int, default=SAMPLES_PER_CHANNEL,
                    help='samples per channel (default: the full 10 s capture)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        with QuietStdout():
            line = run_reference(args)
        if line is not None:
            print(json.dumps(line), flush=True)
        return 0
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               f'--nproc-per-node={args.gpus}', '--master-addr', '127.0.0.1', '--master-port', '29517',
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    with QuietStdout():
        line = run_b200(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == '__main__':
    sys.exit(main())
