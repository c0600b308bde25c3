#!/usr/bin/env python
"""Hardware numbers for the multi-GPU rows of SURVEY.md section 8e that bench.py does not time
(bench.py is row e2: persistence spectrum sharded by channel).  One process per GPU over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29513 tools/multigpu_rows.py [--out profiles/multigpu_rows_rNN_<N>gpu.json] [--quick]

  e1  BASELINE configs[1]: spectrogram dB, 1e9 samples at 100 MS/s, nfft 2048 Blackman-Harris, 50 %
      overlap, ONE capture split in time with a `noverlap` halo per shard (no collective; strong scaling)
  e3  BASELINE configs[2], one channel split in time: exact quantiles of the whole capture through the
      bracketed radix select (the one real exchange step of the path; strong scaling)
  e4  BASELINE configs[3]: iq_to_bin_power mean / peak, 1 ms bins, 60 s at 245.76 MS/s (14.7e9 samples,
      118 GB) split into bin-aligned shards, NCCL all_gather of the 60 000 results (strong scaling)

Times are CUDA events on every rank's launching stream, bracketed by a barrier, MAX over ranks, best of
3 after a warm-up.  Each row states its limiter as measured.  Parity at full size: e3 against the
single-GPU result on rank 0 (bitwise); e4 and e1: rank 0 regenerates the LAST rank's shard from its
seed and compares bit for bit."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import iqwaveform_b200 as iqw  # noqa: E402
from iqwaveform_b200 import distributed as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--quick', action='store_true', help='a tenth of every capture (plumbing check)')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    peak, _ = bench.measured_peak()
    scale = 10 if a.quick else 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=3):
        out = fn(); del out
        best = float('inf')
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item())); del out
        return best

    rows = []
    sampler = bench.ClockSampler(local) if rank == 0 else None

    def row(name, n, ms, bps, **more):
        gbs = n * bps / ms / 1e6
        r = {'row': name, 'n_gpus': world, 'samples_total': n, 'ms': round(ms, 4), 'GS_per_s': round(n / ms / 1e6, 2),
             'algorithmic_bytes_per_sample': bps, 'algorithmic_GBps_per_gpu': round(gbs / world, 1),
             'frac_of_measured_hbm_peak_per_gpu': round(gbs / world / peak, 4), 'scaling': 'strong', **more}
        if rank == 0:
            print(json.dumps(r), flush=True)
        rows.append(r)

    # ---- e1: configs[1] spectrogram, frame shards with halo ------------------------------------------------
    n, nfft, nov = 1_000_000_000 // scale, 2048, 1024
    sh = D.frame_shard(n, nfft, nov, world, rank)
    # every rank generates ITS samples (+ halo) from per-chunk seeds, so any rank can rebuild any shard
    def shard_samples(s0, s1, seed=7):
        x = torch.empty(s1 - s0, dtype=torch.complex64, device=dev)
        xr = torch.view_as_real(x)
        chunk = 1 << 24
        for c in range(s0 // chunk, (s1 + chunk - 1) // chunk):
            lo, hi = max(s0, c * chunk), min(s1, (c + 1) * chunk)
            g = torch.Generator(device=dev).manual_seed(seed * 1_000_003 + c)
            blk = torch.empty((chunk, 2), dtype=torch.float32, device=dev).normal_(0.0, 0.7, generator=g)
            xr[lo - s0:hi - s0] = blk[lo - c * chunk:hi - c * chunk]
        return x
    x = shard_samples(sh.sample0, sh.sample1)
    kw = dict(n_samples=n, nperseg=nfft, noverlap=nov, fs=100e6, window='blackmanharris', dB=True)
    ms = timed(lambda: D.spectrogram_time_sharded(x, **kw))
    parity = None
    if world > 1:       # rank 0 recomputes the first frames of the last rank's shard
        last = D.frame_shard(n, nfft, nov, world, world - 1)
        k = min(4096, last.frame1 - last.frame0)
        mine = D.spectrogram_time_sharded(x, **kw)
        probe = mine[:k].contiguous() if rank == world - 1 else torch.empty((k, nfft), dtype=torch.float32, device=dev)
        dist.broadcast(probe, src=world - 1)
        if rank == 0:
            xs = shard_samples(last.sample0, last.sample0 + (k - 1) * (nfft - nov) + nfft)
            ref = iqw.spectrogram(xs, fs=100e6, window='blackmanharris', nperseg=nfft, noverlap=nov, dB=True,
                                  return_axis_arrays=False)
            parity = bool(torch.equal(ref, probe))
            del xs, ref
        del mine, probe
    row('e1 configs[1] spectrogram dB nfft 2048 blackmanharris 50 %, time-sharded with halo', n, ms, 16,
        halo_samples_per_shard=nov, collective='none', shard_parity=parity,
        limiter='kernel 1 on each shard (same kernel as one GPU); no exchange')
    del x
    torch.cuda.empty_cache()

    # ---- e3: one config-3 channel split in time, exact quantiles -------------------------------------------
    n = 1_000_000_000 // scale
    kw = dict(fs=100e6, window='hann', resolution=100e6 / 4096, fractional_overlap=0.5,
              statistics=[0.1, 0.5, 0.9, 0.999], dB=True)
    xfull = bench.device_capture(torch, n, 1234, dev)
    sh = D.frame_shard(n, 4096, 2048, world, rank)
    mine = xfull[sh.sample0:sh.sample1].clone()
    single, t_single = None, None
    if rank == 0:
        single = iqw.persistence_spectrum(xfull, axis=0, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); single = iqw.persistence_spectrum(xfull, axis=0, **kw); e1.record(); torch.cuda.synchronize()
        t_single = e0.elapsed_time(e1)
    del xfull
    torch.cuda.empty_cache()
    ms = timed(lambda: D.persistence_spectrum_time_sharded(mine, n_samples=n, **kw))
    got = D.persistence_spectrum_time_sharded(mine, n_samples=n, **kw)
    eq = bool(torch.equal(got, single)) if rank == 0 else None
    row('e3 configs[2] one channel time-sharded, exact quantiles [0.1, 0.5, 0.9, 0.999]', n, ms, 24,
        single_gpu_ms=round(t_single, 4) if t_single else None, equal_to_single_gpu=eq,
        collective='nccl all_reduce: 2 x bracket keys (4, 4096) int64, 1 x below counts, 4 x digit counts (4, 4096, 256) int32',
        limiter='local kernel-2 run for the brackets + collect pass + 4 x 16 MB all_reduce (DESIGN section 6)')
    del mine, got, single
    torch.cuda.empty_cache()

    # ---- e4: configs[3] bin power, bin-aligned shards -------------------------------------------------------
    bin_len = 245_760
    n = bin_len * (60_000 // scale)
    bs = D.bin_shard(n, bin_len, world, rank)
    x = shard_samples(bs.sample0, bs.sample1, seed=11)
    for kind in ('mean', 'peak'):
        ms = timed(lambda: D.iq_to_bin_power_sharded(x, 1 / 245.76e6, 1e-3, n_samples=n, kind=kind))
        parity = None
        if world > 1:
            allbins = D.iq_to_bin_power_sharded(x, 1 / 245.76e6, 1e-3, n_samples=n, kind=kind)
            if rank == 0:
                last = D.bin_shard(n, bin_len, world, world - 1)
                # 5000 bins: at least 32 bins per SM, so the recomputation takes the same one-CTA-per-bin
                # summation order as the shard (kernel 3 splits a bin over several CTAs only when there are
                # fewer bins than that; 'mean' then differs in the last bit, 'peak' never does)
                k = min(5000, last.bin1 - last.bin0)
                xs = shard_samples(last.sample0, last.sample0 + k * bin_len, seed=11)
                ref = iqw.iq_to_bin_power(xs, 1 / 245.76e6, 1e-3, kind=kind)
                parity = bool(torch.equal(ref, allbins[last.bin0:last.bin0 + k]))
                del xs
        row(f'e4 configs[3] iq_to_bin_power {kind} 1 ms bins, 245.76 MS/s x {n / 245.76e6:.0f} s, bin-sharded', n, ms, 8,
            bins_per_shard=bs.bin1 - bs.bin0, collective='nccl all_gather of the per-rank bin rows (fp32)',
            shard_parity=parity, limiter='kernel 3 on each shard (HBM); the gather moves 240 KB')
    del x
    clocks = sampler.stop() if sampler else None
    if rank == 0 and a.out:
        json.dump({'hbm_peak_GBps': peak, 'n_gpus': world, 'clocks': clocks, 'rows': rows}, open(a.out, 'w'), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
