#!/usr/bin/env python
"""ONE config-3 channel (1e9 samples, nfft 4096, 50 % overlap, 4 exact quantiles) split in time over
the GPUs of a box: persistence_spectrum_time_sharded against the single-GPU persistence_spectrum.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/timeshard_probe.py [n_samples]

Every rank generates the same capture (same seed) and keeps its shard + halo; rank 0 also runs the
whole capture alone.  Times are CUDA events on each rank, max over ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
import iqwaveform_b200 as iqw
from iqwaveform_b200 import distributed as D

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
kw = dict(fs=100e6, window='hann', resolution=100e6 / 4096, fractional_overlap=0.5,
          statistics=[0.1, 0.5, 0.9, 0.999], dB=True)
x = bench.device_capture(torch, n, 1234, dev)
sh = D.frame_shard(n, 4096, 2048, world, rank)
mine = x[sh.sample0:sh.sample1].clone()
single = None
t_single = None
if rank == 0:
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); single = iqw.persistence_spectrum(x, axis=0, **kw); e1.record(); torch.cuda.synchronize()
        t_single = e0.elapsed_time(e1)
del x
torch.cuda.empty_cache()
times = []
for _ in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); got = D.persistence_spectrum_time_sharded(mine, n_samples=n, **kw); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(t.item())
if rank == 0:
    print(json.dumps({'n_samples': n, 'n_gpus': world, 'single_gpu_ms': round(t_single, 3),
                      'time_sharded_ms': [round(t, 3) for t in times],
                      'GS_per_s_time_sharded': round(n / min(times[1:]) / 1e6, 2),
                      'equal_to_single_gpu': bool(torch.equal(got, single))}))
if world > 1:
    dist.destroy_process_group()
