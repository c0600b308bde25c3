"""one-process-per-GPU sharding of the hot path (no counterpart in the reference, which is
single-process: SURVEY.md section 2.1 / 8e).

The path shards without any collective on the data path:

* channels are independent                     -> `channel_shard`  (persistence spectrum, config 3)
* STFT frames are independent                  -> `frame_shard`    (stft / spectrogram, config 2):
  rank r owns frames [f0, f1) and reads samples [f0*hop, (f1-1)*hop + nfft): the last
  `noverlap` samples are the halo shared with the next rank
* power bins are independent                   -> `bin_shard`      (iq_to_bin_power, config 4):
  shards are aligned to whole bins, so the halo is zero

Only the small results are exchanged: `gather_rows` is one `all_gather` over NCCL (CUDA tensors)
or gloo (CPU tensors, used by the CPU tests) of the per-rank result rows.
"""
from __future__ import annotations

from typing import Callable, NamedTuple

import torch
import torch.distributed as dist

__all__ = ['channel_shard', 'frame_shard', 'bin_shard', 'gather_rows', 'persistence_spectrum_sharded',
           'persistence_spectrum_time_sharded', 'spectrogram_time_sharded', 'iq_to_bin_power_sharded']


def _split(n: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous, balanced split of range(n): the first n % world ranks get one extra item"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f'rank {rank} outside world of {world}')
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def channel_shard(n_channels: int, world: int, rank: int) -> range:
    return range(*_split(n_channels, world, rank))


class FrameShard(NamedTuple):
    frame0: int     # first frame of this rank
    frame1: int     # one past its last frame
    sample0: int    # first sample it reads
    sample1: int    # one past the last sample it reads (includes the halo)
    n_frames: int   # frames of the whole capture


def frame_shard(n_samples: int, nfft: int, noverlap: int, world: int, rank: int) -> FrameShard:
    """time-axis shard of an STFT with the reference's framing (fourier.py:568-569, 1016-1028:
    T = (N - nfft)//hop + 1 with overlap, N//nfft without; no padding, partial tail dropped)"""
    hop = nfft - noverlap
    if hop < 1 or nfft < 1:
        raise ValueError('need 0 <= noverlap < nfft')
    if noverlap == 0:
        T = n_samples // nfft
    else:
        T = (n_samples - nfft) // hop + 1 if n_samples >= nfft else 0
    f0, f1 = _split(T, world, rank)
    if f1 == f0:
        return FrameShard(f0, f1, f0 * hop, f0 * hop, T)
    return FrameShard(f0, f1, f0 * hop, (f1 - 1) * hop + nfft, T)


class BinShard(NamedTuple):
    bin0: int
    bin1: int
    sample0: int
    sample1: int
    n_bins: int


def bin_shard(n_samples: int, bin_len: int, world: int, rank: int) -> BinShard:
    """bin-aligned time shard for iq_to_bin_power (power_analysis.py:380: whole bins only)"""
    if bin_len < 1:
        raise ValueError('bin length must be >= 1 sample')
    n_bins = n_samples // bin_len
    b0, b1 = _split(n_bins, world, rank)
    return BinShard(b0, b1, b0 * bin_len, b1 * bin_len, n_bins)


def gather_rows(local: torch.Tensor, sizes: list[int] | None = None, axis: int = 0, group=None) -> torch.Tensor:
    """all_gather of per-rank result rows along `axis`.  `sizes[r]` = rows of rank r when shards are
    ragged (default: every rank has local.shape[axis] rows).  Works on the backend of the group:
    NCCL for CUDA tensors, gloo for CPU tensors."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    local = local.movedim(axis, 0).contiguous()
    if sizes is None:
        sizes = [local.shape[0]] * world
    if len(sizes) != world or sizes[dist.get_rank(group)] != local.shape[0]:
        raise ValueError('sizes must list the row count of every rank')
    rest = tuple(local.shape[1:])
    # equal-sized exchange (padded to the largest shard): one collective, KB-MB payloads
    m = max(sizes)
    send = local
    if local.shape[0] < m:
        send = torch.zeros((m,) + rest, dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
    recv = [torch.empty((m,) + rest, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    out = torch.cat([r[:n] for r, n in zip(recv, sizes)], dim=0)
    return out.movedim(0, axis)


def _world(group=None) -> tuple[int, int]:
    if dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def persistence_spectrum_sharded(x_local, *, n_channels: int, group=None, compute: Callable | None = None, **kw):
    """channel-sharded persistence spectrum: `x_local` holds this rank's channels
    (`channel_shard(n_channels, world, rank)`, shape (C_local, N)); every rank returns the full
    (n_channels, nstat, nbins) result.  `compute` defaults to the CUDA `persistence_spectrum`."""
    if compute is None:
        from .fourier import persistence_spectrum as compute
    world, rank = _world(group)
    mine = channel_shard(n_channels, world, rank)
    if x_local.shape[0] != len(mine):
        raise ValueError(f'rank {rank} expects {len(mine)} channels, got {x_local.shape[0]}')
    kw = dict(kw, axis=1)
    out = compute(x_local, **kw) if len(mine) else None
    if world == 1:
        return out
    if out is None:     # more ranks than channels: contribute an empty shard of the right row shape
        raise ValueError('every rank needs at least one channel (use a sub-group otherwise)')
    out = torch.as_tensor(out)
    return gather_rows(out, [len(channel_shard(n_channels, world, r)) for r in range(world)], 0, group)


_REDUCIBLE = {'mean': 'sum', 'rms': 'sum', 'max': 'max', 'peak': 'max', 'min': 'min'}


def persistence_spectrum_time_sharded(x_halo, *, n_samples: int, fs: float, resolution: float,
                                      fractional_overlap=0, statistics, group=None,
                                      compute: Callable | None = None, **kw):
    """time-sharded persistence spectrum of ONE long 1-D capture for the statistics that combine
    with a single exchange: 'mean'/'rms' (frame-count-weighted all_reduce SUM of the per-rank
    means, taken over dB values when dB=True, like the reference), 'max'/'peak', 'min'
    (all_reduce MAX / MIN).  `x_halo` holds this rank's samples [shard.sample0, shard.sample1).
    Exact quantiles across time shards would need a candidate exchange per column and are not
    built: shard by channel instead (`persistence_spectrum_sharded`)."""
    for s in statistics:
        if s not in _REDUCIBLE:
            raise NotImplementedError(
                f'statistic {s!r}: only mean/rms/max/peak/min combine across time shards; shard by channel')
    if compute is None:
        from .fourier import persistence_spectrum as compute
    world, rank = _world(group)
    nfft = round(fs / resolution)
    noverlap = round(fractional_overlap * nfft)
    sh = frame_shard(n_samples, nfft, noverlap, world, rank)
    if x_halo.shape[-1] != sh.sample1 - sh.sample0:
        raise ValueError(f'rank {rank} expects {sh.sample1 - sh.sample0} samples, got {x_halo.shape[-1]}')
    local = torch.as_tensor(compute(x_halo, fs=fs, resolution=resolution, fractional_overlap=fractional_overlap,
                                    statistics=list(statistics), axis=0, **kw)).clone()      # (nstat, nbins)
    if world == 1:
        return local
    n_local = sh.frame1 - sh.frame0
    for i, s in enumerate(statistics):
        row = local[i]
        if _REDUCIBLE[s] == 'sum':
            row.mul_(n_local / sh.n_frames)
            dist.all_reduce(row, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(row, op=dist.ReduceOp.MAX if _REDUCIBLE[s] == 'max' else dist.ReduceOp.MIN, group=group)
    return local


def spectrogram_time_sharded(x_halo, *, n_samples: int, nperseg: int, noverlap: int = 0, group=None,
                             compute: Callable | None = None, gather: bool = False, **kw):
    """time-sharded spectrogram of a 1-D capture of `n_samples` samples: `x_halo` holds this rank's
    samples [shard.sample0, shard.sample1) (halo included).  Returns this rank's frames
    (T_local, nfft) or, with gather=True, all frames on every rank."""
    if compute is None:
        from .fourier import spectrogram as compute
    world, rank = _world(group)
    sh = frame_shard(n_samples, nperseg, noverlap, world, rank)
    if x_halo.shape[-1] != sh.sample1 - sh.sample0:
        raise ValueError(f'rank {rank} expects {sh.sample1 - sh.sample0} samples, got {x_halo.shape[-1]}')
    p = compute(x_halo, nperseg=nperseg, noverlap=noverlap, axis=0, return_axis_arrays=False, **kw)
    p = torch.as_tensor(p)
    assert p.shape[0] == sh.frame1 - sh.frame0
    if not gather or world == 1:
        return p
    sizes = [(lambda s: s.frame1 - s.frame0)(frame_shard(n_samples, nperseg, noverlap, world, r)) for r in range(world)]
    return gather_rows(p, sizes, 0, group)


def iq_to_bin_power_sharded(x_local, Ts: float, Tbin: float, *, n_samples: int, kind='mean', group=None,
                            compute: Callable | None = None):
    """bin-sharded iq_to_bin_power of a 1-D capture: `x_local` holds samples
    [shard.sample0, shard.sample1); every rank returns all n_bins values."""
    if compute is None:
        from .power_analysis import iq_to_bin_power as compute
    world, rank = _world(group)
    bin_len = round(Tbin / Ts)
    sh = bin_shard(n_samples, bin_len, world, rank)
    if x_local.shape[-1] != sh.sample1 - sh.sample0:
        raise ValueError(f'rank {rank} expects {sh.sample1 - sh.sample0} samples, got {x_local.shape[-1]}')
    out = torch.as_tensor(compute(x_local, Ts, Tbin, kind=kind, axis=0, truncate=True))
    if world == 1:
        return out
    sizes = [(lambda s: s.bin1 - s.bin0)(bin_shard(n_samples, bin_len, world, r)) for r in range(world)]
    return gather_rows(out, sizes, 0, group)
