"""exact order statistics of a row-split matrix (iqw_radix_count_f32 / iqw_radix_descend /
iqw_order_stats_finish_f32): the four-level radix select equals a full sort, and the rows it
produces equal iqw_time_stats_f32 on the concatenated matrix bit for bit.  Several shards are
emulated on one GPU by summing their counts where the all_reduce would."""
import numpy as np
import pytest
import torch

import iqwaveform_b200 as iqw
from iqwaveform_b200 import _lib, _plan
from iqwaveform_b200 import distributed as D
from iqwaveform_b200.fourier import time_statistics
from _shard_ops_numpy import float_to_key

pytestmark = pytest.mark.gpu


def virtual_select(shards, sel, T, bracket=True, info=None):
    """select_order_statistics with one thread per shard on this GPU (ThreadGroup stands in for the
    NCCL group); every 'rank' must end with the same keys"""
    from concurrent.futures import ThreadPoolExecutor
    tg = D.ThreadGroup(len(shards))
    infos = [{} for _ in shards]

    def run(r):
        torch.cuda.set_device(0)
        return D.select_order_statistics(shards[r], sel, T, group=tg.member(r), bracket=bracket, info=infos[r])

    with ThreadPoolExecutor(len(shards)) as ex:
        keys = list(ex.map(run, range(len(shards))))
    torch.cuda.synchronize()
    for k in keys[1:]:
        assert torch.equal(k, keys[0])
    if info is not None:
        info.update(infos[0])
    return keys[0]


def _matrix(kind, T, nb, seed):
    g = torch.Generator('cuda').manual_seed(seed)
    if kind == 'power':
        return torch.randn(T, nb, device='cuda', generator=g).square_() * 1e-3
    if kind == 'signed':
        return torch.randn(T, nb, device='cuda', generator=g)
    if kind == 'ties':
        return torch.randint(0, 7, (T, nb), device='cuda', generator=g).float() * 0.25
    if kind == 'specials':
        p = torch.randn(T, nb, device='cuda', generator=g)
        p[::7, ::3] = 0.0
        p[1::11, 1::5] = -0.0
        p[2::13] = float('inf')
        p[3::17, ::2] = float('-inf')
        p[5::19, 1::4] = 1e-42           # denormal
        return p
    raise AssertionError(kind)


@pytest.mark.parametrize('kind,T,nb,cuts', [
    ('power', 5000, 257, [1200, 1200, 4999]),        # an empty shard and a one-row shard
    ('signed', 3001, 64, [1000, 2000]),
    ('ties', 4096, 100, [17]),
    ('specials', 2500, 33, [800, 1700]),
    ('power', 1, 40, []),
])
def test_radix_select_equals_sort(kind, T, nb, cuts):
    p = _matrix(kind, T, nb, 5)
    edges = [0] + cuts + [T]
    shards = [p[a:b].contiguous() for a, b in zip(edges[:-1], edges[1:])]
    ranks = sorted({0, T - 1, T // 2, (T - 1) // 2, min(T - 1, 3), max(0, T - 2), T // 10})
    srt = np.sort(float_to_key(p.cpu().numpy()), axis=0)      # key order == the library's total order
    for bracket in (True, False):
        info = {}
        keys = virtual_select(shards, ranks, T, bracket, info)
        assert np.array_equal(keys.cpu().numpy().view(np.uint32), srt[ranks]), bracket
        if bracket and kind != 'power':
            # ties and whole rows of infinities overflow the candidate store, noise in balanced shards
            # does not (the one-row shard of the 'power' case makes its brackets as wide as the data)
            assert info['candidate_store'] is (kind == 'signed'), kind
    keys = D.select_order_statistics(p, ranks)                # one rank holding every row
    assert np.array_equal(keys.cpu().numpy().view(np.uint32), srt[ranks])


@pytest.mark.parametrize('dB', [True, False])
def test_rowsplit_rows_equal_time_stats(dB):
    T, nb = 20000, 300
    p = _matrix('power', T, nb, 9)
    stats = [0.1, 0.5, 'median', 0.999, 1.0, 0.0, 0.37]
    want = time_statistics(p[None], stats, dB=dB)[0]
    reqs = _plan.stat_requests(stats, T)
    sel = sorted(_plan.distinct_ranks(reqs, T))
    shards = [p[:7000].contiguous(), p[7000:7001].contiguous(), p[7001:].contiguous()]
    got = D.CudaShardOps().finish(virtual_select(shards, sel, T), sel, T, reqs, dB)
    assert torch.equal(got, want)


def test_rank_outside_rows_is_poisoned_not_wrong():
    p = _matrix('power', 100, 40, 3)
    keys = D.select_order_statistics(p, [5, 100], bracket=False)
    k = keys.cpu().numpy().view(np.uint32)
    assert (k[1] == 0xFFFFFFFF).all()
    assert np.array_equal(k[0], np.sort(float_to_key(p.cpu().numpy()), axis=0)[5])


def test_rowsplit_argument_errors():
    p = torch.zeros(4, 8, device='cuda')
    ops = D.CudaShardOps()
    z = torch.zeros((2, 8), dtype=torch.int32, device='cuda')
    with pytest.raises(ValueError):
        ops.radix_count(p, z, z, 4, False)
    keys = torch.zeros((1, 8), dtype=torch.int32, device='cuda')
    with pytest.raises(ValueError):          # rank 2 was not selected
        ops.finish(keys, [1], 4, _plan.stat_requests([0.5], 4), True)


def test_time_sharded_single_process_equals_persistence_spectrum():
    """world size 1: the radix-select path end to end against kernel 2"""
    from oracle.make_golden import synth
    x = torch.from_numpy(synth(31, (200000,))).cuda()
    kw = dict(fs=1e6, window='hann', resolution=1e6 / 512, fractional_overlap=0.5, dB=True,
              statistics=[0.1, 'mean', 0.5, 'max', 0.999, 'min', 'median'], bandwidth=0.6e6)
    n_used = D.frame_shard(x.numel(), 512, 256, 1, 0).sample1     # the partial last frame is not part of the shard
    got = D.persistence_spectrum_time_sharded(x[:n_used], n_samples=x.numel(), **kw)
    want = iqw.persistence_spectrum(x, axis=0, **kw)
    assert torch.equal(got, want)
