"""host logic of the distributed order-statistic select (iqwaveform_b200.distributed
.select_order_statistics) on CPU: the shards are threads of this process joined by ThreadGroup, the
per-rank device work is the numpy stand-in.  Checks the bracket guarantee, the grouping of
neighbouring ranks, chunks of more than 8 statistics, empty shards, and the fall-back to whole-matrix
counting when ties overflow the candidate store."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

from iqwaveform_b200 import distributed as D
from _shard_ops_numpy import NumpyShardOps, float_to_key


def run_sharded(p, cuts, sel, **kw):
    edges = [0] + list(cuts) + [p.shape[0]]
    shards = [p[a:b].contiguous() for a, b in zip(edges[:-1], edges[1:])]
    tg = D.ThreadGroup(len(shards))
    infos = [{} for _ in shards]
    with ThreadPoolExecutor(len(shards)) as ex:
        keys = list(ex.map(lambda r: D.select_order_statistics(shards[r], sel, p.shape[0], group=tg.member(r),
                                                               ops=NumpyShardOps(), info=infos[r], **kw),
                           range(len(shards))))
    for k in keys[1:]:
        assert torch.equal(k, keys[0])
    return keys[0].numpy().view(np.uint32), infos[0]


def test_rank_groups():
    assert D._rank_groups([0, 1, 5, 6, 7, 20]) == [[0, 1], [2, 3, 4], [5]]
    assert D._rank_groups([3]) == [[0]]
    assert D._rank_groups([5, 2, 3]) == [[0], [1, 2]]


@pytest.mark.parametrize('cuts', [[], [1000], [700, 700, 2999], [1, 2, 3, 4, 5, 6, 7]])
@pytest.mark.parametrize('bracket', [True, False])
def test_select_equals_sort(cuts, bracket):
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((3000, 19)) ** 2).astype(np.float32)
    a[::5, 3] = 0.0
    a[1::5, 3] = -0.0
    a[:, 4] = np.float32(2.5)                    # a constant column
    a[::3, 5] = np.inf
    a[:, 6] = -a[:, 6]
    sel = [0, 1, 299, 1500, 1501, 2996, 2997, 2998, 2999]      # 9 statistics: two passes
    got, info = run_sharded(torch.from_numpy(a), cuts, sel, bracket=bracket)
    assert np.array_equal(got, np.sort(float_to_key(a), axis=0)[sel])
    if bracket:
        assert info['candidate_store'] is False   # the constant column overflows every store


def test_stationary_noise_stays_on_the_candidate_store():
    rng = np.random.default_rng(2)
    a = (rng.standard_normal((6000, 8)) ** 2).astype(np.float32)
    sel = [600, 601, 3000, 5994]
    got, info = run_sharded(torch.from_numpy(a), [2000, 4000], sel)
    assert info['candidate_store'] is True
    assert np.array_equal(got, np.sort(float_to_key(a), axis=0)[sel])


def test_requires_total_rows_when_sharded():
    tg = D.ThreadGroup(2)
    with pytest.raises(ValueError):
        D.select_order_statistics(torch.zeros(4, 2), [1], group=tg.member(0), ops=NumpyShardOps())


def test_numa_binding_is_best_effort():
    import os
    before = os.sched_getaffinity(0)
    got = D.bind_to_gpu_numa_node(0)          # no GPU here: must not raise and must not change anything
    assert got is None or isinstance(got, str)
    if got is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
