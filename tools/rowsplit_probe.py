#!/usr/bin/env python
"""time the bracketed radix select of the time-sharded persistence spectrum on one GPU:
python tools/rowsplit_probe.py [rows] [cols] [virtual_ranks]
(default: one config-3 channel, 488280 x 4096, as if split over 2 ranks: the bracket comes from the
local order statistics of the two halves, the counting passes then run over ALL rows on this GPU,
i.e. the time printed is what `virtual_ranks` GPUs would each spend on rows/virtual_ranks rows,
times virtual_ranks)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iqwaveform_b200 import _lib, _plan
from iqwaveform_b200 import distributed as D
from iqwaveform_b200.fourier import time_statistics

T = int(sys.argv[1]) if len(sys.argv) > 1 else 488280
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
R = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p = torch.empty(T, nb, device='cuda')
g = torch.Generator('cuda').manual_seed(1)
scale = 1 + 100 * torch.rand(nb, device='cuda', generator=g)
for i in range(0, T, 65536):
    p[i:i + 65536] = torch.randn(p[i:i + 65536].shape, device='cuda', generator=g).square_() * scale
reqs = _plan.stat_requests([0.1, 0.5, 0.9, 0.999], T)
sel = sorted(_plan.distinct_ranks(reqs, T))
ops = D.CudaShardOps()
edges = [T * r // R for r in range(R + 1)]
for it in range(3):
    _lib.profile(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    vals = [ops.local_order_statistics(p[a:b], [min(j * (b - a) // T, b - a - 1) for j in sel])
            for a, b in zip(edges[:-1], edges[1:])]
    klo = torch.stack([D._keys_of(v) for v in vals]).amin(0)
    khi = torch.stack([D._keys_of(v) for v in vals]).amax(0)
    grp = D._rank_groups(sel)
    glo = D._as_u32_storage(torch.stack([klo[g].amin(0) for g in grp]))
    ghi = D._as_u32_storage(torch.stack([khi[g].amax(0) for g in grp]))
    of = torch.tensor([k for k, g in enumerate(grp) for _ in g], device='cuda')
    lo, hi = glo[of].contiguous(), ghi[of].contiguous()
    ev[1].record()
    rank = torch.tensor(sel, dtype=torch.int64, device='cuda').reshape(-1, 1).expand(len(sel), nb).contiguous()
    prefix = torch.zeros((len(sel), nb), dtype=torch.int32, device='cuda')
    use_store = it < 2
    if use_store:
        store, below = ops.bracket_collect(p, glo, ghi)
        rank -= below[of]
        assert not bool(below[len(grp)].any())
    inside = None
    for level in range(4):
        if use_store:
            counts = ops.candidate_count(store, lo, hi, level)
        else:
            counts, below = ops.radix_count(p, lo, hi, level, level == 0)
            if level == 0:
                rank -= below
        if level == 0:
            inside = counts.sum(-1).float().mean().item()
        ops.radix_descend(counts, rank, prefix, lo, hi, level)
    out = ops.finish(prefix, sel, T, reqs, True)
    ev[2].record(); torch.cuda.synchronize()
    rep = {k: (n, round(ms, 3)) for k, (n, ms) in _lib.profile_report().items()
           if k.startswith(('radix', 'order', 'bracket_collect', 'candidate'))}
    print(f'rows {T} cols {nb} sel {len(sel)} ranks {R}: brackets {ev[0].elapsed_time(ev[1]):.3f} ms, '
          f'select {ev[1].elapsed_time(ev[2]):.3f} ms, rows inside a bracket {inside:.0f}', rep)
want = time_statistics(p[None], [0.1, 0.5, 0.9, 0.999], dB=True)[0]
print('equal to kernel 2:', torch.equal(out, want))
