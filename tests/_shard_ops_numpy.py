"""numpy stand-in for iqwaveform_b200.distributed.CudaShardOps: the CPU suite has no GPU, so the
per-rank device work of the time-sharded persistence spectrum is replaced by the oracle's
spectrogram plus a direct numpy statement of the radix count / descend / finish steps.  The host
logic under test (shard planning, the four count -> all_reduce -> descend rounds over gloo, the
row bookkeeping) is the product's own."""
import numpy as np
import torch

from oracle import iqw_oracle as orc


def float_to_key(v: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(v, dtype=np.float32).view(np.uint32)
    return b ^ np.where(b & np.uint32(0x80000000), np.uint32(0xFFFFFFFF), np.uint32(0x80000000))


def key_to_float(k: np.ndarray) -> np.ndarray:
    k = np.ascontiguousarray(k, dtype=np.uint32)
    b = k ^ np.where(k & np.uint32(0x80000000), np.uint32(0x80000000), np.uint32(0xFFFFFFFF))
    return b.view(np.float32)


class NumpyShardOps:
    def power_spectrogram(self, x, *, window, nfft, noverlap, nzero, bin_lo, bin_hi):
        x = np.asarray(x)
        if x.size == 0:
            return torch.empty((0, bin_hi - bin_lo), dtype=torch.float32)
        p = orc.spectrogram(x, fs=1.0, window=window, nperseg=nfft, noverlap=noverlap, nzero=nzero, axis=0,
                            return_axis_arrays=False)
        return torch.from_numpy(np.ascontiguousarray(p[:, bin_lo:bin_hi], dtype=np.float32))

    def named_statistics(self, p, names, dB):
        a = p.numpy()
        a = orc.powtodB(a.copy(), eps=1e-25) if dB else a
        return torch.from_numpy(np.stack([orc._named_stat(n)(a, axis=0) for n in names]).astype(np.float32))

    def local_order_statistics(self, p, local_ranks):
        k = np.sort(float_to_key(p.numpy()), axis=0)
        return torch.from_numpy(key_to_float(k[list(local_ranks)]))

    CAP = 200       # candidates a rank may keep per column before it reports an overflow
    collected = overflowed = False

    def bracket_collect(self, p, lo, hi):
        k = float_to_key(p.numpy())
        lo_u, hi_u = lo.numpy().view(np.uint32), hi.numpy().view(np.uint32)
        n_sel, nb = lo.shape
        below = np.zeros((n_sel + 1, nb), dtype=np.int32)
        keep = np.zeros(k.shape, dtype=bool)
        for s in range(n_sel):
            below[s] = (k < lo_u[s][None, :]).sum(0)
            keep |= (k >= lo_u[s][None, :]) & (k <= hi_u[s][None, :])
        below[n_sel] = keep.sum(0) > self.CAP
        self.collected = True
        self.overflowed |= bool(below[n_sel].any())
        store = torch.from_numpy(np.where(keep, p.numpy(), np.float32(np.nan))), torch.from_numpy(keep)
        return store, torch.from_numpy(below)

    def candidate_count(self, store, lo, hi, level):
        vals, keep = store
        k = float_to_key(vals.numpy())
        lo_u, hi_u = lo.numpy().view(np.uint32), hi.numpy().view(np.uint32)
        digit = ((k >> np.uint32(24 - 8 * level)) & np.uint32(255)).astype(np.int64)
        n_sel, nb = lo.shape
        counts = np.zeros((n_sel, nb, 256), dtype=np.int32)
        cols = np.broadcast_to(np.arange(nb), k.shape)
        for s in range(n_sel):
            m = keep.numpy() & (k >= lo_u[s][None, :]) & (k <= hi_u[s][None, :])
            np.add.at(counts[s], (cols[m], digit[m]), 1)
        return torch.from_numpy(counts)

    def radix_count(self, p, lo, hi, level, want_below):
        k = float_to_key(p.numpy())
        lo_u, hi_u = lo.numpy().view(np.uint32), hi.numpy().view(np.uint32)
        digit = ((k >> np.uint32(24 - 8 * level)) & np.uint32(255)).astype(np.int64)
        n_sel, nb = lo.shape
        counts = np.zeros((n_sel, nb, 256), dtype=np.int32)
        below = np.zeros((n_sel, nb), dtype=np.int32)
        cols = np.broadcast_to(np.arange(nb), k.shape)
        for s in range(n_sel):
            m = (k >= lo_u[s][None, :]) & (k <= hi_u[s][None, :])
            np.add.at(counts[s], (cols[m], digit[m]), 1)
            below[s] = (k < lo_u[s][None, :]).sum(0)
        return torch.from_numpy(counts), (torch.from_numpy(below) if want_below else None)

    def radix_descend(self, counts, rank, prefix, lo, hi, level):
        c = counts.numpy().astype(np.int64)
        r = rank.numpy()
        pre = prefix.numpy().view(np.uint32)
        lo_u, hi_u = lo.numpy().view(np.uint32), hi.numpy().view(np.uint32)
        cum = np.cumsum(c, axis=-1)
        hit = cum > r[..., None]
        digit = hit.argmax(-1)
        assert hit.any(-1).all() and (r >= 0).all()
        below = np.take_along_axis(cum - c, digit[..., None], -1)[..., 0]
        pre[...] = digit.astype(np.uint32) if level == 0 else ((pre << np.uint32(8)) | digit.astype(np.uint32))
        r[...] = r - below
        shift = 24 - 8 * level
        b0 = pre << np.uint32(shift)
        b1 = b0 | np.uint32((1 << shift) - 1)
        lo_u[...] = np.maximum(lo_u, b0)
        hi_u[...] = np.minimum(hi_u, b1)

    def finish(self, keys, sel_rank, n_rows_total, reqs, dB):
        v = key_to_float(keys.numpy().view(np.uint32))
        v = orc.powtodB(v.copy(), eps=1e-25) if dB else v
        rows = []
        for q in reqs:
            lo, hi = (((n_rows_total - 1) // 2, n_rows_total // 2) if q.kind == 4 else (q.rank_lo, q.rank_hi))
            a, b = v[sel_rank.index(lo)], v[sel_rank.index(hi)]
            if q.kind == 4:
                rows.append((a + b) * np.float32(0.5))
                continue
            g = np.float32(q.gamma)
            d = b - a
            rows.append(b - d * (np.float32(1) - g) if g >= 0.5 else a + d * g)
        return torch.from_numpy(np.stack(rows).astype(np.float32))
