import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iqwaveform_b200 as iqw
dev = torch.device('cuda:0')
T, nb = 70001, 257
rng = np.random.default_rng(T)
a = rng.standard_normal((2, T, nb)).astype(np.float32)
a[0, :, 0] = 1.5
a[0, :, 1] = np.round(a[0, :, 1])
a[1, 3, 2], a[1, 4, 2] = 1e30, -1e30
a[1, :, 3 % nb] = np.where(np.arange(T) % 2 == 0, 0.0, -0.0)
a[1, :, 4] = np.exp(6 * a[1, :, 4])
a[0, :, 4] = np.arange(T, dtype=np.float32)
for qs in ([0.0, 0.1, 0.5, 0.999], [0.5], [0.1, 0.5], [0.5, 0.999]):
    got = iqw.time_statistics(torch.from_numpy(a).to(dev), qs, dB=False).cpu().numpy()
    want = np.quantile(a, np.array(qs, dtype=np.float32), axis=1)
    for i, q in enumerate(qs):
        bad = np.argwhere(got[:, i] != want[i])
        for c, col in bad[:6]:
            s = np.sort(a[c, :, col])
            print('qs', qs, 'q', q, 'ch', c, 'col', col, 'got', got[c, i, col], 'want', want[i][c, col], 'rank of got in sorted:', np.searchsorted(s, got[c, i, col]), 'want rank', (T-1)*q)
        print(qs, q, 'n bad', len(bad), 'of', 2*nb)
