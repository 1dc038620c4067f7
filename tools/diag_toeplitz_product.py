"""Diagnostic: frx_toeplitz_product against a dense numpy product (incl. the rectangular n_cols mode)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched
from scipy.linalg import toeplitz
rng = np.random.default_rng(3)
for N, n_cols in ((700, 0), (2048, 0), (2048, 1024), (3000, 1024), (5000, 2048), (5000, 4096), (9000, 1024), (9000, 8192)):
    for n_seq, paired in ((1, False), (3, True)):
        c = rng.standard_normal((n_seq, N)) / (1 + np.arange(N))
        g = rng.standard_normal((n_seq, 1, N) if paired else (2, N))
        if n_cols:
            g[..., n_cols:] = 0.0
        y = batched.toeplitz_product(torch.as_tensor(c, device='cuda'), torch.as_tensor(g, device='cuda'), paired=paired,
                                     scale=-1.0, n_cols=n_cols).cpu().numpy()
        worst = 0.0
        for s in range(n_seq):
            T = np.tril(toeplitz(c[s]))
            fields = g[s] if paired else g
            for f in range(fields.shape[0]):
                ref = -(T @ fields[f])
                d = np.abs(y[s, f] - ref)[n_cols:]
                worst = max(worst, d.max() / np.abs(ref).max())
                if d.max() / np.abs(ref).max() > 1e-12:
                    bad = np.nonzero(d > 1e-12 * np.abs(ref).max())[0] + n_cols
                    print('   BAD rows', bad[:5], '...', bad[-5:], 'count', len(bad))
        print('N', N, 'n_cols', n_cols, 'n_seq', n_seq, 'paired', paired, 'max rel err', worst)
