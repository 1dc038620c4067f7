"""Stress aid: repeated launches of the history kernel (several sizes, rules, batch shapes) must give the same bits
every time -- catches races in the staging / epilogue / split-tile fix-up paths."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched
rng = np.random.default_rng(1)
flush = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
bad = 0
for rule, N, n_alpha, n_fields, reps in (('caputo', 100000, 1, 1, 150), ('caputo', 100000, 3, 2, 30), ('simpson', 100000, 1, 1, 60),
                                         ('trapezoidal', 20000, 5, 3, 60), ('caputo', 1024, 7, 5, 200), ('caputo', 1025, 1, 1, 200),
                                         ('rectangle', 3000, 2, 2, 200), ('caputo', 5, 3, 1, 100), ('grunwald_letnikov', 50000, 2, 1, 40)):
    h = 1.0 / N
    g = torch.as_tensor(rng.standard_normal((n_fields, N + 2)), device='cuda')
    alphas = np.linspace(0.15, 0.85, n_alpha)
    ref = batched.history(rule, alphas, h, g, N=N).clone()
    n_bad = 0
    for r in range(reps):
        if r % 3 == 0:
            flush.zero_()
        y = batched.history(rule, alphas, h, g, N=N)
        if not torch.equal(torch.nan_to_num(y, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)):
            n_bad += 1
    # the host entry point (pinned, in place) gives the same bits
    gp = batched.pinned_empty((n_fields, N + 2)); gp[:] = g.cpu().numpy()
    out = batched.pinned_empty((n_alpha, n_fields, N + 1))
    for r in range(10):
        out[:] = 0.0
        yh = batched.history_host(rule, alphas, h, gp, N=N, out=out)
        if not np.array_equal(np.nan_to_num(yh, nan=-7.0), np.nan_to_num(ref.cpu().numpy(), nan=-7.0)):
            n_bad += 1
    print(rule, N, n_alpha, n_fields, 'mismatching repeats:', n_bad, flush=True)
    bad += n_bad
print('TOTAL mismatches', bad)
sys.exit(1 if bad else 0)
