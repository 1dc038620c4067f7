"""End-to-end time of cfg 4 through sweep.assemble_solve_sweep (page-locked host buffers used in place).  python tools/ab/e2e_k4.py n"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fractulus_b200 import batched, sweep
n = int(sys.argv[1])
n_alpha, n_ls = 400, 250
h = 1.0 / (n + 1)
alpha = np.repeat(np.linspace(0.05, 0.95, n_alpha), n_ls)
res = np.tile(1.0 + (np.arange(n_ls) % 30), n_alpha)
rhs_h = batched.pinned_empty((n_alpha * n_ls, n))
rhs_h[:] = np.sin(np.pi * np.arange(1, n + 1) * h)[None, :]
f = lambda: sweep.assemble_solve_sweep('caputo', alpha, h, res, rhs_h)
f(); f(); torch.cuda.synchronize()
ts = []
for _ in range(7):
    t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print('n=%d SCALAR=%s WARPS=%s e2e ms: min %.3f median %.3f' % (n, os.environ.get('FRX_K4_SCALAR'), os.environ.get('FRX_K4_SCALAR_WARPS'), min(ts), sorted(ts)[3]))
