"""Tuning driver for K4 (assemble + batch-solve): times one batch shape per process (the kernel choice is read from
the environment once: FRX_K4_VARIANT / FRX_K4_DMMA_BINS) and checks residuals through the kernel's own dense assembly.

    python tools/time_k4.py --n 64 --res-lo 1 --res-hi 30 [--n-sys 100000] [--steps 5]
prints one JSON line: solves/s, ms per batch, the per-step solve-kernel time, max relative residual."""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import _lib, batched  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=64)
ap.add_argument('--n-sys', type=int, default=100000)
ap.add_argument('--res-lo', type=int, default=1)
ap.add_argument('--res-hi', type=int, default=30)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--rule', default='caputo')
a = ap.parse_args()
n, n_sys = a.n, a.n_sys
h = 1.0 / (n + 1)
alpha = np.linspace(0.05, 0.95, n_sys)
res = float(a.res_lo) + (np.arange(n_sys) % (a.res_hi - a.res_lo + 1))
xg = torch.arange(1, n + 1, dtype=torch.float64, device='cuda') * h
rhs = (torch.sin(math.pi * xg)[None, :] * (1.0 + 0.001 * torch.arange(n_sys, dtype=torch.float64, device='cuda')[:, None])).contiguous()
ctx = _lib.context(0)
for _ in range(2):
    x, info = batched.assemble_solve(a.rule, alpha, h, res, rhs, return_info=True)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(a.steps):
    x = batched.assemble_solve(a.rule, alpha, h, res, rhs)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / a.steps
ctx.set_timing(True)
for _ in range(a.steps):
    x = batched.assemble_solve(a.rule, alpha, h, res, rhs)
kt = ctx.kernel_times()
ctx.set_timing(False)
idx = np.unique(np.concatenate([np.arange(0, n_sys, max(1, n_sys // 300)), np.arange(min(n_sys, 64))]))
A = batched.assemble_dense(a.rule, alpha[idx], h, res[idx], n)
r = torch.bmm(A, x[idx].unsqueeze(-1)).squeeze(-1) - rhs[idx]
scale = (A.abs().amax(dim=(1, 2)) * x[idx].abs().amax(dim=1)).unsqueeze(-1)
resid = (r.abs() / scale).max().item()
nb = np.minimum(res + 1, n - 1)
flop = float(np.sum(n * (2.0 * nb * (2 * nb + 1) + 3 * nb) + n * (4 * nb + 1)))
solve_ms = kt['solve'][0] / a.steps
print(json.dumps({'n': n, 'n_sys': n_sys, 'res': [a.res_lo, a.res_hi], 'variant': os.environ.get('FRX_K4_VARIANT', '0'),
                  'dmma_bins': os.environ.get('FRX_K4_DMMA_BINS', 'default'), 'two_warp': os.environ.get('FRX_K4_TWO_WARP', 'default'), 'ms_per_batch': ms,
                  'solves_per_s': n_sys / (ms * 1e-3), 'solve_kernels_ms': solve_ms, 'stencil_kernels_ms': kt['stencil'][0] / a.steps,
                  'band_tflops': flop / (solve_ms * 1e-3) / 1e12, 'max_rel_residual': resid, 'info_nonzero': int((info != 0).sum()),
                  'finite': bool(torch.isfinite(x).all())}))
