"""GPU check of the pivot-free first attempt of K4 (csrc/frx_solve_ldl.cuh): residuals through the kernel's own dense
assembly, the number of systems that fell through to the pivoted LU against the number of indefinite matrices (eigenvalues
of the dense assembly), for several n including ones that are not multiples of 8.

    python tools/check_k4_ldl.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import _lib, batched  # noqa: E402

ctx = _lib.context(0)
out = []
for rule in ('caputo', 'trapezoidal'):
    for n in (3, 5, 9, 30, 50, 64, 100, 128, 201, 256):
        alphas = np.linspace(0.05, 0.95, 37)
        ress = np.arange(1, 31)
        alpha = np.repeat(alphas, len(ress))
        res = np.tile(ress, len(alphas)).astype(np.float64)
        n_sys = len(alpha)
        h = 1.0 / (n + 1)
        g = torch.Generator(device='cuda').manual_seed(n)
        rhs = torch.randn(n_sys, n, dtype=torch.float64, device='cuda', generator=g)
        x, info = batched.assemble_solve(rule, alpha, h, res, rhs, return_info=True)
        total, pivoted, piv_bins, refined_bins = ctx.solve_stats(per_bin=True)
        A = batched.assemble_dense(rule, alpha, h, res, n)
        r = torch.bmm(A, x.unsqueeze(-1)).squeeze(-1) - rhs
        scale = (A.abs().amax(dim=(1, 2)) * x.abs().amax(dim=1)).unsqueeze(-1) + rhs.abs().amax(dim=1, keepdim=True)
        resid = (r.abs() / scale).amax(dim=1)
        ev = torch.linalg.eigvalsh(A)
        definite = (ev[:, 0] > 0) | (ev[:, -1] < 0)
        want = torch.linalg.solve(A, rhs)
        cond = (ev.abs().amax(dim=1) / ev.abs().amin(dim=1))
        err = ((x - want).abs().amax(dim=1) / want.abs().amax(dim=1)) / cond
        rec = {'rule': rule, 'n': n, 'n_sys': n_sys, 'stats_total': total, 'pivoted': pivoted, 'refined': int(sum(refined_bins)),
               'indefinite_by_eig': int((~definite).sum()), 'max_resid': float(resid.max()), 'max_err_over_cond': float(err.max()),
               'info_nonzero': int((info != 0).sum()), 'finite': bool(torch.isfinite(x).all())}
        print(json.dumps(rec), flush=True)
        out.append(rec)
ok = all(r['finite'] and r['max_resid'] < 1e-14 and r['stats_total'] == r['n_sys'] and r['max_err_over_cond'] < 1e-14
         and r['pivoted'] + r['refined'] == r['indefinite_by_eig'] for r in out)
print('OK' if ok else 'FAILED')
sys.exit(0 if ok else 1)
