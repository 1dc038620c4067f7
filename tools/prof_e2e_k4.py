"""Where the end-to-end time of cfg 4 goes: raw pinned copies, device solve, and the chunked host-buffer path for several
chunk sizes.   python tools/prof_e2e_k4.py [n]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched, sweep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_alpha, n_ls = 400, 250
n_sys = n_alpha * n_ls
h = 1.0 / (n + 1)
alpha = np.repeat(np.linspace(0.05, 0.95, n_alpha), n_ls)
res = np.tile(1.0 + (np.arange(n_ls) % 30), n_alpha)
rhs_h = batched.pinned_empty((n_sys, n))
rhs_h[:] = np.sin(np.pi * np.arange(1, n + 1) * h)[None, :]
out_h = batched.pinned_empty((n_sys, n))
d = torch.empty((n_sys, n), dtype=torch.float64, device='cuda')


def timed(f, reps=5):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


t_in, t_out = torch.as_tensor(rhs_h), torch.as_tensor(out_h)
print('pinned flags', t_in.is_pinned(), t_out.is_pinned())
print('H2D %.1f MB: %.3f ms' % (rhs_h.nbytes / 1e6, timed(lambda: d.copy_(t_in, non_blocking=True))))
print('D2H: %.3f ms' % timed(lambda: t_out.copy_(d, non_blocking=True)))
print('device solve: %.3f ms' % timed(lambda: batched.assemble_solve('caputo', alpha, h, res, d)))
for chunk in (4096, 8192, 16384, 32768, 65536, 100000):
    print('assemble_solve_host chunk %6d: %.3f ms' % (chunk, timed(lambda: batched.assemble_solve_host('caputo', alpha, h, res, rhs_h, out=out_h, chunk=chunk), 3)))
print('sweep.assemble_solve_sweep: %.3f ms' % timed(lambda: sweep.assemble_solve_sweep('caputo', alpha, h, res, rhs_h), 3))
# host-side cost of one call's parameter pass (no GPU work: tiny n_sys slices do not change the per-system host cost)
t0 = time.perf_counter()
for _ in range(5):
    a, hh, rr = batched._solve_params(alpha, h, res, n_sys)
print('_solve_params: %.3f ms' % ((time.perf_counter() - t0) / 5 * 1e3))
for m in (4096, 8192, 16384, 32768, 50000):
    dd = d[:m]
    t = timed(lambda: batched.assemble_solve('caputo', alpha[:m], h, res[:m], dd))
    print('device solve of the first %6d systems: %.3f ms  (%.1f ns per system)' % (m, t, t * 1e6 / m))
