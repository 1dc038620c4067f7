"""Host-side cost of one batched.assemble_solve call (no synchronisation inside the loop) against its GPU time."""
import os, sys, time, math
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fractulus_b200 import batched, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_alpha, n_ls = 400, 250
n_sys = n_alpha * n_ls
h = 1.0 / (n + 1)
alpha = np.repeat(np.linspace(0.05, 0.95, n_alpha), n_ls)
res = np.tile(1.0 + (np.arange(n_ls) % 30), n_alpha)
xg = torch.arange(1, n + 1, dtype=torch.float64, device='cuda') * h
rhs = torch.sin(math.pi * xg).repeat(n_sys, 1).contiguous()
x = torch.empty_like(rhs)
for _ in range(3):
    batched.assemble_solve('caputo', alpha, h, res, rhs, out=x)
torch.cuda.synchronize()
for steps in (5, 20):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    ev0.record()
    for _ in range(steps):
        t0 = time.perf_counter()
        batched.assemble_solve('caputo', alpha, h, res, rhs, out=x)
        host.append((time.perf_counter() - t0) * 1e3)
    ev1.record()
    torch.cuda.synchronize()
    print('n=%d steps=%d: device-timed ms/step %.3f; host ms per call: %s' % (n, steps, ev0.elapsed_time(ev1) / steps, ' '.join('%.2f' % t for t in host[:8])))
t0 = time.perf_counter()
for _ in range(20):
    a, hh, rr = batched._solve_params(alpha, h, res, n_sys)
print('_solve_params %.3f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
t0 = time.perf_counter()
for _ in range(20):
    info = torch.empty((n_sys,), dtype=torch.int32, device='cuda')
print('torch.empty info %.3f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
