"""Timing aid: frx_history_solve at the BASELINE size (N = 1e5) and its round-trip error."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched
N = 100000; h = 1.0 / N
x = np.arange(N + 1) * h
u_true = torch.as_tensor(np.sin(3 * x) + x * x + 0.5, device='cuda')
for alphas in ([0.5], [0.2, 0.4, 0.6, 0.8]):
    f = batched.history('caputo', alphas, h, u_true, N=N)
    for _ in range(2):
        u = batched.history_solve('caputo', alphas, h, f, float(u_true[0]))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        u = batched.history_solve('caputo', alphas, h, f, float(u_true[0]))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    err = (u - u_true[None, :]).abs().max().item()
    print('n_alpha', len(alphas), 'ms per solve call', dt * 1e3, 'round-trip max err', err)

# attribution (event timing serialises the launches: shares, not absolutes)
from fractulus_b200.batched import _ctx_for
ctx, _ = _ctx_for(u_true.device)
f = batched.history('caputo', [0.5], h, u_true, N=N)
ctx.set_timing(True)
u = batched.history_solve('caputo', [0.5], h, f, float(u_true[0]))
print({k: (round(v[0], 4), v[1]) for k, v in ctx.kernel_times().items() if v[1]})
ctx.set_timing(False)
# host-side issue time of one call (no synchronisation inside)
torch.cuda.synchronize(); t0 = time.perf_counter()
u = batched.history_solve('caputo', [0.5], h, f, float(u_true[0]))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host issue ms', (t1 - t0) * 1e3, 'total ms', (t2 - t0) * 1e3)
