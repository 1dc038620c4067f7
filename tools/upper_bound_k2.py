"""Tuning aid (trace build only): how much would the history kernel gain if staging instructions / the second barrier
cost nothing?  Results are WRONG in these modes; only the kernel time is read."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched, _lib
from fractulus_b200.batched import _ctx_for
N = 100000; h = 1.0 / N
x = np.arange(N + 2) * h
g = torch.as_tensor(3 * np.cos(3 * x) + 2 * x, device='cuda')
alphas = torch.as_tensor(np.array([0.5]), device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ctx, _ = _ctx_for(g.device)
L = _lib.lib()
for mode in (0, 1, 2, 3, 0):
    batched.history('caputo', alphas, h, g, N=N, validate=False)    # module loaded before the symbol is written
    assert L.frx_debug_k2_mode(mode) == 0
    for _ in range(5):
        batched.history('caputo', alphas, h, g, N=N, validate=False)
    ctx.set_timing(True)
    for _ in range(50):
        flush.zero_()
        batched.history('caputo', alphas, h, g, N=N, validate=False)
    t = ctx.kernel_times()
    ctx.set_timing(False)
    print('mode', mode, 'history_main ms', t['history_main'][0] / t['history_main'][1])
L.frx_debug_k2_mode(0)
