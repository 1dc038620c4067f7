"""Diagnostic: frx_history_solve (FRX_ODE_VARIANT from the environment) against the oracle's forward substitution."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched
from oracle import clib
for N in (5000, 20000, 50000):
    h = 1. / N
    x = np.arange(N + 1) * h
    u_true = np.sin(3 * x) + x * x + 0.5
    for rule, alphas in (('caputo', [0.3, 0.7]), ('trapezoidal', [0.5])):
        f = batched.history(rule, alphas, h, torch.as_tensor(u_true, device='cuda'), N=N)
        u = batched.history_solve(rule, alphas, h, f, u_true[0]).cpu().numpy()
        for ai, a in enumerate(alphas):
            want = clib.history_solve(rule, a, h, f[ai].cpu().numpy(), u_true[0])
            d = np.abs(u[ai] - want)
            print(os.environ.get('FRX_ODE_VARIANT', '0'), rule, a, N, 'vs oracle', d.max(), 'at', int(d.argmax()),
                  'gpu vs true', np.abs(u[ai] - u_true).max(), 'oracle vs true', np.abs(want - u_true).max(), flush=True)
