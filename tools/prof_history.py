"""Profiling driver: a few cfg-2 steps (N = 1e5, 'caputo', alpha = 0.5, one smooth field) and nothing else,
so that `ncu` sees only the hot path.  Usage: python tools/prof_history.py [n_steps] [n_alpha] [rule]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched  # noqa: E402

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_alpha = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rule = sys.argv[3] if len(sys.argv) > 3 else 'caputo'
N = 100000
h = 1.0 / N
x = np.arange(N + 2) * h
g = torch.as_tensor(3 * np.cos(3 * x) + 2 * x, device='cuda')
alphas = torch.as_tensor(np.linspace(0.1, 0.9, n_alpha) if n_alpha > 1 else np.array([0.5]), device='cuda')
y = None
for _ in range(n_steps):
    y = batched.history(rule, alphas, h, g, N=N, validate=False)
torch.cuda.synchronize()
print('ok', float(y.reshape(-1, N + 1)[0, N]))
