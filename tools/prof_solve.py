"""Profiling driver for K4: a cfg-4 style batch (alpha x resolution sweep) and nothing else.
Usage: python tools/prof_solve.py [n] [n_sys]"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_sys = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
h = 1.0 / (n + 1)
alpha = np.linspace(0.05, 0.95, n_sys)
res = 1.0 + (np.arange(n_sys) % 30)
xg = torch.arange(1, n + 1, dtype=torch.float64, device='cuda') * h
rhs = torch.sin(math.pi * xg).repeat(n_sys, 1).contiguous()
for _ in range(2):
    x = batched.assemble_solve('caputo', alpha, h, res, rhs)
torch.cuda.synchronize()
print('ok', float(x[0, n // 2]))
