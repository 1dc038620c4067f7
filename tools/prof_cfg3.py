"""Profiling driver: K3-shaped problems (N = 16384, B fields, one alpha).  Usage: python tools/prof_cfg3.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = 16384
G = torch.rand((B, N + 2), dtype=torch.float64, device='cuda')
for _ in range(3):
    Y = batched.toeplitz_apply('caputo', 0.5, 1.0 / N, G, N=N)
torch.cuda.synchronize()
print('ok', float(Y[0, N]))
