"""Profiling driver for the fused compute + gather kernel: ONE process, two GPUs.  GPU 0 runs frx_history_gather with two
targets -- its own result array and one in GPU 1's memory (peer access) -- so that `ncu` (single process) can count the
NVLink bytes the history kernel's epilogue sends.  Usage: python tools/prof_gather_p2p.py [n_alpha]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched  # noqa: E402

n_alpha = int(sys.argv[1]) if len(sys.argv) > 1 else 8
assert torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1)
N = 100000
h = 1.0 / N
x = np.arange(N + 2) * h
torch.cuda.set_device(0)
g = torch.as_tensor(3 * np.cos(3 * x) + 2 * x, device='cuda:0')
alphas = np.linspace(0.1, 0.9, n_alpha)
mine = torch.zeros((n_alpha, N + 1), dtype=torch.float64, device='cuda:0')
peer = torch.zeros((n_alpha, N + 1), dtype=torch.float64, device='cuda:1')
# kernels on GPU 0 dereference GPU 1's memory: peer access has to be enabled explicitly (the multi-process path gets its
# peer mappings from torch's symmetric memory instead)
import ctypes
rt = None
for name in ('libcudart.so.12', 'libcudart.so'):
    try:
        rt = ctypes.CDLL(name)
        break
    except OSError:
        pass
assert rt is not None, 'libcudart not found'
rc = rt.cudaDeviceEnablePeerAccess(ctypes.c_int(1), ctypes.c_uint(0))
assert rc in (0, 704), 'cudaDeviceEnablePeerAccess -> %d' % rc      # 704 = already enabled
torch.cuda.synchronize()
for _ in range(2):
    batched.history_gather('caputo', alphas, h, g, [mine.data_ptr(), peer.data_ptr()], 0, N, validate=True)
torch.cuda.synchronize()
same = torch.equal(mine.cpu(), peer.cpu())
want = batched.history('caputo', alphas, h, g, N=N)
print('ok', same, float((want - mine).abs().max()), 'bytes per launch to the peer: %d' % (n_alpha * (N + 1) * 8))
assert same
