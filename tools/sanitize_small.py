"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractulus_b200 as fr  # noqa: E402
from fractulus_b200 import batched  # noqa: E402

for rule in ('caputo', 'rectangle', 'trapezoidal', 'simpson'):
    fr.create_left_stencil(rule, fr.Settings(0.5, 0.8, 6))
    fr.create_riesz_caputo_stencil(rule, fr.Settings(0.5, 0.8, 6))
    N = 2600
    g = torch.linspace(0, 1, N + 2, dtype=torch.float64, device='cuda').cos()
    y = batched.history(rule, [0.3, 0.7], 1.0 / N, g, N=N)
    if rule != 'simpson':
        batched.history(rule, 0.4, 1.0 / N, g, N=N, side='riesz_caputo')
batched.toeplitz_apply('caputo', 0.5, 1e-3, torch.rand((9, 1500), dtype=torch.float64, device='cuda'))
rhs = torch.ones((40, 64), dtype=torch.float64, device='cuda')
batched.assemble_solve('caputo', np.linspace(0.1, 0.9, 40), 1 / 65, 1 + np.arange(40) % 30, rhs)        # streaming kernel
batched.assemble_solve('caputo', np.linspace(0.1, 0.9, 40), 1 / 65, 1 + np.arange(40) % 40, rhs)        # CTA kernel
batched.weights_table('trapezoidal', [0.2, 0.6], 1e-3, 900)
torch.cuda.synchronize()
print('ok')
