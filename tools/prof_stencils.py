"""Profiling driver for K5: 1e5 Riesz-Caputo stencils (the cfg-4 sweep's settings) in one batched call, plus one single-setting
drop-in call.  Usage: python tools/prof_stencils.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fractulus_b200 as fr  # noqa: E402
from fractulus_b200 import batched  # noqa: E402

n = 100000
alpha = np.repeat(np.linspace(0.05, 0.95, 400), 250)
res = np.tile(1.0 + (np.arange(250) % 30), 400)
lf = res / 129.0
for _ in range(2):
    row_ptr, off, w = batched.stencil_weights('caputo', 'riesz_caputo', alpha, lf, res)
st = fr.create_left_stencil('caputo', fr.Settings(0.5, 10.0, 1000))
torch.cuda.synchronize()
print('ok', int(row_ptr[-1]), float(w[0]), len(st.weights_array))
