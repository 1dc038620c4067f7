"""Diagnostic: per fixture case, how far the GPU history is from the unmodified reference's rows."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched
from oracle.gen_golden import smooth_field
cases = json.load(open('tests/golden/history.json'))['cases']
for c in cases:
    alpha, h, N = float.fromhex(c['alpha']), float.fromhex(c['h']), c['N']
    g = torch.as_tensor(np.array(smooth_field(N + 2, h)), device='cuda')
    y = batched.history(c['rule'], alpha, h, g, N=N).cpu().numpy()
    rows = np.array(c['rows']); want = np.array([float.fromhex(v) for v in c['y']])
    rel = np.abs(y[rows] - want) / np.abs(want)
    print(c['rule'], alpha, N, 'max rel %.2e' % rel.max(), 'rows>1e-12: %d of %d' % ((rel > 1e-12).sum(), len(rows)),
          'worst row', rows[rel.argmax()], 'y', want[rel.argmax()])
