"""Tuning aid: per-CTA phase timestamps of the history kernel (needs a library built with -DFRX_K2_TRACE, see
tools/build_trace_lib.sh).  Prints when CTAs start, get their first data, finish computing and exit."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractulus_b200 import batched, _lib
N = 100000; h = 1.0 / N
x = np.arange(N + 2) * h
g = torch.as_tensor(3 * np.cos(3 * x) + 2 * x, device='cuda')
alphas = torch.as_tensor(np.array([0.5]), device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(5):
    flush.zero_()
    y = batched.history('caputo', alphas, h, g, N=N, validate=False)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (1024 * 16))()
L = _lib.lib()
assert L.frx_debug_k2_trace(buf) == 0
n_cta = int(os.environ.get('FRX_TRACE_CTAS', '592'))
raw = np.array(buf, dtype=np.uint64).reshape(1024, 16)[:n_cta]
t = raw.astype(np.int64)
t0 = t[:, 0].min()
names = ['start', 'after dependency wait', 'first unit in smem', 'last unit in smem', 'last compute done', 'exit']
for k, nm in enumerate(names):
    v = (t[:, k] - t0) / 1e3
    print('%-24s min %8.2f  p10 %8.2f  median %8.2f  p90 %8.2f  max %8.2f us' % (nm, v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max()))
tail = (t[:, 5] - t[:, 4]) / 1e3
print('epilogue after last compute: median %.2f max %.2f us' % (np.median(tail), tail.max()))
print('kernel span (first start -> last exit): %.2f us' % ((t[:, 5].max() - t0) / 1e3))
sm = (raw[:, 6] & np.uint64(0xffffffff)).astype(np.int64)
groups = (raw[:, 6] >> np.uint64(32)).astype(np.int64)
comp = (raw[:, 7] & np.uint64(0xffffffff)).astype(np.int64)
total = (raw[:, 7] >> np.uint64(32)).astype(np.int64)
print('thread-0 cycles inside compute_unit per group (4 column blocks = 32 DMMA per warp, 512 pipe cycles): median %.0f; share of the CTA lifetime spent inside compute_unit: median %.3f' % (np.median(comp / groups), np.median(comp / total)))
per_sm_exit = np.array([t[sm == s, 5].max() for s in np.unique(sm)])
print('SMs used', len(np.unique(sm)), 'CTAs per SM min/max', np.bincount(sm).min(), np.bincount(sm).max(),
      'per-SM last exit spread: min %.2f max %.2f us' % ((per_sm_exit.min() - t0) / 1e3, (per_sm_exit.max() - t0) / 1e3))
order = np.argsort(-t[:, 5])[:12]
print('latest CTAs (blockIdx, smid, first-unit, last-compute, exit us, groups, cycles/group):')
for k in order:
    print('  %4d %4d %8.2f %8.2f %8.2f %6d %6.0f' % (k, sm[k], (t[k, 2] - t0) / 1e3, (t[k, 4] - t0) / 1e3, (t[k, 5] - t0) / 1e3, groups[k], comp[k] / max(groups[k], 1)))
cpg = comp / np.maximum(groups, 1)
print('cycles/group by blockIdx decile:', [round(float(np.median(cpg[i::10])), 0) for i in range(1)], 'min %.0f max %.0f' % (cpg.min(), cpg.max()))
late_sm = np.unique(sm[order])
print('SMs of the latest CTAs:', late_sm.tolist())

fin = np.nonzero(raw[:, 12] == 1)[0]
if len(fin):
    d = lambda a, b: (t[fin, a] - t[fin, b]) / 1e3
    print('finishers of a split tile at their span end: %d CTAs; medians (us): last compute -> slot published %.2f, -> ticket %.2f, '
          '-> slots summed %.2f, -> rows stored %.2f, -> exit %.2f' % (len(fin), np.median(d(8, 4)), np.median(d(9, 8)),
          np.median(d(10, 9)), np.median(d(11, 10)), np.median(d(5, 11))))
