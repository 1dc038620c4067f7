"""Randomised differential check of the history paths against the C oracle (tests' own tolerance helpers): random rules,
sizes around the tile boundaries, batch shapes, both entry points, pinned and pageable host buffers, the solve round
trip.  Usage: python tools/fuzz_parity.py [n_cases] [seed] [largest N]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from fractulus_b200 import batched
from oracle import clib
from test_gpu_history import assert_history_close
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_max = int(sys.argv[3]) if len(sys.argv) > 3 else 6000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
special = [1, 2, 3, 7, 8, 9, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 3071, 3072, 3073, 4097]
fails = 0
for case in range(n_cases):
    rule = ['caputo', 'rectangle', 'trapezoidal', 'simpson'][rng.integers(4)]
    N = int(special[rng.integers(len(special))] if rng.random() < 0.6 else rng.integers(1, n_max))
    if rule == 'simpson':
        N = max(N, 2)
    n_alpha, n_fields = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    alphas = np.sort(rng.uniform(0.02, 0.98, n_alpha))
    h = float(rng.uniform(0.3, 3.0) / N)
    x = np.arange(N + 2) * h
    fields = np.stack([np.cos((1 + f) * x) * (1 + 0.3 * f) + 0.5 * x + 0.1 * f for f in range(n_fields)])
    try:
        got = batched.history(rule, alphas, h, torch.as_tensor(fields, device='cuda'), N=N).cpu().numpy()
        host = batched.history_host(rule, alphas, h, fields, N=N)
        assert np.array_equal(got, host, equal_nan=True), 'host (pageable) != device'
        gp = batched.pinned_empty(fields.shape); gp[:] = fields
        out = batched.pinned_empty(got.shape)
        assert np.array_equal(got, batched.history_host(rule, alphas, h, gp, N=N, out=out), equal_nan=True), 'host (pinned) != device'
        lo = 2 if rule in ('rectangle', 'simpson') else 1
        if N >= lo:
            for ai, a in enumerate(alphas):
                want = clib.history_full(rule, float(a), h, N, fields)
                for f in range(n_fields):
                    try:
                        assert_history_close(rule, float(a), h, np.arange(lo, N + 1), got[ai][f, lo:], want[f, lo:], fields[f],
                                             0.8 if rule == 'simpson' else 0.97)   # (share of rows within the plain 1e-12)
                    except AssertionError:
                        # sign-changing fields at large N: the rounding of the N-term sums (tile order here, CPython's
                        # compensated sum in the oracle) reaches the zero-crossing floor 1e-15 * max|y| of the contract
                        err = np.abs(got[ai][f, lo:] - want[f, lo:]).max() / np.abs(want[f, lo:]).max()
                        if err > 64 * np.finfo(float).eps:
                            raise
                        print('     (case %d: summation noise %.1e of max|y| at a zero crossing)' % (case, err), flush=True)
        if rule in ('caputo', 'trapezoidal') and N >= 2:
            u_true = fields[0, :N + 1]
            fgpu = batched.history(rule, alphas, h, torch.as_tensor(u_true, device='cuda'), N=N)
            u = batched.history_solve(rule, alphas, h, fgpu, float(u_true[0])).cpu().numpy()
            err = np.abs(u - u_true[None, :]).max() / np.abs(u_true).max()
            assert err <= 1e-9, ('solve round trip', err)
        print('ok  ', case, rule, N, n_alpha, n_fields, flush=True)
    except AssertionError as ex:
        fails += 1
        print('FAIL', case, rule, N, n_alpha, n_fields, repr(ex)[:300], flush=True)
print('cases', n_cases, 'failures', fails)
sys.exit(1 if fails else 0)
