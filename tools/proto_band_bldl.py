"""NumPy prototype of frx_k4_band_bldl<T> (csrc/frx_solve_ldl.cuh): BLOCK LDL^T of a symmetric definite banded Toeplitz
system with 8 x 8 block pivots -- the diagonal tile is inverted in place by Gauss-Jordan steps without pivoting (the
pivots are those of the scalar LDL^T: definiteness check), L21 = A21 Dinv and the trailing update are tile products, the
trailing block is T(T+1)/2 lower tiles that move one tile up and left per panel, entering tiles come from the Toeplitz
coefficient vector, rows beyond n are identity padding; forward substitution rides along, back substitution is
x_p = Dinv z_p - sum_t X_t^T x_{p+t}.  Checks index bookkeeping AND the residuals the explicit tile inverse leaves on
the real cfg-4 matrices (oracle coefficients).

    python tools/proto_band_bldl.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def gj_inverse(M, sgn, n_valid):
    """in-place Gauss-Jordan inverse without pivoting; returns (inverse, ok)"""
    M = M.copy()
    for c in range(8):
        p = M[c, c]
        if c < n_valid and not (p * sgn > 0):
            return None, False
        r = 1.0 / p
        f = M[:, c].copy()
        prow = M[c, :] * r
        new = M - np.outer(f, prow)
        new[c, :] = prow
        new[:, c] = -f * r
        new[c, c] = r
        M = new
    return M, True


def solve_bldl(a_low, n, nb, T, b):
    assert nb <= 8 * T
    OFF = 8 * T + 8
    apad = np.zeros(2 * OFF + 1)
    for k in range(nb + 1):
        apad[OFF + k] = apad[OFF - k] = a_low[k]

    def value(R, C):
        if R < n and C < n:
            return apad[OFF + R - C] if abs(R - C) <= OFF else 0.0
        return 1.0 if R == C else 0.0

    def fresh(bi, bj):
        return np.array([[value(8 * bi + g, 8 * bj + c) for c in range(8)] for g in range(8)])

    P = (n + 7) // 8
    bpad = np.zeros(8 * (P + T + 1))
    bpad[:n] = b
    S = {(i, j): fresh(i, j) for i in range(T) for j in range(i + 1)}
    bv = [bpad[8 * t:8 * t + 8].copy() for t in range(T)]
    sgn = np.sign(a_low[0])
    store = []
    for p in range(P):
        Dinv, ok = gj_inverse(S[0, 0], sgn, n - 8 * p)
        if not ok:
            return None, False
        A21 = [None] + [S[t, 0] if t <= T - 1 else fresh(p + T, p) for t in range(1, T + 1)]
        Xn = [None] + [A21[t] @ (-Dinv) for t in range(1, T + 1)]            # X' = -X
        z = bv[0]
        store.append((-Dinv, Xn, z))
        newS = {}
        for i in range(T):
            for j in range(i + 1):
                src = S[i + 1, j + 1] if i + 1 <= T - 1 else fresh(p + 1 + i, p + 1 + j)
                newS[i, j] = src + Xn[i + 1] @ A21[j + 1].T
        S = newS
        nbv = []
        for t in range(1, T + 1):
            cur = bv[t] if t <= T - 1 else bpad[8 * (p + T):8 * (p + T) + 8].copy()
            nbv.append(cur + Xn[t] @ z)
        bv = nbv
    x = np.zeros(8 * (P + T + 1))
    for p in range(P - 1, -1, -1):
        nDinv, Xn, z = store[p]
        acc = -(nDinv.T @ z)
        for t in range(1, T + 1):
            acc += Xn[t].T @ x[8 * (p + t):8 * (p + t) + 8]
        x[8 * p:8 * p + 8] = acc
    return x[:n], True


def toeplitz(a_low, n):
    A = np.zeros((n, n))
    for k, v in enumerate(a_low):
        if k < n:
            A += np.diag(np.full(n - k, v), -k)
            if k:
                A += np.diag(np.full(n - k, v), k)
    return A


def main():
    rng = np.random.default_rng(3)
    n_ok = n_flag = 0
    for trial in range(300):
        T = int(rng.integers(1, 5))
        n = int(rng.integers(1, 70))
        nb = int(rng.integers(0, min(8 * T, max(n - 1, 0)) + 1))
        a = rng.normal(size=nb + 1)
        if trial % 3:
            a[0] = (np.abs(a[1:]).sum() * 2 + 1.0) * (1 if trial % 2 else -1)
        A = toeplitz(a, n)
        b = rng.normal(size=n)
        x, ok = solve_bldl(a, n, nb, T, b)
        ev = np.linalg.eigvalsh(A)
        is_def = ev.min() > 0 or ev.max() < 0
        assert ok == is_def or (not ok and abs(ev).min() < 1e-9), (trial, ok, is_def)
        if ok:
            want = np.linalg.solve(A, b)
            assert np.abs(x - want).max() / np.abs(want).max() < 1e-10 * np.linalg.cond(A), trial
            n_ok += 1
        else:
            n_flag += 1
    print('block LDL prototype: %d definite systems solved, %d indefinite flagged' % (n_ok, n_flag))
    # residual quality on the real operator (oracle coefficients)
    from oracle import clib
    worst = (0, None)
    worst_err = (0, None)
    cnt = 0
    for n in (64, 128, 256):
        h = 1.0 / (n + 1)
        for alpha in np.linspace(0.05, 0.95, 10):
            for res in (1, 2, 3, 5, 8, 13, 21, 30):
                first, w = clib.riesz_caputo_arrays('caputo', (alpha, res * h, res))
                rcp = np.zeros(2 * res + 5)
                rcp[2:-2] = w
                k = np.arange(0, res + 2)
                a = (rcp[k + res + 2 - 1] - 2 * rcp[k + res + 2] + rcp[k + res + 2 + 1]) / h ** 2
                nb = min(res + 1, n - 1)
                a = a[:nb + 1]
                T = (nb + 7) // 8 if nb > 7 else 1
                T = 1 if nb <= 7 else 2 if nb <= 15 else 3 if nb <= 23 else 4
                A = toeplitz(a, n)
                b = np.sin(np.pi * np.arange(1, n + 1) * h) * (1 + rng.normal(size=n) * 0.1)
                x, ok = solve_bldl(a, n, nb, T, b)
                if not ok:
                    continue
                cnt += 1
                r = np.abs(A @ x - b).max() / (np.abs(A).max() * np.abs(x).max() + np.abs(b).max())
                want = np.linalg.solve(A, b)
                e = np.abs(x - want).max() / np.abs(want).max() / np.linalg.cond(A)
                if r > worst[0]:
                    worst = (r, (n, alpha, res))
                if e > worst_err[0]:
                    worst_err = (e, (n, alpha, res))
    print('real operator: %d definite systems, worst residual %.2e at %s, worst err/cond %.2e at %s' % (cnt, worst[0], worst[1], worst_err[0], worst_err[1]))


def refinement_statistics():
    """How many refinement steps the INDEFINITE cfg-4 matrices need when they are factorised without pivoting (no sign test)
    and accepted on  ||b - A x|| <= 16 eps (max|a| ||x|| + ||b||):  the numbers behind kBldlRefineSteps in
    csrc/frx_solve_ldl.cuh."""
    import collections
    from oracle import clib
    global gj_inverse
    checked = gj_inverse

    def gj_no_sign_test(M, sgn, n_valid):
        M = M.copy()
        for c in range(8):
            p = M[c, c]
            if p == 0:
                return None, False
            r = 1.0 / p
            f = M[:, c].copy()
            prow = M[c, :] * r
            new = M - np.outer(f, prow)
            new[c, :] = prow
            new[:, c] = -f * r
            new[c, c] = r
            M = new
        return M, True

    gj_inverse = gj_no_sign_test
    eps = np.finfo(float).eps
    try:
        for n in (128, 256):
            stats, tot = collections.Counter(), 0
            h = 1.0 / (n + 1)
            for alpha in np.linspace(0.05, 0.95, 37):
                for res in range(1, 31):
                    first, w = clib.riesz_caputo_arrays('caputo', (alpha, res * h, res))
                    rcp = np.zeros(2 * res + 5)
                    rcp[2:-2] = w
                    k = np.arange(0, res + 2)
                    a = (rcp[k + res + 2 - 1] - 2 * rcp[k + res + 2] + rcp[k + res + 2 + 1]) / h ** 2
                    nb = min(res + 1, n - 1)
                    a = a[:nb + 1]
                    T = 1 if nb <= 7 else 2 if nb <= 15 else 3 if nb <= 23 else 4
                    A = toeplitz(a, n)
                    ev = np.linalg.eigvalsh(A)
                    if ev.min() > 0 or ev.max() < 0:
                        continue
                    tot += 1
                    b = np.sin(np.pi * np.arange(1, n + 1) * h)
                    with np.errstate(all='ignore'):
                        x, ok = solve_bldl(a, n, nb, T, b)
                        if not ok or not np.isfinite(x).all():
                            stats['never'] += 1
                            continue
                        amax, done = np.abs(a).max(), False
                        for it in range(9):
                            r = b - A @ x
                            if np.abs(r).max() <= 16 * eps * (amax * np.abs(x).max() + np.abs(b).max()):
                                stats[it] += 1
                                done = True
                                break
                            d, ok = solve_bldl(a, n, nb, T, r)
                            if not ok or not np.isfinite(d).all():
                                break
                            x = x + d
                        if not done:
                            stats['never'] += 1
            print('n = %d: %d indefinite matrices, accepted after k refinement steps: %s'
                  % (n, tot, sorted(stats.items(), key=lambda kv: str(kv[0]))))
    finally:
        gj_inverse = checked


if __name__ == '__main__':
    if '--refine' in sys.argv:
        refinement_statistics()
    else:
        main()
