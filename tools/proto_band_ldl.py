"""NumPy prototype of the index bookkeeping of frx_k4_band_ldl<T> (csrc/frx_solve_ldl.cuh): blocked pivot-free elimination
of a symmetric banded Toeplitz system in panels of 8, the trailing block kept as T(T+1)/2 lower 8x8 tiles, rows and columns
that enter the band built from the coefficient vector, identity padding beyond n, U rows in the layout the back
substitution reads.  Checked against numpy.linalg.solve on random definite band matrices (and that indefinite ones are
flagged).

    python tools/proto_band_ldl.py
"""
import numpy as np


def solve_ldl(a_low, n, nb, T, b):
    """a_low[k] = a_k, k = 0..nb (row - col = k); returns (x, ok)."""
    assert nb <= 8 * T and nb <= max(n - 1, 0)
    OFF = 8 * T + 8
    apad = np.zeros(2 * OFF + 1)
    for k in range(0, nb + 1):
        apad[OFF + k] = a_low[k]
        apad[OFF - k] = a_low[k]

    def value(R, C):
        if R < n and C < n:
            return apad[OFF + R - C]
        return 1.0 if R == C else 0.0

    us = nb + 1
    P = (n + 7) // 8
    U = np.full((P * 8, us), np.nan)
    xs = np.zeros(P * 8 + 8 * T + 8)
    xs[:n] = b
    SL = {}
    for i in range(T):
        for j in range(i + 1):
            SL[i, j] = np.array([[value(8 * i + g, 8 * j + c) for c in range(8)] for g in range(8)])
    sgn = np.sign(a_low[0])
    for p in range(P):
        k0 = 8 * p
        d = np.zeros((8, 9))
        d[:, :8] = SL[0, 0]
        d[:, 8] = xs[k0:k0 + 8]
        r = np.zeros((8 * T, 9))
        for l in range(8 * T):
            t = l >> 3
            R = k0 + 8 + l
            if t + 1 <= T - 1:
                r[l, :8] = SL[t + 1, 0][l & 7]
            else:
                r[l, :8] = [value(R, k0 + c) for c in range(8)]
            r[l, 8] = xs[R] if R < n else 0.0
        x = np.zeros((8 * T, 8))
        y = np.zeros((8 * T, 8))
        rinvs = np.zeros(8)
        for c in range(8):
            pv = d[c].copy()
            pval = pv[c]
            if k0 + c < n and not (pval * sgn > 0):
                return None, False
            rinv = 1.0 / pval
            rinvs[c] = rinv
            for l in range(c + 1, 8):
                m = d[l, c] * rinv
                d[l, c] = m
                d[l, c + 1:] -= m * pv[c + 1:]
            y[:, c] = r[:, c]
            m = r[:, c] * rinv
            x[:, c] = m
            r[:, c + 1:] -= np.outer(m, pv[c + 1:])
        for q in range(8):
            if k0 + q < n:
                U[k0 + q, 0] = rinvs[q]
                for cc in range(q + 1, 8):
                    if cc - q < us and k0 + cc < n:
                        U[k0 + q, cc - q] = d[q, cc]
                xs[k0 + q] = d[q, 8]
        for l in range(8 * T):
            R = k0 + 8 + l
            if R < n:
                xs[R] = r[l, 8]
                for q in range(8):
                    rel = 8 + l - q
                    if rel < us and k0 + q < n:
                        U[k0 + q, rel] = y[l, q]
        new = {}
        for i in range(T):
            for j in range(i + 1):
                if i + 1 <= T - 1:
                    src = SL[i + 1, j + 1].copy()
                else:
                    src = np.array([[value(k0 + 8 + 8 * i + g, k0 + 8 + 8 * j + c) for c in range(8)] for g in range(8)])
                new[i, j] = src - x[8 * i:8 * i + 8] @ y[8 * j:8 * j + 8].T
        SL = new
    # back substitution on the stored rows
    xsol = np.zeros(n)
    for k in range(n - 1, -1, -1):
        acc = xs[k]
        for rel in range(1, us):
            if k + rel < n:
                assert not np.isnan(U[k, rel]), (k, rel)
                acc -= U[k, rel] * xsol[k + rel]
        xsol[k] = acc * U[k, 0]
    return xsol, True


def main():
    rng = np.random.default_rng(3)
    n_ok = n_flag = 0
    for trial in range(300):
        T = int(rng.integers(1, 5))
        n = int(rng.integers(1, 70))
        nb = int(rng.integers(0, min(8 * T, max(n - 1, 0)) + 1))
        a = rng.normal(size=nb + 1)
        definite = trial % 3 != 0
        if definite:
            a[0] = (np.abs(a[1:]).sum() * 2 + 1.0) * (1 if trial % 2 else -1)
        A = np.zeros((n, n))
        for k in range(nb + 1):
            if k < n:
                A += np.diag(np.full(n - k, a[k]), -k)
                if k:
                    A += np.diag(np.full(n - k, a[k]), k)
        b = rng.normal(size=n)
        x, ok = solve_ldl(a, n, nb, T, b)
        ev = np.linalg.eigvalsh(A)
        is_def = ev.min() > 0 or ev.max() < 0
        assert ok == is_def or (not ok and abs(ev).min() < 1e-9), (trial, ok, is_def)
        if ok:
            want = np.linalg.solve(A, b)
            err = np.abs(x - want).max() / np.abs(want).max()
            assert err < 1e-10 * np.linalg.cond(A), (trial, err)
            n_ok += 1
        else:
            n_flag += 1
    print('band LDL prototype: %d definite systems solved, %d indefinite flagged' % (n_ok, n_flag))


if __name__ == '__main__':
    main()
