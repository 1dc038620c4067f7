"""Index-level prototype (NumPy, CPU) of the blocked band LU that kernel K4 `frx_k4_band_dmma` runs, used to pin
down the window / ring / slot bookkeeping before it is written as CUDA (development aid, not product, not oracle).

One system per warp.  Row slots never move (implicit pivoting, slots of retired rows are refilled with the rows
that enter the band window), column tiles of 8 form a ring, panels of 8 columns are factored with partial
pivoting, the trailing update of the whole window is one rank-8 product per 8x8 tile (DMMA m8n8k4 x 2 on the GPU).

    python tools/proto_band_lu.py        # random banded systems against numpy.linalg.solve
"""
import numpy as np


def geometry(nbt):
    rt = (nbt + 1 + 7) // 8             # row tiles: nb + 1 rows are alive at any time (the slot of a retired row is refilled at once)
    ct = (2 * nbt + 8 + 7) // 8         # column tiles: panel (8) + fill-in reach (2 nb)
    ct += 1 - ct % 2                    # odd: conflict-free C-fragment accesses
    return rt, ct


def solve_blocked(a, nb, n, rhs, nbt):
    """a[k + nb], k = -nb..nb: band coefficients of the Toeplitz matrix A[i][j] = a[j - i + nb]."""
    RT, CT = geometry(nbt)
    S, WC = RT * 8, CT * 8
    assert nb <= nbt and (nb <= n - 1 or n == 1)
    ustride = 2 * nb + 1
    A = lambda i, j: a[j - i + nb] if (abs(j - i) <= nb and 0 <= j < n) else 0.0
    W = np.zeros((S, WC))
    PB = np.zeros((S, 8))
    label = -np.ones(S, dtype=int)
    b = np.zeros(S)
    xs = np.array(rhs, dtype=float)
    U = np.zeros((n, ustride))
    info = 0
    for s in range(min(n, nb + 1)):
        label[s] = s
        b[s] = xs[s]
        for j in range(WC):
            W[s, j] = A(s, j)
        PB[s, :] = [A(s, j) for j in range(8)]
    n_panels = (n + 7) // 8
    TW = (CT - 1) * 8
    for p in range(n_panels):
        k0 = 8 * p
        pw = min(8, n - k0)
        col = lambda t: ((p + 1 + t // 8) % CT) * 8 + t % 8
        r = np.concatenate([PB.copy(), b[:, None]], axis=1)   # registers: the 8 panel entries + rhs of the lane's row
        L11 = np.zeros((8, 8))
        raw = np.zeros((8, TW))                               # registers u[jj][c] of lane t
        for c in range(pw):
            k = k0 + c
            act = [s for s in range(S) if label[s] >= 0]
            best = max(abs(r[s, c]) for s in act)
            P = min((s for s in act if abs(r[s, c]) == best), key=lambda s: label[s])
            blog = label[P]
            prow = r[P].copy()
            pval = prow[c]
            if pval == 0.0 and info == 0:
                info = k + 1
            rinv = 1.0 / pval if pval != 0.0 else 0.0
            for s in act:
                if s == P:
                    continue
                l = r[s, c] * rinv
                r[s, c] = l
                r[s, c + 1:] -= l * prow[c + 1:]
                if label[s] == k:
                    label[s] = blog
            raw[c, :] = [W[P, col(t)] for t in range(TW)]      # the pivot row leaves the window ...
            L11[c, :c] = r[P, :c]
            for cc in range(c, pw):
                if cc - c < ustride:
                    U[k, cc - c] = rinv if cc == c else r[P, cc]
            xs[k] = prow[8]
            inew = k + nb + 1                                  # ... and the row that enters the band takes its slot
            if inew < n:
                label[P] = inew
                for jj in range(8, WC):
                    W[P, ((p + jj // 8) % CT) * 8 + jj % 8] = A(inew, k0 + jj)
                r[P, :c + 1] = 0.0
                for cc in range(c + 1, 8):
                    r[P, cc] = A(inew, k0 + cc)
                r[P, 8] = xs[inew]
            else:
                label[P] = -1
                r[P, :] = 0.0
        M = np.where((label >= 0)[:, None], r[:, :8], 0.0)
        M[:, pw:] = 0.0
        b = r[:, 8].copy()
        Ub = np.zeros((8, TW))
        for q in range(pw):
            for t in range(TW):
                u = raw[q, t]
                for c in range(q):
                    u -= L11[q, c] * Ub[c, t]
                Ub[q, t] = u
                rel = 8 - q + t
                if k0 + 8 + t < n and rel < ustride:
                    U[k0 + q, rel] = u
                else:
                    assert u == 0.0, (p, q, t, u)
        for s in range(S):
            for t in range(TW):
                v = W[s, col(t)] - float(np.dot(M[s], Ub[:, t]))
                if t < 8:
                    PB[s, t] = v
                else:
                    W[s, col(t)] = v
        W[:, (p % CT) * 8:(p % CT) * 8 + 8] = 0.0
    x = np.zeros(n)
    for k in range(n - 1, -1, -1):
        ln = min(n - 1, k + 2 * nb) - k
        sacc = 0.0
        for c in range(1, ln + 1):
            sacc += U[k, c] * x[k + c]
        x[k] = (xs[k] - sacc) * U[k, 0]
    return x, info


def dense(a, nb, n):
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - nb), min(n, i + nb + 1)):
            A[i, j] = a[j - i + nb]
    return A


def main():
    rng = np.random.default_rng(0)
    worst = 0.0
    cases = 0
    for nbt in (7, 15, 23, 31):
        for n in (1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33, 40, 64, 71, 100, 128):
            for trial in range(3):
                nb_full = int(rng.integers(max(1, nbt - 7), nbt + 1))
                nb = min(nb_full, max(n - 1, 0))
                half = rng.standard_normal(nb + 1)
                if trial == 0:
                    half[0] = 0.1             # weak diagonal: forces pivoting
                a = np.concatenate([half[:0:-1], half])   # symmetric band
                if n == 1:
                    a = np.array([half[0] if half[0] != 0 else 1.0])
                    nb = 0
                rhs = rng.standard_normal(n)
                A = dense(a, nb, n)
                want = np.linalg.solve(A, rhs)
                got, info = solve_blocked(a, nb, n, rhs, nbt)
                err = np.abs(got - want).max() / np.abs(want).max()
                cond = np.linalg.cond(A)
                worst = max(worst, err / cond)
                assert info == 0
                assert err <= 1e-13 * cond + 1e-12, (nbt, n, nb, err, cond)
                cases += 1
    print('ok: %d cases, worst err/cond = %.2e' % (cases, worst))


if __name__ == '__main__':
    main()
