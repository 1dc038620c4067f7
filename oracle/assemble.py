"""TEST INFRASTRUCTURE -- CPU oracle for the assemble+solve row.

PARITY UNPINNED: the reference contains no assembler and no solver (upstream that is the job of the
absent `fdm` package).  This oracle assembles the matrix the way `fdm`-style composition of the
reference's stencils would -- Operator(central(h), Operator(riesz_caputo_stencil, Operator(central(h))))
expanded at every interior node through oracle/fdm_shim.py, i.e. the operator d/dx o D^alpha_RC o d/dx of
SURVEY.md section 8(f)(1) -- and solves it with numpy.linalg.solve (LAPACK dgesv, partial pivoting).

Boundary rule (defined by this repo, the reference has none): n interior unknowns at x_k = k*h,
k = 1..n; homogeneous Dirichlet data, u = 0 at and beyond both ends (stencil nodes that fall outside
the interior are dropped).

Only tests/ and bench.py's CPU legs may import this."""
import numpy as np

from . import clib, fdm_shim


def rc_stencil(rule, alpha, h, res):
    first, w = clib.riesz_caputo_arrays(rule, (alpha, res * h, res))
    hh = (res * h) / float(res)
    return fdm_shim.Stencil({fdm_shim.Point((first + k) * hh): float(v) for k, v in enumerate(w)})


def assemble(rule, alpha, h, res, n):
    """Dense n x n matrix of the composed operator on the interior nodes x_k = k*h, k = 1..n."""
    op = fdm_shim.Operator(fdm_shim.Stencil.central(h),
                           fdm_shim.Operator(rc_stencil(rule, alpha, h, res),
                                             fdm_shim.Operator(fdm_shim.Stencil.central(h))))
    A = np.zeros((n, n))
    for i in range(1, n + 1):
        for node, w in op.expand(fdm_shim.Point(i * h)).items():
            j = int(round(node.x / h))
            assert abs(node.x / h - j) < 1e-6
            if 1 <= j <= n:
                A[i - 1, j - 1] += w
    return A


def solve(rule, alpha, h, res, n, rhs):
    A = assemble(rule, alpha, h, res, n)
    return np.linalg.solve(A, np.asarray(rhs, dtype=np.float64)), A
