"""EXTENSION: the time-stepping fractional-ODE driver around frx_history_solve.  No reference oracle exists
(the reference has no solver): manufactured solutions.  u = 1 + t^2 has a piecewise-linear derivative v = 2t, for which the
product-integration rule is exact, so the discrete solution must reproduce u to rounding; u = 1 + t^3 checks the O(h^2)
convergence order."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from fractulus_b200 import ode


@pytest.mark.parametrize('rule', ['caputo', 'trapezoidal'])
@pytest.mark.parametrize('N', [2000, 20000])
def test_linear_and_nonlinear_fode_manufactured_exact(rule, N):
    alpha, T = 0.6, 1.0
    h = T / N
    prob = ode.FractionalODE(rule, alpha, h, N)
    t = prob.t
    u_true = 1. + t * t
    du = ode.caputo_power(alpha, 2., t)                        # D^alpha (1 + t^2)
    for name, F in (('linear', lambda tt, u: -3. * (u - (1. + tt * tt)) + ode.caputo_power(alpha, 2., tt)),
                    ('nonlinear', lambda tt, u: -(u * u - (1. + tt * tt) ** 2) + ode.caputo_power(alpha, 2., tt))):
        u, info = prob.solve(F, 1.0, v0=0.0, windows=4)
        err = (u - u_true).abs().max().item()
        assert err <= 2e-10, (rule, N, name, err, info['sweeps'])
        assert info['residual'] <= 1e-8 * du.abs().max().item(), (rule, N, name, info['residual'])
        assert info['sweeps'] < 400
    # the two kernel-level operations are inverse to each other
    v = 2. * t
    assert (prob.derivative(v)[1:] - du[1:]).abs().max().item() <= 1e-11 * du.abs().max().item()
    back = prob.invert(prob.derivative(v), 0.0)
    assert (back - v).abs().max().item() <= 1e-9


def test_fode_second_order_convergence():
    alpha = 0.4
    errs = []
    for N in (250, 500, 1000):
        h = 1. / N
        prob = ode.FractionalODE('caputo', alpha, h, N)
        F = lambda tt, u: -(u - (1. + tt ** 3)) + ode.caputo_power(alpha, 3., tt)
        u, info = prob.solve(F, 1.0, v0=0.0, windows=2)
        errs.append((u - (1. + prob.t ** 3)).abs().max().item())
    assert errs[0] / errs[1] > 3.3 and errs[1] / errs[2] > 3.3, errs      # ~4 per halving of h
    with pytest.raises(ValueError):
        ode.FractionalODE('simpson', 0.5, 0.1, 10)
    with pytest.raises(ValueError):
        ode.FractionalODE('caputo', 1.5, 0.1, 10)
