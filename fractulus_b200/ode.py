"""Time-stepping fractional ODEs on top of the growing-history operator (EXTENSION: the reference
has neither a solver nor an ODE driver -- upstream that is `fdm`'s job).

The rows of the history operator are the reference's product-integration stencils (reference fractulus/equation.py:42-59 for
'caputo'): applied to samples of v = u' they give the Caputo derivative

        (D^alpha u)(t_i)  =  history(rule, alpha, h, v)[i],        v = u',   0 < alpha < 1.

A Caputo initial-value problem  D^alpha u = F(t, u),  u(0) = u0  is therefore, on the grid t_i = i h,

        history(v)[i] = F(t_i, u_i),    u_i = u0 + sum of trapezoids of v up to t_i        (i = 1..N)

-- the classical implicit product-integration scheme in derivative form.  Marching it node by node costs one O(i) history
sum per step (O(N^2) dependent work).  Here the whole time axis is advanced at once by waveform relaxation

        v^{k+1} = history^{-1}( F(t, u[v^k]) )

where history^{-1} is kernel-level work (frx_history_solve: Newton iteration for the inverse series on the tensor-core
Toeplitz kernel, ~1 ms at N = 1e5) and everything else is elementwise / a prefix sum on the device.  The Volterra structure
makes the iteration converge for every Lipschitz F on a bounded interval; it is carried out window by window in time
(`windows`) so that stiff problems converge in a few sweeps per window.

    FractionalODE(rule, alpha, h, N).solve(F, u0)   ->  u[N+1], info
"""
import math

import numpy as np
import torch

from . import batched


class FractionalODE(object):
    """Caputo initial-value problem D^alpha u = F(t, u), u(0) = u0 on t_i = i*h, i = 0..N, 0 < alpha < 1.

    rule: 'caputo' | 'trapezoidal' (the rules frx_history_solve inverts).  All state lives on `device`."""

    def __init__(self, rule, alpha, h, N, device=None):
        if rule not in ('caputo', 'trapezoidal'):
            raise ValueError("the ODE driver needs a rule with a non-zero diagonal weight: 'caputo' or 'trapezoidal'")
        if not 0. < float(alpha) < 1.:
            raise ValueError('alpha must lie in (0, 1)')
        self.rule, self.alpha, self.h, self.N = rule, float(alpha), float(h), int(N)
        self.device = torch.device(device) if device is not None else batched.default_device()
        self.t = torch.arange(self.N + 1, dtype=torch.float64, device=self.device) * self.h

    # ---- the two kernel-level operations ---------------------------------------------------------------------------
    def derivative(self, v):
        """(D^alpha u)(t_i), i = 0..N, from samples v_i = u'(t_i) (row i = reference stencil of Settings(alpha, i*h, i))."""
        return batched.history(self.rule, self.alpha, self.h, v, N=self.N)

    def invert(self, f, v0):
        """v with derivative(v)[i] = f[i], i = 1..N, and v[0] = v0."""
        return batched.history_solve(self.rule, [self.alpha], self.h, f.reshape(1, -1), v0)[0]

    def integrate(self, v, u0):
        """u_i = u0 + trapezoids of v (exact for the piecewise-linear v the quadrature rule assumes)."""
        inc = 0.5 * self.h * (v[1:] + v[:-1])
        u = torch.empty_like(v)
        u[0] = u0
        u[1:] = u0 + torch.cumsum(inc, 0)
        return u

    # ---- the driver ------------------------------------------------------------------------------------------------
    def solve(self, F, u0, v0=None, tol=1e-12, max_sweeps=200, windows=1):
        """F(t, u) -> tensor (elementwise, torch ops on the device).  v0 = u'(0) if known (default: F-consistent guess 0).
        windows: the time axis is converged in this many consecutive windows (each sweep still runs the kernels over
        [0, end of the window]; earlier windows are frozen).  Returns (u, info)."""
        N = self.N
        u0 = float(u0)
        v = torch.zeros(N + 1, dtype=torch.float64, device=self.device)
        if v0 is not None:
            v[0] = float(v0)
        u = self.integrate(v, u0)
        sweeps, history = 0, []
        edges = [int(round(N * (k + 1) / windows)) for k in range(windows)]
        begin = 0
        for end in edges:
            if end <= begin:
                continue
            n_w = end
            sub = FractionalODE(self.rule, self.alpha, self.h, n_w, self.device) if n_w != N else self
            for _ in range(max_sweeps):
                f = F(self.t[:n_w + 1], u[:n_w + 1])
                v_new = sub.invert(f, v[0].item())
                v_new[:begin + 1] = v[:begin + 1]                      # converged windows stay frozen
                u_new = sub.integrate(v_new, u0)
                delta = (u_new[begin:] - u[begin:n_w + 1]).abs().max().item()
                scale = max(u_new.abs().max().item(), 1e-300)
                v[:n_w + 1], u[:n_w + 1] = v_new, u_new
                sweeps += 1
                history.append(delta / scale)
                if delta <= tol * scale:
                    break
            begin = end
        residual = (self.derivative(v)[1:] - F(self.t, u)[1:]).abs().max().item()
        return u, {'sweeps': sweeps, 'last_update': history[-1] if history else 0., 'residual': residual, 'v': v}


def caputo_power(alpha, m, t):
    """D^alpha t^m = Gamma(m+1)/Gamma(m+1-alpha) t^(m-alpha) (m > 0): manufactured solutions for tests."""
    return math.gamma(m + 1.) / math.gamma(m + 1. - alpha) * t ** (m - alpha)
