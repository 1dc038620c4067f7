"""Drop-in for reference ``fractulus/equation.py``: the same four public names with the same
signatures, keys and exceptions, the weights computed by hand-written sm_100a CUDA kernels
(fractulus_b200/csrc/frx_weights.cu, kernel K5) behind the C ABI of include/frx.h.

    Settings(alpha, lf, resolution)                      reference equation.py:10
    create_left_stencil(approximation, settings)         reference equation.py:28-29
    create_right_stencil(approximation, settings)        reference equation.py:32-35
    create_riesz_caputo_stencil(approximation, settings) reference equation.py:13-17
    approximation in 'caputo' | 'rectangle' | 'trapezoidal' | 'simpson'   reference equation.py:179-184

The returned object is an ``fdm.equation.Stencil`` when the reference's ``fdm`` dependency is
importable, otherwise the equivalent carrier of fractulus_b200/fdm_carrier.py.  There is no CPU path:
without the CUDA library or without a GPU these functions raise.
"""
import collections
import ctypes
import os
import sys

import numpy as np

from . import _lib

__all__ = 'Settings', 'create_left_stencil', 'create_right_stencil', 'create_riesz_caputo_stencil'

# namedtuple type name 'CaputoSettings' as in the reference (equation.py:10)
Settings = collections.namedtuple('CaputoSettings', ('alpha', 'lf', 'resolution'))


def _carrier():
    """(Stencil, Point) classes: real fdm unless absent / disabled / a test shim."""
    mode = os.environ.get('FRACTULUS_B200_FDM', 'auto')
    if mode != 'builtin':
        try:
            import fdm
            if not getattr(fdm, '__fdm_shim__', False):
                from fdm.equation import Stencil
                from fdm.geometry import Point
                return Stencil, Point, False
        except ImportError:
            if mode == 'external':
                raise
    from . import fdm_carrier
    return fdm_carrier.Stencil, fdm_carrier.Point, True


def _node_coordinates(approximation, side, lf, resolution, first, count):
    """Node coordinates the way the reference places them: start + j * (end - start) / n_intervals
    (equation.py:59,70-71,98,174-176 through Stencil.uniform, equation.py:81), mirrored by
    symmetry(Point(0.)) for the right stencil (equation.py:38-39); Riesz-Caputo nodes are the union."""
    p = int(resolution)
    if approximation == 'rectangle':
        start = -lf + lf / float(resolution)
        intervals, end = p - 1, 0.
    elif approximation == 'simpson' and p % 2 == 1:
        start, end, intervals = -lf, lf / float(resolution), p + 1
    else:
        start, end, intervals = -lf, 0., p
    delta = (end - start) / intervals
    left_first = -(p - 1) if approximation == 'rectangle' else -p
    left = {left_first + j: start + delta * j for j in range(intervals + 1)}
    xs = []
    for o in range(first, first + count):
        if side == 'left':
            xs.append(left[o])
        elif side == 'right':
            xs.append(0. * 2. - left[-o])
        else:
            xs.append(left[o] - 0. if o in left else (0. * 2. - left[-o]) - 0.)
    return xs


def _lazy_nodes(approximation, side, lf, resolution, first, count):
    return lambda: _node_coordinates(approximation, side, lf, resolution, first, count)


def _stencil(approximation, side, settings):
    rule = _lib.RULES[approximation]          # KeyError for an unknown key, like _factory[approximation]
    alpha, lf, resolution = settings          # same unpacking (and exceptions) as the reference
    L = _lib.lib()
    alpha_f, lf_f, res_f = float(alpha), float(lf), float(resolution)
    first_c, count_c = ctypes.c_int32(), ctypes.c_int32()
    # the domain checks (KeyError / ZeroDivisionError / ValueError like the reference) need no GPU
    _lib.check(L.frx_stencil_layout(rule, _lib.SIDES[side], alpha_f, res_f, ctypes.byref(first_c), ctypes.byref(count_c)))
    cap = count_c.value
    weights = np.empty(cap, dtype=np.float64)
    offsets = np.empty(cap, dtype=np.int32)
    got = ctypes.c_int32()
    ctx = _lib.context()
    _lib.check(L.frx_stencil_weights_host(ctx.handle, rule, _lib.SIDES[side], alpha_f, lf_f, res_f, cap,
                                          offsets.ctypes.data_as(_lib.c_int32_p), weights.ctypes.data_as(_lib.c_double_p),
                                          ctypes.byref(got)))
    n = got.value
    first = int(offsets[0])
    weights = weights[:n]
    Stencil, Point, builtin = _carrier()
    if builtin:       # node coordinates and the {Point: weight} dictionary are produced when first asked for
        return Stencil.from_layout(lambda: _node_coordinates(approximation, side, lf, resolution, first, n), weights, first,
                                   lf / float(resolution))
    xs = _node_coordinates(approximation, side, lf, resolution, first, n)
    return Stencil({Point(x): float(w) for x, w in zip(xs, weights)})


def create_left_stencil(approximation, settings):
    return _stencil(approximation, 'left', settings)


def create_right_stencil(approximation, settings):
    return _stencil(approximation, 'right', settings)


def create_riesz_caputo_stencil(approximation, settings):
    return _stencil(approximation, 'riesz_caputo', settings)


# ---- batched conveniences (no counterpart in the reference): a list of Settings -> a list of stencils ------------------
# One vectorised layout call, ONE K5 launch and one device->host copy for the whole list instead of two launches and two
# copies per stencil; the stencils are the same objects the single-setting functions return.

def _stencils(approximation, side, settings_list):
    _lib.RULES[approximation]                 # KeyError for an unknown key
    triples = [tuple(s) for s in settings_list]
    if not triples:
        return []
    for t in triples:
        alpha, lf, resolution = t             # same unpacking (and exceptions) as the reference
    from . import batched
    alpha = np.array([float(t[0]) for t in triples])
    lf = np.array([float(t[1]) for t in triples])
    res = np.array([float(t[2]) for t in triples])
    row_ptr, first = batched.stencil_layouts(approximation, side, alpha, res)
    _, _, weights = batched.stencil_weights(approximation, side, alpha, lf, res)
    w = weights.cpu().numpy()
    Stencil, Point, builtin = _carrier()
    out = []
    for k, (a, l, r) in enumerate(triples):
        lo, hi = int(row_ptr[k]), int(row_ptr[k + 1])
        xs = None if builtin else _node_coordinates(approximation, side, l, r, int(first[k]), hi - lo)
        if builtin:
            out.append(Stencil.from_layout(_lazy_nodes(approximation, side, l, r, int(first[k]), hi - lo), w[lo:hi].copy(),
                                           int(first[k]), l / float(r)))
        else:
            out.append(Stencil({Point(x): float(v) for x, v in zip(xs, w[lo:hi])}))
    return out


def create_left_stencils(approximation, settings_list):
    return _stencils(approximation, 'left', settings_list)


def create_right_stencils(approximation, settings_list):
    return _stencils(approximation, 'right', settings_list)


def create_riesz_caputo_stencils(approximation, settings_list):
    return _stencils(approximation, 'riesz_caputo', settings_list)
