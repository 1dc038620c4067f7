"""TEST INFRASTRUCTURE -- ctypes loader for oracle/_build/libfrx_oracle.so (the plain-C restatement
in oracle/frx_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this."""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libfrx_oracle.so')
# the same C file built with the product's correctly rounded pow / exp (see the header of frx_oracle.c)
_SO_CR = os.path.join(_HERE, '_build', 'libfrx_oracle_cr.so')
RULE_CODE = {'caputo': 0, 'rectangle': 1, 'trapezoidal': 2, 'simpson': 3}
_libs = {}


def build(force=False):
    srcs = [os.path.join(_HERE, 'frx_oracle.c'), os.path.join(_HERE, 'cr_math.cpp'), os.path.join(_HERE, 'Makefile'),
            os.path.join(os.path.dirname(_HERE), 'fractulus_b200', 'csrc', 'frx_math.cuh')]
    newest = max(os.path.getmtime(p) for p in srcs)
    if force or any(not os.path.exists(so) or os.path.getmtime(so) < newest for so in (_SO, _SO_CR)):
        subprocess.check_call(['make', '-C', _HERE, '-B', '-s'])
    return _SO


def lib(cr=False):
    """cr=False: glibc pow (the reference's arithmetic); cr=True: correctly rounded pow (summation-order checks)."""
    if cr not in _libs:
        build()
        L = ctypes.CDLL(_SO_CR if cr else _SO)
        dp, lp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_int)
        L.orc_tgamma.restype = ctypes.c_double
        L.orc_tgamma.argtypes = [ctypes.c_double]
        L.orc_left_layout.argtypes = [ctypes.c_int, ctypes.c_double, ip, ip]
        L.orc_left_weights.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ip, ip, dp]
        L.orc_tables.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_long, dp, dp, dp, dp, dp]
        L.orc_history_rows_naive.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, dp, ctypes.c_long,
                                             lp, ctypes.c_long, dp, ctypes.c_int]
        L.orc_history_full.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_long, dp,
                                       ctypes.c_long, ctypes.c_long, ctypes.c_long, dp, ctypes.c_long, ctypes.c_int]
        L.orc_history_full_abs.argtypes = L.orc_history_full.argtypes
        L.orc_history_rows_naive_abs.argtypes = L.orc_history_rows_naive.argtypes
        L.orc_history_solve.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_long, dp, ctypes.c_double, dp]
        L.orc_pow.restype = ctypes.c_double
        L.orc_pow.argtypes = [ctypes.c_double, ctypes.c_double]
        L.orc_pow_many.argtypes = [dp, dp, dp, ctypes.c_long]
        _libs[cr] = L
    return _libs[cr]


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _raise(rc):
    if rc == -3:
        raise ZeroDivisionError('float division by zero')
    if rc == -1:
        raise KeyError('unknown rule')
    if rc:
        raise ValueError('settings outside the domain of the rule (rc=%d)' % rc)


def num_threads():
    return lib().orc_num_threads()


def tgamma(x):
    return lib().orc_tgamma(float(x))


def left_arrays(rule, settings, cr=False):
    """(first_offset, np.float64[count]) of the reference's create_left_stencil(rule, settings)."""
    alpha, lf, res = settings
    code = RULE_CODE[rule]
    if res == 0:
        raise ZeroDivisionError('float division by zero')
    first, count = ctypes.c_int(), ctypes.c_int()
    w = np.empty(int(res) + 2, dtype=np.float64)
    _raise(lib(cr).orc_left_weights(code, float(alpha), float(lf), float(res), ctypes.byref(first),
                                  ctypes.byref(count), _dp(w)))
    return first.value, w[:count.value].copy()


def right_arrays(rule, settings):
    first, w = left_arrays(rule, settings)           # equation.py:32-39
    return -(first + len(w) - 1), (w * -1.)[::-1].copy()


def riesz_caputo_arrays(rule, settings):
    """equation.py:13-25 with Python-float scalar arithmetic (identical to oracle/ref_port.py)."""
    alpha = settings[0]
    first, w = left_arrays(rule, settings)
    n = math.floor(alpha) + 1.
    K = 1. / 2. * math.gamma(2. - alpha) / math.gamma(2.)
    sgn = (-1.) ** n
    left = {first + i: float(x) for i, x in enumerate(w)}
    lo = min(first, -(first + len(w) - 1))
    out = []
    for o in range(lo, -lo + 1):
        a = left.get(o)
        b = sgn * (left[-o] * -1.) if -o in left else None
        out.append(K * (a if b is None else (b if a is None else a + b)))
    return lo, np.array(out)


def tables(rule, alpha, h, N, cr=False):
    cA, cB = np.zeros(N), np.zeros(N)
    w0, mult = np.zeros(N + 1), np.zeros(N + 1)
    cm1 = ctypes.c_double()
    _raise(lib(cr).orc_tables(RULE_CODE[rule], float(alpha), float(h), N, _dp(cA), _dp(cB), _dp(w0), _dp(mult),
                            ctypes.byref(cm1)))
    return cA, cB, w0, mult, cm1.value


def history_rows_naive(rule, alpha, h, g, rows, n_threads=0, cr=False, abs_terms=False):
    """The reference's own algorithm: every row rebuilds its stencil (O(i) pow) and sums.
    abs_terms: sum of |g_j w_ij| instead (the rounding scale of an uncompensated sum of that row)."""
    g = np.ascontiguousarray(g, dtype=np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    y = np.empty(len(rows))
    fn = lib(cr).orc_history_rows_naive_abs if abs_terms else lib(cr).orc_history_rows_naive
    fn(RULE_CODE[rule], float(alpha), float(h), _dp(g), len(g),
                                 rows.ctypes.data_as(ctypes.POINTER(ctypes.c_long)), len(rows), _dp(y), n_threads)
    return y


def history_full(rule, alpha, h, N, g, n_threads=0, cr=False, abs_terms=False):
    """Same values for rows 0..N (row 0 = 0), table-accelerated; g is [n_g] or [n_fields, n_g]."""
    g = np.ascontiguousarray(g, dtype=np.float64)
    g2 = g.reshape(-1, g.shape[-1])
    y = np.empty((g2.shape[0], N + 1))
    fn = lib(cr).orc_history_full_abs if abs_terms else lib(cr).orc_history_full
    _raise(fn(RULE_CODE[rule], float(alpha), float(h), N, _dp(g2), g2.shape[1], g2.shape[0],
                                  g2.shape[1], _dp(y), N + 1, n_threads))
    return y.reshape(g.shape[:-1] + (N + 1,))


def history_solve(rule, alpha, h, f, u0, cr=False):
    """u with history(u)[i] = f[i], i = 1..N, u[0] = u0: forward substitution with the reference's weights."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    u = np.empty_like(f)
    _raise(lib(cr).orc_history_solve(RULE_CODE[rule], float(alpha), float(h), len(f) - 1, _dp(f), float(u0), _dp(u)))
    return u


def pow_many(x, y, cr=False):
    """Element-wise x ** y with this library's pow (glibc's, or the correctly rounded one for cr=True)."""
    x = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), np.broadcast(x, y).shape))
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(y, dtype=np.float64), x.shape))
    out = np.empty(x.shape)
    lib(cr).orc_pow_many(_dp(x.reshape(-1)), _dp(y.reshape(-1)), _dp(out.reshape(-1)), x.size)
    return out
