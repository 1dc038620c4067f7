"""CPU suite: the oracle (oracle/ref_port.py, oracle/frx_oracle.c) against the fixtures generated from
the UNMODIFIED reference (tests/golden/, made by oracle/gen_golden.py) and against the reference's own
hard-coded golden vectors (reference fractulus/test/unit/test_equation.py)."""
import math

import numpy as np
import pytest

from conftest import unhex
from oracle import clib, ref_port

S = ref_port.Settings

# the reference's five 17-digit vectors, keyed by integer offset
# (reference fractulus/test/unit/test_equation.py:20-26, 47-52, 61-67, 76-82, 89-97)
REFERENCE_UNIT_VECTORS = [
    ('caputo', S(0.5, 0.6, 4), -4, [0.057148272422657305, 0.12706258982171437, 0.1571224994043748,
                                     0.24135913466702896, 0.29134624815788773]),
    ('rectangle', S(0.5, 0.8, 4), -3, [0.1352142643344008, 0.16038909801255471, 0.20902314205707648,
                                        0.504626504404032]),
    ('trapezoidal', S(0.5, 0.8, 4), -4, [0.06598914093388651, 0.14671924087499555, 0.1814294346537252,
                                          0.2786975227427686, 0.33641766960268793]),
    ('simpson', S(0.5, 0.8, 4), -4, [0.04139442395762801, 0.19590867482751353, 0.10587690665922159,
                                      0.38061314477925734, 0.28545985858444356]),
    ('simpson', S(0.5, 1.0, 5), -5, [0.03727938104424598, 0.1690131707314828, 0.09488746599316666,
                                      0.24273847930859488, 0.214401233455065, 0.40370120352322564,
                                      -0.033641766960268805]),
]


@pytest.mark.parametrize('impl', [ref_port, clib], ids=['python_port', 'c_oracle'])
def test_reference_unit_vectors_bit_exact(impl):
    for rule, settings, first, weights in REFERENCE_UNIT_VECTORS:
        f, w = impl.left_arrays(rule, settings)
        assert f == first
        assert [float(x) for x in w] == weights, (rule, settings)


@pytest.mark.parametrize('impl', [ref_port, clib], ids=['python_port', 'c_oracle'])
def test_stencils_vs_unmodified_reference_bit_exact(impl, golden_stencils):
    checked = 0
    for c in golden_stencils:
        settings = S(float.fromhex(c['alpha']), float.fromhex(c['lf']), c['resolution'])
        for side, fn in (('left', impl.left_arrays), ('right', impl.right_arrays),
                         ('riesz_caputo', impl.riesz_caputo_arrays)):
            rec = c[side]
            if 'raises' in rec:              # alpha = 2: the Riesz-Caputo constant gamma(2. - alpha) is at a pole
                with pytest.raises(ValueError):
                    fn(c['rule'], settings)
                continue
            first, w = fn(c['rule'], settings)
            assert first == rec['first_offset'] and len(w) == rec['count'], (c['rule'], settings, side)
            assert [float(x).hex() for x in w] == rec['weights'], (c['rule'], settings, side)
            checked += 1
    assert checked >= 400


def test_c_gamma_is_cpython_gamma():
    rng = np.random.default_rng(7)
    for x in np.concatenate([rng.uniform(0.5, 6.0, 20000), [1., 2., 3., 4., 5., 2.5, 1.5]]):
        assert clib.tgamma(x) == math.gamma(x)


def test_errors_match_reference_behaviour():
    for impl in (ref_port, clib):
        with pytest.raises(KeyError):
            impl.left_arrays('grunwald', S(0.5, 1., 4))          # equation.py:29 dict lookup
        with pytest.raises(ZeroDivisionError):
            impl.left_arrays('caputo', S(0.5, 1., 0))            # equation.py:56
        with pytest.raises(ZeroDivisionError):
            impl.left_arrays('rectangle', S(1.5, 1., 4))         # equation.py:68, 0.0 ** negative


def test_c_history_naive_vs_reference_rows_bit_exact(golden_history):
    from oracle.gen_golden import smooth_field
    for c in golden_history:
        alpha, h, N = float.fromhex(c['alpha']), float.fromhex(c['h']), c['N']
        g = np.array(smooth_field(N + 2, h))
        y = clib.history_rows_naive(c['rule'], alpha, h, g, c['rows'])
        assert [float(v).hex() for v in y] == c['y'], (c['rule'], alpha, N)


def test_c_history_full_equals_naive():
    from oracle.gen_golden import smooth_field
    N, h = 3000, 1. / 3000
    g = np.array(smooth_field(N + 2, h))
    for rule in ref_port.RULES:
        lo = {'caputo': 1, 'trapezoidal': 1, 'rectangle': 2, 'simpson': 2}[rule]
        rows = np.arange(lo, N + 1)
        full = clib.history_full(rule, 0.37, h, N, g)
        naive = clib.history_rows_naive(rule, 0.37, h, g, rows)
        assert np.array_equal(full[lo:], naive), rule
        assert full[0] == 0.


def test_python_port_history_equals_c(golden_history):
    from oracle.gen_golden import smooth_field
    N, h = 200, 1. / 200
    g = smooth_field(N + 2, h)
    for rule in ref_port.RULES:
        rows = list(range(3, N + 1, 7))
        yp = ref_port.history(rule, 0.61, h, g, rows)
        yc = clib.history_rows_naive(rule, 0.61, h, np.array(g), rows)
        assert yp == [float(v) for v in yc], rule


def test_tables_reassemble_stencils():
    """The Toeplitz decomposition (cA/cB/w0/mult/cm1) that the GPU kernels use reproduces every left
    stencil of the growing-history pattern bit-exactly."""
    N, h, alpha = 67, 0.013, 0.42
    for rule in ref_port.RULES:
        cA, cB, w0, mult, cm1 = clib.tables(rule, alpha, h, N)
        lo = {'caputo': 1, 'trapezoidal': 1, 'rectangle': 2, 'simpson': 2}[rule]
        for i in range(lo, N + 1):
            first, w = clib.left_arrays(rule, S(alpha, i * h, i))
            for k, wk in enumerate(w):
                d = -(first + k)
                if d == -1:
                    raw = cm1
                elif d == i:
                    raw = w0[i]
                else:
                    raw = cB[d] if (rule == 'simpson' and (i - d) % 2 == 1) else cA[d]
                assert mult[i] * raw == wk, (rule, i, d)
