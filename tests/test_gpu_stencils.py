"""GPU parity, stencil level (kernel K5 through the drop-in API and the C ABI), against
 (a) the reference's own golden vectors and tests, restated (reference fractulus/test/unit/test_equation.py,
     fractulus/test/integration/test_equation.py),
 (b) tests/golden/stencils.json, produced by the UNMODIFIED reference (oracle/gen_golden.py),
 (c) the C oracle (oracle/frx_oracle.c) on seeded random settings.
Bar: integer layouts bit-exact; weights bit-exact for small stencils and within a few ulp of the
largest power in the cancelling sum otherwise (glibc pow is not correctly rounded, ours is)."""
import math

import numpy as np
import pytest

from conftest import unhex
from test_cpu_host import weight_tolerance

pytestmark = pytest.mark.gpu

import fractulus_b200 as fr
from fractulus_b200 import batched, fdm_carrier
from fractulus_b200.fdm_carrier import Point, Stencil
from oracle import clib

Settings = fr.Settings


# ---- (a) the reference's unit tests, restated against this package -------------------------------

def test_LeftSide_Always_ReturnStencilWithCorrectWeights():
    # reference fractulus/test/unit/test_equation.py:9-28
    settings = Settings(alpha=0.5, lf=0.6, resolution=4)
    results = fr.create_left_stencil('caputo', settings)._weights
    expected = {
        Point(-0.44999999999999996, 0.0, 0.0): 0.12706258982171437,
        Point(-0.6, 0.0, 0.0): 0.057148272422657305,
        Point(0.0, 0.0, 0.0): 0.29134624815788773,
        Point(-0.15000000000000002, 0.0, 0.0): 0.24135913466702896,
        Point(-0.3, 0.0, 0.0): 0.1571224994043748,
    }
    assert expected == results


def test_Expand_AlphaAlmostOne_ReturnOnlyOneWeightForGivenNodeAddress():
    # reference fractulus/test/unit/test_equation.py:30-40
    stencil = fr.create_riesz_caputo_stencil('caputo', Settings(0.999999, 1., 1))
    result = stencil.expand(Point(0.))
    assert abs(1. - result[Point(0)].real) < 1e-5
    assert abs(1. - sum([v.real for v in result.values()])) < 1e-5


@pytest.mark.parametrize('rule,settings,expected', [
    ('rectangle', Settings(0.5, 0.8, 4), {-0.6: 0.1352142643344008, -0.4: 0.16038909801255471,
                                          -0.2: 0.20902314205707648, -0.0: 0.504626504404032}),          # :43-54
    ('trapezoidal', Settings(0.5, 0.8, 4), {-0.8: 0.06598914093388651, -0.6: 0.14671924087499555,
                                            -0.4: 0.1814294346537252, -0.2: 0.2786975227427686,
                                            -0.0: 0.33641766960268793}),                               # :57-69
    ('simpson', Settings(0.5, 0.8, 4), {-0.8: 0.04139442395762801, -0.6: 0.19590867482751353,
                                        -0.4: 0.10587690665922159, -0.2: 0.38061314477925734,
                                        -0.0: 0.28545985858444356}),                                   # :72-84
    ('simpson', Settings(0.5, 1.0, 5), {-1.0: 0.03727938104424598, -0.8: 0.1690131707314828,
                                        -0.6: 0.09488746599316666, -0.4: 0.24273847930859488,
                                        -0.2: 0.214401233455065, -0.0: 0.40370120352322564,
                                        0.2: -0.033641766960268805}),                                  # :86-99
])
def test_Create_Left_ReturnCorrectStencil(rule, settings, expected):
    result = fr.create_left_stencil(rule, settings)
    assert Stencil({Point(x): w for x, w in expected.items()}) == result


# ---- the reference's integration studies (fractulus/test/integration/test_equation.py) -----------

def test_Expand_NodeZeroAndAlphaAlmostOne_ReturnNodeZeroWeightAlmostOne():            # :10-19
    stencil = fr.create_left_stencil('caputo', Settings(0.9999, 10., 10))
    assert abs(1. - stencil.expand(Point(0))[Point(0)]) < 1e-3


def _left_caputo_linear(alpha, test_range):                                              # :48-66
    out = []
    for i in test_range:
        op = fdm_carrier.Operator(fr.equation.create_left_stencil('caputo', fr.equation.Settings(alpha, i, i)),
                                  fdm_carrier.Operator(Stencil.central(.1)))
        out.append(sum([node.x * weight for node, weight in op.expand(Point(i)).items()]))
    return out


def test_LeftCaputoForLinearFunctionStudies():
    np.testing.assert_almost_equal([1.] * 14, _left_caputo_linear(0.9999, range(1, 15)), decimal=3)   # :26-35
    np.testing.assert_almost_equal([1], _left_caputo_linear(0.00001, range(1, 2)), decimal=3)          # :37-46


def _riesz_caputo(function, alpha, lf, test_range):                                      # :74-94
    out = []
    for i in test_range:
        op = fdm_carrier.Operator(fr.create_riesz_caputo_stencil('caputo', fr.Settings(alpha, lf, lf)),
                                  fdm_carrier.Number(function))
        out.append(sum(op.expand(Point(i)).values()))
    return out


def test_RieszCaputoStudies():
    lin, quad = (lambda p: p.x), (lambda p: p.x ** 2)
    r = range(0, 15)
    np.testing.assert_almost_equal([i for i in r], _riesz_caputo(lin, 0.99999, 4, r), decimal=3)       # :101-111
    np.testing.assert_almost_equal([i * 4 for i in r], _riesz_caputo(lin, 0.000000000001, 4, r), decimal=2)  # :113-123
    np.testing.assert_almost_equal([i * 4 ** (1. - 0.8) for i in r], _riesz_caputo(lin, 0.8, 4, r), decimal=3)  # :125-136
    np.testing.assert_almost_equal([i ** 2 for i in r], _riesz_caputo(quad, 0.999999, 4, r), decimal=3)  # :143-153


def test_errors_like_the_reference():
    with pytest.raises(KeyError):
        fr.create_left_stencil('nope', Settings(0.5, 1., 4))
    with pytest.raises(ZeroDivisionError):
        fr.create_right_stencil('trapezoidal', Settings(0.5, 1., 0))
    with pytest.raises(ZeroDivisionError):
        fr.create_riesz_caputo_stencil('rectangle', Settings(1.2, 1., 3))


# ---- (b) fixtures from the unmodified reference ---------------------------------------------------

def test_all_golden_stencils_layout_exact_weights_within_ulps(golden_stencils):
    total = differing = 0
    for c in golden_stencils:
        a, lf, res = float.fromhex(c['alpha']), float.fromhex(c['lf']), c['resolution']
        s = Settings(a, lf, res)
        tol = weight_tolerance(c['rule'], a, lf, res)
        for side, fn in (('left', fr.create_left_stencil), ('right', fr.create_right_stencil),
                         ('riesz_caputo', fr.create_riesz_caputo_stencil)):
            rec = c[side]
            if 'raises' in rec:              # alpha = 2: ValueError like the reference (gamma(2. - alpha), equation.py:23)
                with pytest.raises(ValueError):
                    fn(c['rule'], s)
                continue
            st = fn(c['rule'], s)
            assert st.first_offset == rec['first_offset'] and len(st.weights_array) == rec['count']
            ref = np.array(unhex(rec['weights']))
            ne = st.weights_array != ref
            total += len(ref)
            differing += int(ne.sum())
            if res <= 16:
                assert not ne.any(), (c['rule'], s, side)
            # the Riesz-Caputo constant 1/2 gamma(2 - alpha) (equation.py:23) scales the left weights' differences: > 1 for alpha -> 2
            tol_side = tol * (max(1., math.gamma(2. - a)) if side == 'riesz_caputo' else 1.)
            assert np.all(np.abs(st.weights_array - ref) <= tol_side), (c['rule'], s, side)
            if 'nodes' in rec:      # node coordinates: the same Points
                assert set(st._weights) == {Point(x) for x in unhex(rec['nodes'])}
    assert differing <= 0.005 * total, (differing, total)


# ---- (c) batched C-ABI call vs the C oracle on seeded random settings ------------------------------

def test_batched_stencil_weights_vs_c_oracle():
    rng = np.random.default_rng(2024)
    for rule in ('caputo', 'rectangle', 'trapezoidal', 'simpson'):
        n = 300
        alpha = rng.uniform(0.02, 0.98, n)
        res = rng.integers(2, 120, n).astype(np.float64)
        lf = rng.uniform(0.1, 5.0, n)
        for side in ('left', 'right', 'riesz_caputo'):
            row_ptr, offsets, weights = batched.stencil_weights(rule, side, alpha, lf, res)
            offsets, weights = offsets.cpu().numpy(), weights.cpu().numpy()
            exact = 0
            for s in range(n):
                fn = {'left': clib.left_arrays, 'right': clib.right_arrays, 'riesz_caputo': clib.riesz_caputo_arrays}[side]
                first, w = fn(rule, (alpha[s], lf[s], res[s]))
                lo, hi = row_ptr[s], row_ptr[s + 1]
                assert hi - lo == len(w)
                assert np.array_equal(offsets[lo:hi], np.arange(first, first + len(w)))     # integer layout: exact
                assert np.all(np.abs(weights[lo:hi] - w) <= weight_tolerance(rule, alpha[s], lf[s], res[s]))
                exact += np.array_equal(weights[lo:hi], w)
            assert exact >= 0.9 * n, (rule, side, exact)


def test_caputo_alpha_between_one_and_two_vs_c_oracle():
    # n = 2 branch of equation.py:22,44-45 (not covered by the reference's tests)
    for alpha in (1.0, 1.25, 1.5, 1.99):
        for res in (1, 2, 5, 32):
            s = Settings(alpha, 0.7 * res, res)
            for fn, orc in ((fr.create_left_stencil, clib.left_arrays), (fr.create_riesz_caputo_stencil, clib.riesz_caputo_arrays)):
                st = fn('caputo', s)
                first, w = orc('caputo', s)
                assert st.first_offset == first
                assert np.array_equal(st.weights_array, w), (alpha, res)


def test_grunwald_letnikov_extension_vs_mpmath():
    """EXTENSION key 'grunwald_letnikov' (not a key of the reference): prefix-product kernel in
    double-double against an exact mpmath evaluation of the recurrence."""
    from oracle import gl_exact
    for alpha, lf, p in ((0.5, 1.0, 10), (0.3, 0.7, 257), (1.6, 2.0, 1500), (0.9, 1.0, 5000)):
        st = fr.create_left_stencil('grunwald_letnikov', Settings(alpha, lf, p))
        first, want = gl_exact.left_arrays(alpha, lf, p)
        assert st.first_offset == first == -p and len(st.weights_array) == p + 1
        err = np.abs(st.weights_array - want)
        assert np.all(err <= np.spacing(np.abs(want))), (alpha, p, err.max())
        assert np.mean(st.weights_array == want) > 0.98
    with pytest.raises(ValueError):
        fr.create_right_stencil('grunwald_letnikov', Settings(0.5, 1.0, 8))


def test_large_single_stencil_and_float_resolution():
    """a 20001-node stencil through the API path, and resolution given as a float like the reference's own
    fr.Settings(alpha, lf, lf) (test/integration/test_equation.py:91)."""
    s = Settings(0.45, 2.0, 20000)
    st = fr.create_left_stencil('trapezoidal', s)
    first, w = clib.left_arrays('trapezoidal', s)
    assert st.first_offset == first and len(st.weights_array) == len(w) == 20001
    assert np.all(np.abs(st.weights_array - w) <= weight_tolerance('trapezoidal', 0.45, 2.0, 20000))
    assert np.mean(st.weights_array == w) > 0.99
    a = fr.create_riesz_caputo_stencil('caputo', Settings(0.8, 4, 4.0))
    b = fr.create_riesz_caputo_stencil('caputo', Settings(0.8, 4, 4))
    assert a == b
    with pytest.raises(ValueError):
        fr.create_left_stencil('caputo', Settings(0.5, 1.0, 4.5))


def test_stencil_composition_equals_operator_expansion():
    """(f)(1) frx_stencil_compose: d/dx o D^alpha_RC o d/dx composed on the device in two steps, against the oracle's
    dict-merging expansion Operator(central, Operator(rc, Operator(central))) -- same terms in the same order: bit-exact."""
    from oracle import assemble as oasm, fdm_shim
    cases = [('caputo', 0.5, 0.25, 4), ('caputo', 0.8, 0.1, 7), ('trapezoidal', 0.3, 0.05, 12), ('rectangle', 0.6, 0.5, 3),
             ('simpson', 0.4, 0.125, 6), ('caputo', 1.5, 0.2, 5)]
    items, want = [], []
    for rule, alpha, h, res in cases:
        rc = oasm.rc_stencil(rule, alpha, h, res)
        central = fdm_shim.Stencil.central(h)
        op = fdm_shim.Operator(central, fdm_shim.Operator(rc, fdm_shim.Operator(central)))
        exp = op.expand(fdm_shim.Point(0.))
        want.append({int(round(node.x / (h / 2.))): w for node, w in exp.items()})
        first, w = clib.riesz_caputo_arrays(rule, (alpha, res * h, res))
        rc_off = 2 * (first + np.arange(len(w)))                   # units of h/2
        items.append((np.array([-1, 1]), np.array([-1. / h, 1. / h]), rc_off, np.asarray(w, dtype=np.float64)))

    def ragged(parts):
        ptr = np.zeros(len(parts) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum([len(p) for p in parts])
        return ptr, np.concatenate(parts)

    # step 1: B1 = rc o central   (outer = rc in its own node order, inner = central)
    pa, oa = ragged([it[2] for it in items]); _, wa = ragged([it[3] for it in items])
    pb, ob = ragged([it[0] for it in items]); _, wb = ragged([it[1] for it in items])
    p1, o1, w1 = batched.stencil_compose(pa, oa, wa, pb, ob, wb)
    # step 2: C = central o B1
    p2, o2, w2 = batched.stencil_compose(pb, ob, wb, p1, o1, w1.cpu().numpy())
    w2 = w2.cpu().numpy()
    for k in range(len(cases)):
        got = {int(o): float(w) for o, w in zip(o2[p2[k]:p2[k + 1]], w2[p2[k]:p2[k + 1]])}
        assert set(got) == set(want[k]), cases[k]
        for o in got:
            assert got[o] == want[k][o], (cases[k], o, got[o], want[k][o])      # bit-exact
    with pytest.raises(ValueError):
        batched.stencil_compose([0, 2], [0, 1], [1., 1.], [0, 2], [1, 0], [1., 1.])     # B not ascending


def test_batched_stencil_creators_equal_single_calls():
    """create_*_stencils (one K5 launch for a list of Settings) return the stencils of the single-setting functions."""
    import fractulus_b200 as frb
    settings = [Settings(0.5, 0.6, 4), Settings(0.3, 2.0, 7), Settings(0.9, 1.0, 64), Settings(1.5, 1.4, 2), Settings(0.2, 3.3, 33)]
    for many, one in ((frb.create_left_stencils, fr.create_left_stencil), (frb.create_right_stencils, fr.create_right_stencil),
                      (frb.create_riesz_caputo_stencils, fr.create_riesz_caputo_stencil)):
        for rule in ('caputo', 'trapezoidal', 'simpson'):
            got = many(rule, settings)
            assert len(got) == len(settings)
            for st, s in zip(got, settings):
                ref = one(rule, s)
                assert st == ref and st.first_offset == ref.first_offset
                assert np.array_equal(st.weights_array, ref.weights_array)
    assert frb.create_left_stencils('caputo', []) == []
    with pytest.raises(KeyError):
        frb.create_left_stencils('nope', settings)
    with pytest.raises(ZeroDivisionError):
        frb.create_left_stencils('caputo', [Settings(0.5, 1., 4), Settings(0.5, 1., 0)])
