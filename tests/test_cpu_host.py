"""CPU suite for the product's host side: the C-ABI library loads and exports every symbol that
include/frx.h declares, the integer layout logic (no GPU work) matches the unmodified reference
bit-exactly, errors map to the reference's exception types, the carrier objects behave like the
reference's call sites need, and the FP64 scalar math of the kernels (compiled for the host by a
test helper -- same explicitly rounded operations) is correctly rounded."""
import ctypes
import math
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import REPO, unhex

import fractulus_b200 as fr
from fractulus_b200 import _lib, fdm_carrier


def _header_symbols():
    text = open(os.path.join(REPO, 'include', 'frx.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(frx_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 14
    L = _lib.lib()
    for s in syms:
        assert hasattr(L, s), s
        assert s in _lib.SIGNATURES, 'ctypes signature missing for ' + s
    assert L.frx_version() == 100


def test_public_names_match_reference():
    assert fr.equation.__all__ == ('Settings', 'create_left_stencil', 'create_right_stencil',
                                   'create_riesz_caputo_stencil')          # reference equation.py:7
    assert fr.Settings.__name__ == 'CaputoSettings'                         # reference equation.py:10
    assert fr.Settings._fields == ('alpha', 'lf', 'resolution')
    for name in fr.equation.__all__:
        assert hasattr(fr, name)                                            # reference __init__.py:1


def test_layout_matches_unmodified_reference(golden_stencils):
    L = _lib.lib()
    first, count = ctypes.c_int32(), ctypes.c_int32()
    for c in golden_stencils:
        for side in ('left', 'right', 'riesz_caputo'):
            rc = L.frx_stencil_layout(_lib.RULES[c['rule']], _lib.SIDES[side], float.fromhex(c['alpha']),
                                      float(c['resolution']), ctypes.byref(first), ctypes.byref(count))
            if 'raises' in c[side]:          # alpha = 2: gamma(2. - alpha) at a pole (reference: ValueError, equation.py:23)
                assert rc == -5
                continue
            assert rc == 0
            assert (first.value, count.value) == (c[side]['first_offset'], c[side]['count']), (c['rule'], side)


def test_error_mapping_without_gpu_work():
    S = fr.Settings
    with pytest.raises(KeyError):
        fr.create_left_stencil('grunwald', S(0.5, 1., 4))       # equation.py:29
    with pytest.raises(ZeroDivisionError):
        fr.create_left_stencil('caputo', S(0.5, 1., 0))         # equation.py:56
    with pytest.raises(ZeroDivisionError):
        fr.create_left_stencil('rectangle', S(1.5, 1., 4))      # equation.py:68
    with pytest.raises((TypeError, ValueError)):
        fr.create_left_stencil('caputo', (0.5, 1.))             # equation.py:43 unpack
    with pytest.raises(ValueError):
        fr.create_left_stencil('simpson', S(0.5, 1., 1))        # complex weights in the reference


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        fr.create_left_stencil('caputo', fr.Settings(0.5, 0.6, 4))
    from fractulus_b200 import batched
    with pytest.raises(RuntimeError):
        batched.history_host('caputo', 0.5, 1e-3, np.zeros(10))
    with pytest.raises(TypeError):
        batched.history('caputo', 0.5, 1e-3, torch.zeros(10, dtype=torch.float64))   # CPU tensor: rejected


def test_carrier_semantics_from_reference_call_sites():
    P, St, Op, Num = fdm_carrier.Point, fdm_carrier.Stencil, fdm_carrier.Operator, fdm_carrier.Number
    assert P(-0.6000000000000001) == P(-0.6) and hash(P(-0.6000000000000001)) == hash(P(-0.6))
    assert P(-1.1102230246251565e-16) == P(0.) == P(-0.0)
    c = St.central(.1)
    assert c._weights == {P(-.05): -10., P(.05): 10.}
    s = St({P(-1.): 2., P(0.): 3.})
    assert s.expand(P(5.)) == {P(4.): 2., P(5.): 3.}
    assert s.symmetry(P(0.)).multiply(-1.)._weights == {P(1.): -2., P(0.): -3.}
    op = Op(s, Op(St.central(2.)))
    assert op.expand(P(0.)) == {P(-2.): -1., P(-1.): -1.5, P(0.): 1., P(1.): 1.5}
    rc = (Num(0.5) * (s + Num(-1.) * s.symmetry(P(0.)).multiply(-1.))).to_stencil(P(0.))
    assert rc._weights == {P(-1.): 1., P(0.): 3., P(1.): 1.}
    assert Op(s, Num(lambda p: p.x)).expand(P(2.)) == {P(1.): 2., P(2.): 6.}


# ---- FP64 scalar math of the kernels, host build -------------------------------------------------

@pytest.fixture(scope='module')
def hostmath():
    out = os.path.join(REPO, 'tests', '_build', 'libfrx_math_host.so')
    src = os.path.join(REPO, 'tests', 'helpers', 'frx_math_host.cpp')
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(['/usr/bin/g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-o', out,
                           src, '-lm'])
    L = ctypes.CDLL(out)
    L.t_pow_cr.restype = L.t_exp_cr.restype = L.t_gamma_py.restype = ctypes.c_double
    L.t_pow_cr.argtypes = [ctypes.c_double, ctypes.c_double]
    L.t_exp_cr.argtypes = L.t_gamma_py.argtypes = [ctypes.c_double]
    dp = ctypes.POINTER(ctypes.c_double)
    L.t_left_weights.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                 ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), dp]
    return L


def test_pow_is_correctly_rounded(hostmath):
    mp = pytest.importorskip('mpmath')
    mp.mp.prec = 400
    rng = np.random.default_rng(11)
    n_glibc_not_cr = 0
    for k in range(12000):
        if k % 3 == 0:
            x, y = float(rng.integers(1, 100003)), rng.uniform(0., 3.)
        elif k % 3 == 1:
            x, y = rng.uniform(1e-7, 1.), rng.uniform(0., 2.)
        else:
            x, y = rng.uniform(5.5, 10.), rng.uniform(0.5, 4.)
        exact = float(mp.power(mp.mpf(x), mp.mpf(y)))
        assert hostmath.t_pow_cr(x, y) == exact, (x, y)
        n_glibc_not_cr += math.pow(x, y) != exact
    # the reason GPU and reference weights can differ at all: glibc's pow is < 1 ULP, not CR
    assert n_glibc_not_cr < 60
    for x, y, want in ((4., .5, 2.), (4., 1.5, 8.), (0., 1.5, 0.), (7., 0., 1.), (1., 2.5, 1.), (1e5, 2., 1e10)):
        assert hostmath.t_pow_cr(x, y) == want
    for x in rng.uniform(-30, 30, 3000):
        assert hostmath.t_exp_cr(x) == float(mp.exp(mp.mpf(x)))


def test_gamma_is_cpython_lanczos(hostmath):
    rng = np.random.default_rng(5)
    xs = rng.uniform(1.0, 4.0, 20000)
    diff = [(x, hostmath.t_gamma_py(x), math.gamma(x)) for x in xs if hostmath.t_gamma_py(x) != math.gamma(x)]
    assert len(diff) < 100                                   # only where glibc exp/pow are not CR
    assert all(abs(a - b) <= 2 * np.spacing(b) for _, a, b in diff)
    for x in (1., 2., 3., 4.):
        assert hostmath.t_gamma_py(x) == math.gamma(x)


def test_rule_functions_vs_unmodified_reference(hostmath, golden_stencils):
    """Same source as the kernels, compiled for the host: weights are bit-identical to the unmodified
    reference except where glibc's pow is not correctly rounded; small stencils are all exact."""
    dp = ctypes.POINTER(ctypes.c_double)
    total = differing = 0
    for c in golden_stencils:
        a, lf, res = float.fromhex(c['alpha']), float.fromhex(c['lf']), c['resolution']
        w = np.zeros(res + 3)
        f, n = ctypes.c_int(), ctypes.c_int()
        hostmath.t_left_weights(_lib.RULES[c['rule']], a, lf, float(res), ctypes.byref(f), ctypes.byref(n),
                                w.ctypes.data_as(dp))
        ref = np.array(unhex(c['left']['weights']))
        assert (f.value, n.value) == (c['left']['first_offset'], c['left']['count'])
        got = w[:n.value]
        ne = got != ref
        total += len(ref)
        differing += int(ne.sum())
        if res <= 16:
            assert not ne.any(), (c['rule'], a, lf, res)
        assert np.all(np.abs(got - ref) <= weight_tolerance(c['rule'], a, lf, res)), (c['rule'], a, lf, res)
    assert differing <= 0.005 * total


def weight_tolerance(rule, alpha, lf, res):
    """|w_gpu - w_ref| bound: a few ulp of the largest power entering the cancelling sum, times the
    multiplier."""
    n = math.floor(alpha) + 1.
    emax = {'caputo': n - alpha + 1., 'trapezoidal': 2. - alpha, 'rectangle': 1. - alpha, 'simpson': 3. - alpha}[rule]
    mult = (lf / res) ** ({'caputo': n - alpha}.get(rule, 1. - alpha))
    return 8 * np.finfo(float).eps * mult * (res + 2.) ** emax


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under fractulus_b200/ (Python or CUDA) may import, include or
    load anything from oracle/, and the shared library must not link it."""
    pkg = os.path.join(REPO, 'fractulus_b200')
    for root, _, files in os.walk(pkg):
        if '_build' in root or '__pycache__' in root:
            continue
        for name in files:
            if name.endswith(('.py', '.cu', '.cuh', '.h', '.inc')):
                for line in open(os.path.join(root, name)):
                    code = line.split('//')[0] if name.endswith(('.cu', '.cuh', '.h', '.inc')) else line.split('#')[0]
                    assert not re.search(r'\b(import|from|include|CDLL)\b.*oracle', code), (name, line)
    out = subprocess.run(['ldd', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'oracle' not in out


def test_host_side_argument_handling_without_gpu():
    """Host logic of the batched layer that runs before any device work."""
    from fractulus_b200 import batched
    a = np.linspace(0., 1., 5)
    assert batched._f64_c(a) is a                                        # conforming arrays pass through untouched
    assert batched._f64_c(a[::2]).flags.c_contiguous and batched._f64_c([1, 2]).dtype == np.float64
    assert batched._f64_c(0.5, 1).shape == (1,)
    with pytest.raises(ValueError):                                      # item counts differ
        batched.stencil_compose([0, 1], [0], [1.], [0, 1, 2], [0, 1], [1., 1.])
    with pytest.raises(ValueError):                                      # B's offsets must ascend
        batched.stencil_compose([0, 2], [0, 1], [1., 1.], [0, 2], [1, 0], [1., 1.])


def test_compose_structure_matches_per_item_unique():
    """batched._compose_structure (tensor ops; runs on the device in production) against the per-item np.unique it replaced."""
    import torch
    from fractulus_b200 import batched
    rng = np.random.default_rng(4)
    n = 50
    na, nb = rng.integers(0, 9, n), rng.integers(0, 7, n)
    ptr_a, ptr_b = np.concatenate([[0], np.cumsum(na)]), np.concatenate([[0], np.cumsum(nb)])
    off_a = rng.integers(-20, 20, ptr_a[-1]).astype(np.int32)
    off_b = np.concatenate([np.sort(rng.choice(np.arange(-15, 15), k, replace=False)) for k in nb] + [np.zeros(0)]).astype(np.int32)
    t = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x), dtype=dt)
    ptr_c, off_c = batched._compose_structure(t(ptr_a, torch.int64), t(off_a, torch.int32), t(ptr_b, torch.int64), t(off_b, torch.int32))
    ptr_c, off_c = ptr_c.numpy(), off_c.numpy()
    for k in range(n):
        oa, ob = off_a[ptr_a[k]:ptr_a[k + 1]], off_b[ptr_b[k]:ptr_b[k + 1]]
        want = np.unique(oa[:, None].astype(np.int64) + ob[None, :]) if len(oa) and len(ob) else np.zeros(0)
        assert np.array_equal(off_c[ptr_c[k]:ptr_c[k + 1]], want), k
    with pytest.raises(ValueError):
        batched._compose_structure(t([0, 1], torch.int64), t([0], torch.int32), t([0, 2], torch.int64), t([3, 3], torch.int32))


def test_operator_array_fast_path_equals_generic_expansion():
    """fdm_carrier.Operator._expand_from_arrays (kernel-built stencil o small stencil, the reference's call pattern of
    test/integration/test_equation.py:61-66) gives the generic dict-merging expansion: same nodes, same order, same bits."""
    P, St, Op = fdm_carrier.Point, fdm_carrier.Stencil, fdm_carrier.Operator
    rng = np.random.default_rng(0)
    for p in (1, 2, 5, 40, 333):
        h = 0.001
        lf = p * h
        xs = [-lf + (lf / p) * j for j in range(p + 1)]
        w = rng.standard_normal(p + 1)
        eager = St({P(x): float(v) for x, v in zip(xs, w)})
        for inner in (St.central(h), St.central(2 * h), St({P(-h): 1., P(0.): -2., P(h): 1.})):
            lazy = St.from_layout(lambda xs=xs: list(xs), w, -p, h)
            assert lazy._dict is None
            a = Op(lazy, Op(inner)).expand(P(p * h))
            assert lazy._dict is None                                   # no Point per stencil node was built
            b = Op(eager, Op(inner)).expand(P(p * h))
            assert [q.x for q in a] == [q.x for q in b] and list(a.values()) == list(b.values())
    lazy = St.from_layout(lambda: [0., 1.], np.array([1., 2.]), 0, 1.)
    assert Op(lazy, fdm_carrier.Number(3.)).expand(P(0.)) == {P(0.): 3., P(1.): 6.}      # other elements: generic path


def test_cr_oracle_tables_equal_the_kernel_rule_functions(hostmath):
    """The correctly-rounded-pow build of the C oracle and the product's rule functions (host build of csrc/frx_rules.cuh, the
    source kernel K1 compiles) give bit-identical Toeplitz tables: GPU vs that oracle therefore isolates summation order."""
    from oracle import clib
    dp = ctypes.POINTER(ctypes.c_double)
    hostmath.t_tables.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_long, dp, dp, dp, dp, dp]
    N, h = 20000, 5e-5
    for rule, code, alphas in (('caputo', 0, (0.1, 0.5, 0.9, 1.5)), ('rectangle', 1, (0.5,)), ('trapezoidal', 2, (0.5, 1.25)),
                               ('simpson', 3, (0.5, 1.5))):
        for alpha in alphas:
            cA, cB, w0, mult = np.zeros(N), np.zeros(N), np.zeros(N + 1), np.zeros(N + 1)
            cm1 = ctypes.c_double()
            hostmath.t_tables(code, alpha, h, N, cA.ctypes.data_as(dp), cB.ctypes.data_as(dp), w0.ctypes.data_as(dp),
                              mult.ctypes.data_as(dp), ctypes.byref(cm1))
            o = clib.tables(rule, alpha, h, N, cr=True)
            lo = 2 if rule == 'simpson' else 1          # (row 1 of 'simpson' does not exist: (m-2)**x with m = 1)
            assert np.array_equal(cA, o[0]) and np.array_equal(w0[lo:], o[2][lo:]) and np.array_equal(mult[1:], o[3][1:]), (rule, alpha)
            if rule == 'simpson':
                assert np.array_equal(cB, o[1]) and cm1.value == o[4]
