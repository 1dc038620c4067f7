"""The reference's OWN test files, unmodified, against the drop-in ("the same 13 tests pass with
the new package aliased as `fractulus`"): `fractulus` -> fractulus_b200, `fdm` -> the carrier (the real fdm package is
absent everywhere), then reference fractulus/test/unit/test_equation.py:8-99 and
fractulus/test/integration/test_equation.py:10-153 are loaded by path and run under unittest.

The test files come from baseline/_ref (an unmodified copy made by oracle/make_ref.py, shipped to the GPU box) or
from /root/reference where that exists; without either the test is skipped."""
import importlib.util
import os
import sys
import types
import unittest

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_test_files():
    for root in (os.path.join(REPO, 'baseline', '_ref'), '/root/reference'):
        files = [os.path.join(root, 'fractulus', 'test', kind, 'test_equation.py') for kind in ('unit', 'integration')]
        if all(os.path.isfile(f) for f in files):
            return files
    return None


def _fake_fdm():
    """A module tree named fdm WITHOUT the shim marker: what fractulus_b200.equation sees when the real package exists."""
    from fractulus_b200 import fdm_carrier
    made = []

    class Stencil(fdm_carrier.Stencil):
        def __init__(self, weights, *a, **k):
            made.append(len(weights))
            super().__init__(weights, *a, **k)

    fdm = types.ModuleType('fdm')
    eq, geo = types.ModuleType('fdm.equation'), types.ModuleType('fdm.geometry')
    eq.Stencil, eq.Number, geo.Point = Stencil, fdm_carrier.Number, fdm_carrier.Point
    fdm.equation, fdm.geometry = eq, geo
    fdm.__path__ = []
    return {'fdm': fdm, 'fdm.equation': eq, 'fdm.geometry': geo}, Stencil, made


def test_carrier_selection_prefers_a_real_fdm_package(monkeypatch):
    """fractulus_b200/equation.py:_carrier -- real `fdm` importable => fdm.equation.Stencil / fdm.geometry.Point are what
    create_*_stencil builds (reference equation.py:4-5); a module marked __fdm_shim__ or none => the built-in carrier."""
    from fractulus_b200 import equation, fdm_carrier
    mods, Stencil, _ = _fake_fdm()
    for k, v in mods.items():
        monkeypatch.setitem(sys.modules, k, v)
    monkeypatch.delenv('FRACTULUS_B200_FDM', raising=False)
    st, pt, builtin = equation._carrier()
    assert st is Stencil and pt is fdm_carrier.Point and builtin is False
    monkeypatch.setenv('FRACTULUS_B200_FDM', 'builtin')
    assert equation._carrier() == (fdm_carrier.Stencil, fdm_carrier.Point, True)
    monkeypatch.delenv('FRACTULUS_B200_FDM')
    mods['fdm'].__fdm_shim__ = True
    assert equation._carrier()[2] is True


@pytest.mark.gpu
def test_real_fdm_return_path_builds_fdm_stencils(monkeypatch):
    """The branch nobody exercised in round 1: with an importable `fdm`, create_*_stencil returns fdm.equation.Stencil
    built from {fdm.geometry.Point: float} -- same nodes and weights as the carrier path."""
    import fractulus_b200 as fr
    mods, Stencil, made = _fake_fdm()
    want = fr.create_riesz_caputo_stencil('caputo', fr.Settings(0.5, 0.6, 4))
    for k, v in mods.items():
        monkeypatch.setitem(sys.modules, k, v)
    got = fr.create_riesz_caputo_stencil('caputo', fr.Settings(0.5, 0.6, 4))
    assert type(got) is Stencil and made == [9]
    assert got._weights == want._weights and all(type(w) is float for w in got._weights.values())


@pytest.mark.gpu
def test_reference_own_tests_pass_against_the_dropin(monkeypatch):
    files = _reference_test_files()
    if files is None:
        pytest.skip('no copy of the reference test files (baseline/_ref: python -m oracle.make_ref)')
    import fractulus_b200
    import fractulus_b200.equation
    from fractulus_b200 import fdm_carrier
    saved = {k: sys.modules.get(k) for k in ('fdm', 'fdm.equation', 'fdm.geometry', 'fractulus', 'fractulus.equation')}
    try:
        fdm_carrier.install_as_fdm()
        sys.modules['fractulus'] = fractulus_b200
        sys.modules['fractulus.equation'] = fractulus_b200.equation
        suite = unittest.TestSuite()
        for i, path in enumerate(files):
            spec = importlib.util.spec_from_file_location('reference_test_equation_%d' % i, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            suite.addTests(unittest.defaultTestLoader.loadTestsFromModule(mod))
        result = unittest.TextTestRunner(verbosity=0).run(suite)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert result.testsRun == 13, result.testsRun
    assert result.wasSuccessful(), (result.failures, result.errors)
