"""TEST INFRASTRUCTURE -- generate golden fixtures from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, which is absent on the GPU box).
Imports /root/reference/fractulus/equation.py as-is on top of oracle/fdm_shim.py (the
reference's ``fdm`` dependency is absent), then writes tests/golden/*.json with every
float stored as ``float.hex()`` so the fixtures are bit-exact.

    python -m oracle.gen_golden --selftest     # reference's own 13 tests + port-vs-reference bit check
    python -m oracle.gen_golden                # (re)write tests/golden/
"""
import argparse
import json
import math
import os
import random
import sys
import time
import unittest

sys.dont_write_bytecode = True
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def load_reference():
    sys.path.insert(0, REPO)
    from oracle import fdm_shim
    fdm = fdm_shim.install_as_fdm()
    if REF not in sys.path:
        sys.path.insert(1, REF)
    import fractulus  # the unmodified reference
    assert os.path.realpath(fractulus.__file__).startswith(REF), fractulus.__file__
    return fractulus, fdm, fdm_shim


def stencil_record(stencil, h):
    """Sorted (integer offset, node coordinate, weight) triples of a reference stencil."""
    items = sorted(stencil._weights.items(), key=lambda kv: kv[0].x)
    offs = [int(round(pt.x / h)) for pt, _ in items]
    assert offs == list(range(offs[0], offs[0] + len(offs))), offs
    ws = [w for _, w in items]
    assert all(isinstance(w, float) for w in ws)
    rec = {'first_offset': offs[0], 'count': len(offs), 'weights': [w.hex() for w in ws]}
    if len(offs) <= 40:   # node coordinates only for small stencils (keeps the fixture small)
        rec['nodes'] = [float(pt.x).hex() for pt, _ in items]
    return rec


def stencil_cases():
    rnd = random.Random(20171)
    cases = []
    # the reference's own unit-test settings (fractulus/test/unit/test_equation.py)
    cases += [('caputo', 0.5, 0.6, 4), ('rectangle', 0.5, 0.8, 4), ('trapezoidal', 0.5, 0.8, 4),
              ('simpson', 0.5, 0.8, 4), ('simpson', 0.5, 1.0, 5), ('caputo', 0.999999, 1., 1),
              ('caputo', 0.9999, 10., 10)]
    for rule in ('caputo', 'rectangle', 'trapezoidal', 'simpson'):
        lo = 2 if rule in ('rectangle', 'simpson') else 1
        for res in (lo, lo + 1, 3, 4, 7, 8, 16, 31, 64, 255):
            for _ in range(3):
                alpha = round(rnd.uniform(0.01, 0.99), rnd.choice((1, 2, 6, 15)))
                lf = rnd.choice((1.0, 0.1 * res, rnd.uniform(0.05, 7.0), float(res)))
                cases.append((rule, alpha, lf, res))
        cases.append((rule, 0.5, 1.0, 1000))
        cases.append((rule, 0.37, 2.5, 1001))
    # alpha in (1, 2): n = 2 branch of 'caputo' (equation.py:44-45)
    for res in (1, 2, 5, 32):
        for alpha in (1.0, 1.25, 1.5, 1.99):
            cases.append(('caputo', alpha, 0.7 * res, res))
    # (appended in round 2, after the cases above so that their random draws stay what they were)
    # 'trapezoidal' and 'simpson' also give real weights for alpha in [1, 2): every pow base is >= 0 for
    # resolution >= 2 and (1.-alpha)**2 has an integer exponent (equation.py:84-176); 'trapezoidal' up to alpha = 2
    for res in (2, 3, 4, 5, 8, 33, 64):
        for alpha in (1.0, 1.25, 1.5, 1.99):
            cases.append(('trapezoidal', alpha, 0.7 * res, res))
            cases.append(('simpson', alpha, 0.7 * res, res))
    # alpha = 2: gamma(2. - alpha) of the Riesz-Caputo constant is at a pole (ValueError, equation.py:23) -> recorded as 'raises'
    cases += [('trapezoidal', 2.0, 1.0, 4), ('trapezoidal', 1.5, 1.0, 1), ('caputo', 2.0, 2.0, 5)]
    return cases


def gen_stencils(fr):
    out = []
    for rule, alpha, lf, res in stencil_cases():
        s = fr.Settings(alpha, lf, res)
        h = lf / float(res)
        rec = {'rule': rule, 'alpha': float(alpha).hex(), 'lf': float(lf).hex(), 'resolution': res,
               'left': stencil_record(fr.create_left_stencil(rule, s), h),
               'right': stencil_record(fr.create_right_stencil(rule, s), h)}
        try:
            rec['riesz_caputo'] = stencil_record(fr.create_riesz_caputo_stencil(rule, s), h)
        except ValueError as ex:
            rec['riesz_caputo'] = {'raises': type(ex).__name__}
        out.append(rec)
    return out


def history_row_reference(fr, rule, alpha, h, g, i):
    """Pattern of reference test/integration/test_equation.py:51-57 with the inner element a
    sampled field: sum(g(node) * weight)."""
    st = fr.create_left_stencil(rule, fr.Settings(alpha, i * h, i))
    hh = (i * h) / float(i)
    return sum([g[i + int(round(pt.x / hh))] * w for pt, w in st._weights.items()])


def smooth_field(n_nodes, h):
    # g = f' for f = sin 3x + x^2 (SURVEY.md section 8(d) cfg 2)
    return [3. * math.cos(3. * (j * h)) + 2. * (j * h) for j in range(n_nodes)]


def gen_history(fr, fdm):
    out = []
    # cfg 1 in full: N = 1000, alpha = 0.5, all four rules, every row
    N, h = 1000, 1e-3
    g = smooth_field(N + 2, h)
    for rule in ('caputo', 'rectangle', 'trapezoidal', 'simpson'):
        lo = 3 if rule == 'simpson' else (2 if rule == 'rectangle' else 1)
        rows = list(range(lo, N + 1))
        ys = [history_row_reference(fr, rule, 0.5, h, g, i) for i in rows]
        out.append({'rule': rule, 'alpha': (0.5).hex(), 'h': h.hex(), 'N': N, 'field': 'smooth',
                    'rows': rows, 'y': [y.hex() for y in ys]})
    # alpha in (1, 2) (n = 2 branch, equation.py:22,44-45,49,54), every row
    for rule, alphas in (('caputo', (1.25, 1.5, 1.9)), ('trapezoidal', (1.5,)), ('simpson', (1.5,))):
        for alpha in alphas:
            lo = 3 if rule == 'simpson' else 1
            rows = list(range(lo, N + 1))
            ys = [history_row_reference(fr, rule, alpha, h, g, i) for i in rows]
            out.append({'rule': rule, 'alpha': float(alpha).hex(), 'h': h.hex(), 'N': N, 'field': 'smooth',
                        'rows': rows, 'y': [y.hex() for y in ys]})
    # cfg 2 sampled rows: N = 1e5
    N, h = 100000, 1e-5
    g = smooth_field(N + 2, h)
    rows = sorted(set([1, 2, 3, 4, 5, 7, 8, 9, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 4097, 50001, 77777]
                      + [N // 32 * k for k in range(1, 33)] + [N - 1]))
    for rule, alphas in (('caputo', (0.1, 0.5, 0.9, 0.8444218515250481, 1.5)), ('trapezoidal', (0.5,)),
                         ('rectangle', (0.5,)), ('simpson', (0.5,))):
        for alpha in alphas:
            lo = 3 if rule == 'simpson' else (2 if rule == 'rectangle' else 1)
            rr = [i for i in rows if i >= lo]
            t0 = time.time()
            ys = [history_row_reference(fr, rule, alpha, h, g, i) for i in rr]
            out.append({'rule': rule, 'alpha': float(alpha).hex(), 'h': h.hex(), 'N': N, 'field': 'smooth',
                        'rows': rr, 'y': [y.hex() for y in ys], 'ref_seconds': round(time.time() - t0, 2)})
            print('history', rule, alpha, 'rows', len(rr), 'sec', round(time.time() - t0, 1), flush=True)
    return out


def gen_cfg1_api(fr, fdm):
    """cfg 1 exactly as SURVEY.md section 8(d): Operator(left('caputo'), Operator(central(h))) on x^2 and x."""
    N, h, alpha = 1000, 1e-3, 0.5
    out = []
    for name, f in (('x2', lambda pt: pt.x ** 2), ('x1', lambda pt: pt.x)):
        t0 = time.time()
        ys = []
        for i in range(1, N + 1):
            op = fdm.Operator(fr.create_left_stencil('caputo', fr.Settings(alpha, i * h, i)),
                              fdm.Operator(fdm.Stencil.central(h)))
            scheme = op.expand(fdm.Point(i * h))
            ys.append(sum([f(node) * weight for node, weight in scheme.items()]))
        out.append({'f': name, 'alpha': alpha.hex(), 'h': h.hex(), 'N': N, 'y': [y.hex() for y in ys],
                    'ref_seconds': round(time.time() - t0, 2)})
        print('cfg1', name, round(time.time() - t0, 1), 's', flush=True)
    return out


def selftest(fr, fdm):
    suite = unittest.defaultTestLoader.discover(os.path.join(REF, 'fractulus', 'test'), top_level_dir=REF)
    res = unittest.TextTestRunner(verbosity=1).run(suite)
    assert res.wasSuccessful() and res.testsRun == 13, res
    from oracle import ref_port
    bad = n = 0
    for rule, alpha, lf, res_ in stencil_cases():
        s = fr.Settings(alpha, lf, res_)
        h = lf / float(res_)
        for side, ref_fn, port_fn in (('left', fr.create_left_stencil, ref_port.left_arrays),
                                      ('right', fr.create_right_stencil, ref_port.right_arrays),
                                      ('rc', fr.create_riesz_caputo_stencil, ref_port.riesz_caputo_arrays)):
            try:
                rec = stencil_record(ref_fn(rule, s), h)
            except ValueError:
                try:
                    port_fn(rule, s)
                    bad += 1
                    print('MISMATCH (port does not raise)', rule, alpha, lf, res_, side)
                except ValueError:
                    pass
                n += 1
                continue
            first, ws = port_fn(rule, s)
            n += 1
            if first != rec['first_offset'] or [w.hex() for w in ws] != rec['weights']:
                bad += 1
                print('MISMATCH', rule, alpha, lf, res_, side)
    print('port vs unmodified reference: %d stencils, %d mismatching (must be 0)' % (n, bad))
    assert bad == 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--selftest', action='store_true')
    ap.add_argument('--only', default='', help='comma list of stencils,cfg1_api,history')
    args = ap.parse_args()
    fr, fdm, _ = load_reference()
    if args.selftest:
        selftest(fr, fdm)
        return
    os.makedirs(GOLDEN, exist_ok=True)
    meta = {'generator': 'oracle/gen_golden.py', 'reference': 'szajek/fractulus fractulus/equation.py (unmodified) + oracle/fdm_shim.py',
            'python': sys.version.split()[0], 'floats': 'float.hex()'}
    gens = (('stencils.json', lambda: gen_stencils(fr)), ('cfg1_api.json', lambda: gen_cfg1_api(fr, fdm)),
            ('history.json', lambda: gen_history(fr, fdm)))
    only = [x for x in args.only.split(',') if x]
    for name, gen in gens:
        if only and name[:-5] not in only:
            continue
        payload = gen()
        with open(os.path.join(GOLDEN, name), 'w') as f:
            json.dump({'meta': meta, 'cases': payload}, f, separators=(',', ':'))
        print('wrote', name, os.path.getsize(os.path.join(GOLDEN, name)), 'bytes')


if __name__ == '__main__':
    main()
