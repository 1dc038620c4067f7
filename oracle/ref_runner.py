"""TEST / BENCH INFRASTRUCTURE -- runs the UNMODIFIED reference (baseline/_ref, see oracle/make_ref.py) on top of
oracle/fdm_shim.py for bench.py's CPU legs: the reference's own public API, the call pattern of its own tests
(reference fractulus/test/integration/test_equation.py:48-66), on all host cores through multiprocessing.

No torch, no CUDA in here (worker processes are spawned and must stay light)."""
import math
import multiprocessing as mp
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_state = {}


def _load():
    if 'fr' not in _state:
        if REPO not in sys.path:
            sys.path.insert(0, REPO)
        from oracle import make_ref
        fr, fdm = make_ref.load()
        _state['fr'], _state['fdm'] = fr, fdm
    return _state['fr'], _state['fdm']


def available():
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from oracle import make_ref
    return make_ref.available()


def smooth_field(n_nodes, h):
    # g = f' for f = sin 3x + x^2 (SURVEY.md section 8(d) cfg 2)
    return [3. * math.cos(3. * (j * h)) + 2. * (j * h) for j in range(n_nodes)]


def _init_worker(n_nodes, h):
    _load()
    _state['g'] = smooth_field(n_nodes, h)


def _row(args):
    """one history row through the reference API: stencil of Settings(alpha, i*h, i), then sum(g(node) * weight)"""
    rule, alpha, h, i = args
    fr = _state['fr']
    g = _state['g']
    st = fr.create_left_stencil(rule, fr.Settings(alpha, i * h, i))
    hh = (i * h) / float(i)
    return sum([g[i + int(round(pt.x / hh))] * w for pt, w in st._weights.items()])


class RowPool(object):
    """Persistent pool of reference workers for one (N, h): rows(rule, alpha, row_list) -> (seconds, values)."""

    def __init__(self, N, h, processes=None):
        self.N, self.h = N, h
        self.processes = processes or len(os.sched_getaffinity(0))
        ctx = mp.get_context('spawn')
        self.pool = ctx.Pool(self.processes, initializer=_init_worker, initargs=(N + 2, h))
        self.pool.map(_noop, range(self.processes * 2))      # workers up and the reference imported before any timing

    def rows(self, rule, alpha, rows):
        jobs = [(rule, alpha, self.h, int(i)) for i in sorted(rows, reverse=True)]     # longest first
        t0 = time.perf_counter()
        ys = self.pool.map(_row, jobs, chunksize=1)
        dt = time.perf_counter() - t0
        return dt, dict(zip([j[3] for j in jobs], ys))

    def close(self):
        self.pool.close()
        self.pool.join()


def _noop(_):
    return 0


def cfg1(alpha=0.5, N=1000, h=1e-3):
    """BASELINE configs[0] exactly: D^alpha of x^2 on N nodes, one core, through
    Operator(create_left_stencil('caputo', Settings(alpha, i*h, i)), Operator(Stencil.central(h))).expand(Point(i*h))
    and sum(f(node) * weight) (reference test/integration/test_equation.py:51-66).  -> (seconds, [y_1..y_N])"""
    fr, fdm = _load()
    f = lambda pt: pt.x ** 2
    t0 = time.perf_counter()
    ys = []
    for i in range(1, N + 1):
        op = fdm.Operator(fr.create_left_stencil('caputo', fr.Settings(alpha, i * h, i)), fdm.Operator(fdm.Stencil.central(h)))
        scheme = op.expand(fdm.Point(i * h))
        ys.append(sum([f(node) * weight for node, weight in scheme.items()]))
    return time.perf_counter() - t0, ys


def create_stencil_latency(points, repeats=3):
    """seconds per create_left_stencil('caputo', Settings(0.5, p*0.01, p)) call for every p in points (best of repeats)"""
    fr, _ = _load()
    out = {}
    for p in points:
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            fr.create_left_stencil('caputo', fr.Settings(0.5, p * 0.01, p))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out[int(p)] = best
    return out
