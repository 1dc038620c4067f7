"""TEST INFRASTRUCTURE -- CPU restatement (pure Python, small cases) of the reference's
fractional-stencil algorithm, organised by *integer distance* instead of by closures.

Follows reference fractulus/equation.py expression for expression (same operand order,
same int/float mixing, Python-float ``**`` == glibc ``pow``, ``math.gamma`` == CPython's
Lanczos ``m_tgamma``) so that every weight is BIT-IDENTICAL to what the unmodified
reference returns.  Pinned by:

* the five 17-digit golden vectors of reference fractulus/test/unit/test_equation.py:20-26,
  47-52,61-67,76-82,89-97 (tests/test_oracle_golden.py);
* bit-exact comparison with the unmodified reference imported from /root/reference on a
  randomised set of settings (oracle/gen_golden.py, fixtures in tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
Pure-Python loops: use for small cases; oracle/frx_oracle.c is the large-N restatement.
"""
import collections
import math

from . import fdm_shim

RULES = ('caputo', 'rectangle', 'trapezoidal', 'simpson')

# reference fractulus/equation.py:10 -- note the namedtuple *type name*.
Settings = collections.namedtuple('CaputoSettings', ('alpha', 'lf', 'resolution'))


def _as_int_resolution(resolution):
    r = int(resolution)
    if r != resolution:
        raise ValueError('resolution must be integral, got %r' % (resolution,))
    return r


# --------------------------------------------------------------------------------------
# raw (un-multiplied) weight of the node at distance d = -(offset) from the evaluation
# point, and the stencil multiplier, per rule.
# --------------------------------------------------------------------------------------

def _caputo(alpha, lf, p):
    """reference fractulus/equation.py:42-59.  Nodes j=0..p at offset j-p."""
    n = math.floor(alpha) + 1.
    index = (n - alpha + 1.)

    def raw(d):                      # d = p - j
        if d == p:                   # node number 0 -> first column, equation.py:49
            return (p - 1.) ** index - (p - n + alpha - 1.) * p ** (n - alpha)
        if d == 0:                   # equation.py:51
            return 1.
        j = p - d                    # equation.py:53-54
        return (p - j + 1.) ** index - 2. * (p - j) ** index + (p - j - 1.) ** index

    h = (0. - (-lf)) / p             # equation.py:56
    mult = ((h ** (n - alpha)) / math.gamma(n - alpha + 2.))
    return -p, [raw(p - j) for j in range(p + 1)], mult


def _rectangle(alpha, lf, res):
    """reference fractulus/equation.py:62-73.  Nodes i=0..res-1 at offset i-(res-1)."""
    if res - 1 == 0:
        # Stencil.uniform(start, end, 0 intervals, ...): behaviour lives in the absent fdm;
        # the shim divides the span by the interval count.
        raise ZeroDivisionError('rectangle rule with resolution 1 has zero intervals')
    raws = []
    for i in range(res):
        k = -res + i                 # equation.py:67-68
        raws.append((-k) ** (1. - alpha) - (-k - 1) ** (1. - alpha))
    mult = (lf / float(res)) ** (1. - alpha) / math.gamma(2. - alpha)
    return -(res - 1), raws, mult


def _trapezoidal(alpha, lf, p):
    """reference fractulus/equation.py:84-100.  Nodes j=0..p at offset j-p."""
    raws = []
    for number in range(p + 1):
        if number == 0:              # equation.py:90
            raws.append((p - 1.) ** (2 - alpha) + (2. - alpha - p) * p ** (1 - alpha))
        elif number == p:
            raws.append(1.)
        else:                        # equation.py:94-95
            k = -p + number
            raws.append((-k + 1.) ** (2 - alpha) - 2 * (-k) ** (2 - alpha) + (-k - 1.) ** (2 - alpha))
    mult = (lf / float(p)) ** (1. - alpha) / math.gamma(3. - alpha)
    return -p, raws, mult


def _simpson(alpha, lf, m):
    """reference fractulus/equation.py:103-176 (evaluation node index i == 0 there).

    even m: nodes j=0..m   at offset k=j-m in [-m, 0]
    odd  m: nodes j=0..m+1 at offset k=j-m in [-m, +1] (one node right of the point).
    """
    dt = lf / float(m)
    m_is_even = math.fmod(m, 2.) == 0.
    g4, g3, g2 = math.gamma(4 - alpha), math.gamma(3 - alpha), math.gamma(2 - alpha)
    g4f, g3f, g2f = math.gamma(4. - alpha), math.gamma(3. - alpha), math.gamma(2. - alpha)

    def first_column():              # equation.py:113-117
        a = (m ** (3 - alpha) - (m - 2) ** (3 - alpha)) / g4
        b = - (3 * m ** (2 - alpha) + (m - 2) ** (2 - alpha)) / (2 * g3)
        c = (m ** (1 - alpha)) / g2
        return a + b + c

    def odd_node(d):                 # equation.py:119-122, d = i - k
        a = -2 * ((d + 1) ** (3 - alpha) - (d - 1) ** (3 - alpha)) / g4
        b = 2 * ((d + 1) ** (2 - alpha) + (d - 1) ** (2 - alpha)) / g3
        return a + b

    def even_node(d):                # equation.py:124-128
        a = ((d + 2) ** (3 - alpha) - (d - 2) ** (3 - alpha)) / g4
        b = ((d + 2) ** (2 - alpha) + 6 * d ** (2 - alpha) + (d - 2) ** (2 - alpha)) / (2 * g3)
        return a - b

    def diag_even_m():               # equation.py:130-133
        a = (2 ** (3 - alpha)) / g4
        b = -(2 ** (2 - alpha)) / (2 * g3)
        return a + b

    def subdiag_odd_m():             # equation.py:135-139
        a = (3 ** (3 - alpha) - 1) / g4f
        b = -(3 ** (2 - alpha) + 3) / (2 * g3f)
        c = -1. / g2f
        return a + b + c

    def w(k):                        # equation.py:141-154, branch order kept
        if k == -m:
            return first_column()
        elif m_is_even and k == 0:
            return diag_even_m()
        elif not m_is_even and k == -1:
            return subdiag_odd_m()
        elif math.fmod(-m + k, 2) == 0. and k < -1:
            return even_node(-k)
        else:
            return odd_node(-k)

    def u(k):                        # equation.py:156-164
        if k == -1:
            return (2. * (1. - alpha) ** 2 + 3. - 3. * alpha) / (2 * g4f)
        elif k == 0:
            return (4. - 2. * alpha) / g4f
        else:
            return (-1. + alpha) / (2. * g4f)

    mult = dt ** (1. - alpha)
    if m_is_even:
        return -m, [w(j - m) for j in range(m + 1)], mult
    raws = []
    for j in range(m + 2):           # equation.py:166-170
        k = j - m
        wk = w(k) if k <= -1 else 0.
        uk = u(k) if k in (-1, 0, 1) else 0.
        raws.append(wk + uk)
    return -m, raws, mult


_RULE_FN = {'caputo': _caputo, 'rectangle': _rectangle, 'trapezoidal': _trapezoidal, 'simpson': _simpson}


def left_arrays(approximation, settings):
    """-> (first_offset, [weight per consecutive integer offset]).  equation.py:28-29,76-81."""
    fn = _RULE_FN[approximation]                 # KeyError for an unknown key, as the reference
    alpha, lf, resolution = settings
    if resolution == 0:
        raise ZeroDivisionError('float division by zero')
    first, raws, mult = fn(alpha, lf, _as_int_resolution(resolution))
    return first, [mult * r for r in raws]       # equation.py:78-79: multiplier * weight(node)


def right_arrays(approximation, settings):
    """equation.py:32-39: offset o -> -o, weight -> weight * -1."""
    first, ws = left_arrays(approximation, settings)
    last = first + len(ws) - 1
    return -last, [w * -1. for w in reversed(ws)]


def riesz_caputo_arrays(approximation, settings):
    """equation.py:13-25 flattened at Point(0): K * (left + (-1.)**n * right)."""
    alpha, lf, resolution = settings
    first, ws = left_arrays(approximation, settings)
    n = math.floor(alpha) + 1.
    K = 1. / 2. * math.gamma(2. - alpha) / math.gamma(2.)
    sgn = (-1.) ** n
    left = {first + i: w for i, w in enumerate(ws)}
    lo = min(first, -(first + len(ws) - 1))
    out = []
    for o in range(lo, -lo + 1):
        terms = []
        if o in left:
            terms.append(left[o])
        if -o in left:
            terms.append(sgn * (left[-o] * -1.))
        if not terms:
            out.append(None)
        else:
            out.append(K * (terms[0] if len(terms) == 1 else terms[0] + terms[1]))
    # offsets present are contiguous for every rule (left part reaches -lo, mirror reaches +lo)
    assert all(v is not None for v in out)
    return lo, out


# --------------------------------------------------------------------------------------
# the same four public names as the reference, returning shim stencils
# --------------------------------------------------------------------------------------

def _node(settings, approximation, offset):
    alpha, lf, resolution = settings
    return offset * (lf / float(resolution))


def _to_stencil(settings, approximation, arrays):
    first, ws = arrays
    return fdm_shim.Stencil({fdm_shim.Point(_node(settings, approximation, first + i)): w
                             for i, w in enumerate(ws)})


def create_left_stencil(approximation, settings):
    return _to_stencil(settings, approximation, left_arrays(approximation, settings))


def create_right_stencil(approximation, settings):
    return _to_stencil(settings, approximation, right_arrays(approximation, settings))


def create_riesz_caputo_stencil(approximation, settings):
    return _to_stencil(settings, approximation, riesz_caputo_arrays(approximation, settings))


# --------------------------------------------------------------------------------------
# growing-history operator application (call pattern of reference
# fractulus/test/integration/test_equation.py:51-66): row i uses Settings(alpha, i*h, i).
# --------------------------------------------------------------------------------------

def history_row(approximation, alpha, h, g, i):
    """y_i = sum_j w_j * g[i + offset_j] with the stencil of Settings(alpha, i*h, i)."""
    first, ws = left_arrays(approximation, Settings(alpha, i * h, i))
    return sum([g[i + first + k] * w for k, w in enumerate(ws)])


def history(approximation, alpha, h, g, rows):
    return [history_row(approximation, alpha, h, g, i) for i in rows]
