"""TEST INFRASTRUCTURE -- minimal stand-in for the absent third-party package ``fdm``.

The reference imports ``fdm.equation.Stencil``, ``fdm.equation.Number`` and
``fdm.geometry.Point`` (reference fractulus/equation.py:4-5) and its tests also use
``fdm.Operator`` / ``fdm.Stencil.central`` / ``fdm.Number``
(reference fractulus/test/integration/test_equation.py:61-66,88-94).  ``fdm``
(github szajek/fdm, version unpinned: no requirements file in the reference) is not
vendored and not installable here, so this module restates ONLY the semantics the
reference's call sites pin.  It is an inferred stand-in, not upstream ``fdm``.
Acceptance test: the reference's own 13 tests pass on top of it
(oracle/gen_golden.py --selftest).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import sys
import types

__fdm_shim__ = True

_ROUND = 9  # tolerant Point identity: coordinates hashed after rounding


class Point(object):
    __slots__ = ('x', 'y', 'z', '_key')

    def __init__(self, x=0., y=0., z=0.):
        self.x, self.y, self.z = x, y, z
        self._key = (round(x, _ROUND) + 0., round(y, _ROUND) + 0., round(z, _ROUND) + 0.)

    def __hash__(self):
        return hash(self._key)

    def __eq__(self, other):
        return isinstance(other, Point) and self._key == other._key

    def __add__(self, o):
        return Point(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return Point(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, s):
        return Point(self.x * s, self.y * s, self.z * s)

    __rmul__ = __mul__

    def __neg__(self):
        return Point(-self.x, -self.y, -self.z)

    def __repr__(self):
        return 'Point(%r, %r, %r)' % (self.x, self.y, self.z)


def _merge(into, node, weight):
    into[node] = into[node] + weight if node in into else weight


class Element(object):
    def expand(self, point):
        raise NotImplementedError

    def __add__(self, other):
        return _Lazy('+', self, other)

    def __mul__(self, other):
        return _Lazy('*', self, other)

    def to_stencil(self, origin):
        return Stencil({node - origin: w for node, w in self.expand(origin).items()})


class Stencil(Element):
    def __init__(self, weights):
        self._weights = dict(weights)

    @classmethod
    def uniform(cls, left, right, resolution, weights_provider):
        # resolution = number of INTERVALS -> resolution + 1 nodes, node i at
        # left + i * (right - left) / resolution (pinned by the float reprs at
        # reference fractulus/test/unit/test_equation.py:21,24).
        span = right - left
        delta = Point(span.x / resolution, span.y / resolution, span.z / resolution)
        out = {}
        for i in range(int(resolution) + 1):
            node = left + delta * i
            out[node] = weights_provider(i, node)
        return cls(out)

    @classmethod
    def central(cls, span=1.):
        return cls({Point(-span / 2.): -1. / span, Point(span / 2.): 1. / span})

    def expand(self, point):
        return {point + node: w for node, w in self._weights.items()}

    def symmetry(self, point):
        return Stencil({point * 2. - node: w for node, w in self._weights.items()})

    def multiply(self, factor):
        return Stencil({node: w * factor for node, w in self._weights.items()})

    def __eq__(self, other):
        return isinstance(other, Stencil) and self._weights == other._weights

    __hash__ = None


class Number(Element):
    def __init__(self, value):
        self._value = value

    def value_at(self, point):
        return self._value(point) if callable(self._value) else self._value

    def expand(self, point):
        return {point: self.value_at(point)}


class Operator(Element):
    def __init__(self, stencil, element=None):
        self._stencil, self._element = stencil, element

    def expand(self, point):
        if self._element is None:
            return self._stencil.expand(point)
        out = {}
        for node, w in self._stencil.expand(point).items():
            for inner, wi in self._element.expand(node).items():
                _merge(out, inner, w * wi)
        return out


class _Lazy(Element):
    def __init__(self, op, a, b):
        self._op, self._a, self._b = op, a, b

    def expand(self, point):
        a, b = self._a, self._b
        if self._op == '+':
            out = dict(a.expand(point))
            for node, w in b.expand(point).items():
                _merge(out, node, w)
            return out
        if isinstance(a, Number):
            s = a.value_at(point)
            return {node: s * w for node, w in b.expand(point).items()}
        if isinstance(b, Number):
            s = b.value_at(point)
            return {node: w * s for node, w in a.expand(point).items()}
        raise TypeError('shim supports Number * element only')


def install_as_fdm():
    """Register this shim under the import name ``fdm`` (+ ``fdm.equation``, ``fdm.geometry``)."""
    me = sys.modules[__name__]
    fdm = types.ModuleType('fdm')
    fdm.__fdm_shim__ = True
    eq = types.ModuleType('fdm.equation')
    geo = types.ModuleType('fdm.geometry')
    for name in ('Stencil', 'Number', 'Operator', 'Element'):
        setattr(fdm, name, getattr(me, name))
        setattr(eq, name, getattr(me, name))
    geo.Point = Point
    fdm.Point = Point
    fdm.equation, fdm.geometry = eq, geo
    fdm.__path__ = []
    sys.modules['fdm'], sys.modules['fdm.equation'], sys.modules['fdm.geometry'] = fdm, eq, geo
    return fdm
