"""Host-side carrier objects for the stencils this package returns.

The reference hands its weights to the third-party package ``fdm`` (``fdm.equation.Stencil``,
``fdm.equation.Number``, ``fdm.geometry.Point``: reference fractulus/equation.py:4-5) and its users
compose them with ``fdm.Operator`` (reference fractulus/test/integration/test_equation.py:61-66,
88-94).  ``fdm`` is not part of the reference checkout; when it is importable this package returns
real ``fdm`` objects (see fractulus_b200/equation.py), otherwise these minimal equivalents, whose
behaviour is exactly what the reference's call sites need:

    Point(x, y=0, z=0)             hashable, tolerant equality, + - * by scalar
    Stencil({Point: weight})       ._weights, .expand(point), .symmetry(point), .multiply(f), ==,
                                   Stencil.uniform(left, right, n_intervals, provider), Stencil.central(span)
    Number(value | callable)       .expand(point) -> {point: value}
    Operator(stencil, element)     .expand(point): stencil applied to element
    a + b, Number * element        lazy algebra with .to_stencil(origin)

In addition a Stencil built by this package remembers its integer layout (``first_offset``, ``step``)
and keeps its weights as a NumPy array (``weights_array``) so batched code never touches the dict.
"""
import numpy as np

_ROUND = 9   # Points are identified after rounding: -0.6000000000000001 and -0.6 are the same node


class Point(object):
    __slots__ = ('x', 'y', 'z', '_key')

    def __init__(self, x=0., y=0., z=0.):
        self.x, self.y, self.z = x, y, z
        self._key = (round(x, _ROUND) + 0., round(y, _ROUND) + 0., round(z, _ROUND) + 0.)

    def __hash__(self):
        return hash(self._key)

    def __eq__(self, other):
        return isinstance(other, Point) and self._key == other._key

    def __ne__(self, other):
        return not self == other

    def __add__(self, other):
        return Point(self.x + other.x, self.y + other.y, self.z + other.z)

    def __sub__(self, other):
        return Point(self.x - other.x, self.y - other.y, self.z - other.z)

    def __mul__(self, scalar):
        return Point(self.x * scalar, self.y * scalar, self.z * scalar)

    __rmul__ = __mul__

    def __truediv__(self, scalar):
        return Point(self.x / scalar, self.y / scalar, self.z / scalar)

    def __neg__(self):
        return Point(-self.x, -self.y, -self.z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __repr__(self):
        return 'Point(%r, %r, %r)' % (self.x, self.y, self.z)


def _accumulate(into, node, weight):
    if node in into:
        into[node] = into[node] + weight
    else:
        into[node] = weight


class Element(object):
    def expand(self, point):
        raise NotImplementedError

    def __add__(self, other):
        return LazyOperation('+', self, other)

    def __mul__(self, other):
        return LazyOperation('*', self, other)

    def to_stencil(self, origin):
        return Stencil({node - origin: w for node, w in self.expand(origin).items()})


class Stencil(Element):
    def __init__(self, weights, first_offset=None, step=None, weights_array=None):
        self._dict = dict(weights) if weights is not None else None
        self._nodes = None
        self.first_offset = first_offset
        self.step = step
        self.weights_array = weights_array

    @property
    def _weights(self):
        """{Point: weight} -- what the reference's call sites read (test/unit/test_equation.py:18).  A stencil built from
        kernel output keeps its weights as an array and materialises the dictionary on first use."""
        if self._dict is None:
            xs = self._nodes() if callable(self._nodes) else self._nodes
            self._dict = {Point(float(x)): float(v) for x, v in zip(xs, self.weights_array)}
        return self._dict

    @classmethod
    def from_layout(cls, nodes_x, weights, first_offset, step):
        """nodes_x: node coordinates, or a callable producing them on demand; weights: float64 array (kept); integer
        layout remembered.  The {Point: weight} dictionary is built lazily."""
        st = cls(None, first_offset, step, np.asarray(weights, dtype=np.float64))
        st._nodes = nodes_x
        return st

    @classmethod
    def uniform(cls, left, right, resolution, weights_provider):
        span = right - left
        delta = span / resolution
        return cls({left + delta * i: weights_provider(i, left + delta * i) for i in range(int(resolution) + 1)})

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

    def __ne__(self, other):
        return not self == other

    __hash__ = None

    def __repr__(self):
        return 'Stencil(%r)' % (self._weights,)


class Number(Element):
    def __init__(self, value):
        self._value = value

    def value_at(self, point):
        return self._value(point) if callable(self._value) else self._value

    def expand(self, point):
        return {point: self.value_at(point)}


class Operator(Element):
    def __init__(self, stencil, element=None):
        self._stencil = stencil
        self._element = element

    def expand(self, point):
        fast = self._expand_from_arrays(point)
        if fast is not None:
            return fast
        outer = self._stencil.expand(point)
        if self._element is None:
            return outer
        out = {}
        for node, w in outer.items():
            for inner, wi in self._element.expand(node).items():
                _accumulate(out, inner, w * wi)
        return out

    def _expand_from_arrays(self, point):
        """Operator(kernel-built stencil, Operator(small stencil)) on the x axis -- the reference's call pattern
        (test/integration/test_equation.py:61-66) -- without materialising a Point per intermediate node: the same
        floating-point operations in the same order as the generic path below (node coordinates point + a + b, products
        wa * wb, first term assigned and later ones added in the outer stencil's node order, nodes identified after the
        same rounding), so the result is the same dictionary, in the same order.  None: not this pattern."""
        st, el = self._stencil, self._element
        if type(st) is not Stencil or st.weights_array is None or st._dict is not None or st._nodes is None:
            return None
        if type(el) is Operator and el._element is None:
            el = el._stencil
        if type(el) is not Stencil or point.y != 0. or point.z != 0.:
            return None
        inner = list(el._weights.items())
        if not 0 < len(inner) <= 16 or any(p.y != 0. or p.z != 0. for p, _ in inner):
            return None
        if callable(st._nodes):
            st._nodes = st._nodes()
        xs = np.asarray(st._nodes, dtype=np.float64)
        if xs.shape != st.weights_array.shape:
            return None
        X = point.x + xs                                                  # point + node
        Y = X[:, None] + np.array([p.x for p, _ in inner])[None, :]        # node + inner node
        Wt = st.weights_array[:, None] * np.array([float(w) for _, w in inner])[None, :]
        acc = {}
        for y, w in zip(Y.reshape(-1).tolist(), Wt.reshape(-1).tolist()):
            k = round(y, _ROUND) + 0.
            e = acc.get(k)
            if e is None:
                acc[k] = [y, w]
            else:
                e[1] = e[1] + w
        return {Point(y): w for y, w in acc.values()}


class LazyOperation(Element):
    def __init__(self, op, left, right):
        self._op, self._left, self._right = op, left, right

    def expand(self, point):
        a, b = self._left, self._right
        if self._op == '+':
            out = dict(a.expand(point))
            for node, w in b.expand(point).items():
                _accumulate(out, node, w)
            return out
        if isinstance(a, Number):
            s = a.value_at(point)
            return {node: s * w for node, w in b.expand(point).items()}
        if isinstance(b, Number):
            s = b.value_at(point)
            return {node: w * s for node, w in a.expand(point).items()}
        raise TypeError('only Number * element products are supported')


def install_as_fdm():
    """Register this carrier under the import name ``fdm`` (+ ``fdm.equation``, ``fdm.geometry``) for programs written
    against the reference's dependency (``from fdm import Stencil``, ``import fdm; fdm.Operator(...)``, reference
    fractulus/test/unit/test_equation.py:3-4, test/integration/test_equation.py:1,7) when the real ``fdm`` package is not
    installed.  The module carries ``__fdm_shim__`` so that fractulus_b200.equation keeps returning carrier objects."""
    import sys
    import types
    me = sys.modules[__name__]
    fdm = types.ModuleType('fdm')
    fdm.__fdm_shim__ = True
    eq, geo = types.ModuleType('fdm.equation'), types.ModuleType('fdm.geometry')
    for name in ('Stencil', 'Number', 'Operator', 'Element'):
        setattr(fdm, name, getattr(me, name))
        setattr(eq, name, getattr(me, name))
    geo.Point = fdm.Point = Point
    fdm.equation, fdm.geometry = eq, geo
    fdm.__path__ = []
    sys.modules['fdm'], sys.modules['fdm.equation'], sys.modules['fdm.geometry'] = fdm, eq, geo
    return fdm
