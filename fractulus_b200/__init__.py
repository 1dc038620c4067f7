"""fractulus_b200 -- B200 (sm_100a) implementation of the fractional-operator hot path of
szajek/fractulus.  ``from fractulus_b200 import *`` gives the reference's four public names
(reference fractulus/__init__.py:1); ``fractulus_b200.batched`` is the tensor-level API."""
from .equation import *          # noqa: F401,F403  (Settings, create_*_stencil)
from . import equation           # noqa: F401
# batched conveniences of this package (not names of the reference; equation.__all__ stays the reference's)
from .equation import create_left_stencils, create_right_stencils, create_riesz_caputo_stencils   # noqa: F401
