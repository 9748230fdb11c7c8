"""Import stub (test infrastructure): lets the reference package import without matplotlib."""
from . import cm, colors, animation, tri, pyplot  # noqa: F401
