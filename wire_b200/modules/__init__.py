"""Drop-in counterpart of the reference's ``modules`` package for the WIRE models (wire, wire2d, models)."""
from . import wire, wire2d, models  # noqa: F401
