"""Import location scenarios use for ``HolonomicWithRotation`` (defined in :mod:`.basic`)."""
from .basic import HolonomicWithRotation  # noqa: F401
