"""Import location scenarios use for ``Holonomic`` (defined in :mod:`.basic`)."""
from .basic import Holonomic  # noqa: F401
