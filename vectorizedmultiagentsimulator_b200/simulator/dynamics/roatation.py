"""Import location scenarios use for ``Rotation`` (defined in :mod:`.basic`)."""
from .basic import Rotation  # noqa: F401
