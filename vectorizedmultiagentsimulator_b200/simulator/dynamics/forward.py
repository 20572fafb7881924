"""Import location scenarios use for ``Forward`` (defined in :mod:`.basic`)."""
from .basic import Forward  # noqa: F401
