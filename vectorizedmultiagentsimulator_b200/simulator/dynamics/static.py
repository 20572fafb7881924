"""Import location scenarios use for ``Static`` (defined in :mod:`.basic`)."""
from .basic import Static  # noqa: F401
