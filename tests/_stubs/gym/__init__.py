"""Test-only stand-in for ``gym`` so the UNMODIFIED reference (``/root/reference``) imports in
this container (it needs ``gym.Env`` and ``gym.spaces``; gym is not installed and there is no
network).  Never imported by the product package."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}


class Wrapper(Env):
    def __init__(self, env=None):
        self.env = env
