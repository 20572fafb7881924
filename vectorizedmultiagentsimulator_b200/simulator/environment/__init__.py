"""Environment and the wrapper selector (ref vmas/simulator/environment/__init__.py:10-34)."""
from enum import Enum

from .environment import Environment


class Wrapper(Enum):
    RLLIB = 0
    GYM = 1
    GYMNASIUM = 2
    GYMNASIUM_VEC = 3

    def get_env(self, env: Environment, **kwargs):
        # The gym / gymnasium / rllib adapters are thin list<->tensor shims over Environment and
        # are outside the hot-path scope of this build; the reference's own
        # adapters work unchanged on this Environment because its surface is identical.
        raise NotImplementedError(
            f"Wrapper {self.name} is not bundled: wrap the returned Environment with the "
            "reference's adapter (its API is unchanged)"
        )


__all__ = ["Environment", "Wrapper"]
