"""Environment and the wrapper selector (ref vmas/simulator/environment/__init__.py:10-34)."""
from enum import Enum

from .environment import Environment


class Wrapper(Enum):
    RLLIB = 0
    GYM = 1
    GYMNASIUM = 2
    GYMNASIUM_VEC = 3

    def get_env(self, env: Environment, **kwargs):
        """``make_env(..., wrapper=...)`` (ref vmas/simulator/environment/__init__.py:16-34)."""
        if self is Wrapper.RLLIB:
            # ray is not a dependency of this package and the adapter is a 250-line VectorEnv shim
            # outside the physics path: fail up front, with the way out
            raise ImportError(
                "wrapper='rllib' is not bundled with vectorizedmultiagentsimulator_b200: wrap the Environment "
                "returned by make_env(..., wrapper=None) in the reference's "
                "vmas.simulator.environment.rllib.VectorEnvWrapper (the Environment API is the reference's)"
            )
        from . import gym_adapters

        cls = {
            Wrapper.GYM: gym_adapters.GymWrapper,
            Wrapper.GYMNASIUM: gym_adapters.GymnasiumWrapper,
            Wrapper.GYMNASIUM_VEC: gym_adapters.GymnasiumVectorizedWrapper,
        }[self]
        return cls(env, **kwargs)


__all__ = ["Environment", "Wrapper"]
