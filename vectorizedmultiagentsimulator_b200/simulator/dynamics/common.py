"""Base class of the action→force models (ref vmas/simulator/dynamics/common.py:12-53)."""
from __future__ import annotations

import abc


class Dynamics(abc.ABC):
    def __init__(self):
        self._agent = None

    def reset(self, index=None):
        return

    def zero_grad(self):
        return

    @property
    def agent(self):
        if self._agent is None:
            raise ValueError(
                "You need to add the dynamics to an agent during construction before accessing its properties"
            )
        return self._agent

    @agent.setter
    def agent(self, value):
        if self._agent is not None:
            raise ValueError("Agent in dynamics has already been set")
        self._agent = value

    def check_and_process_action(self):
        u = self.agent.action.u
        if u.shape[1] < self.needed_action_size:
            raise ValueError(
                f"Agent action size {u.shape[1]} is less than the required dynamics action size {self.needed_action_size}"
            )
        self.process_action()

    @property
    @abc.abstractmethod
    def needed_action_size(self) -> int:
        raise NotImplementedError

    @abc.abstractmethod
    def process_action(self):
        raise NotImplementedError
