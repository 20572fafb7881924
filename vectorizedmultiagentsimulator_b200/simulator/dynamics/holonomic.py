"""Default model: the first two action components are the force (ref dynamics/holonomic.py:9-15)."""
from .common import Dynamics


class Holonomic(Dynamics):
    @property
    def needed_action_size(self) -> int:
        return 2

    def process_action(self):
        # the state setter copies into the world's force slab
        self.agent.state.force = self.agent.action.u[:, :2]
