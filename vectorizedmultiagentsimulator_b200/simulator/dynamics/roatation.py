"""One action component used as torque (ref dynamics/roatation.py; file name kept as shipped)."""
from .common import Dynamics


class Rotation(Dynamics):
    @property
    def needed_action_size(self) -> int:
        return 1

    def process_action(self):
        self.agent.state.torque = self.agent.action.u[:, 0:1]
