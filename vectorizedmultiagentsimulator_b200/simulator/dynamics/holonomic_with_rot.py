"""Force from u[:, :2], torque from u[:, 2] (ref dynamics/holonomic_with_rot.py)."""
from .common import Dynamics


class HolonomicWithRotation(Dynamics):
    @property
    def needed_action_size(self) -> int:
        return 3

    def process_action(self):
        u = self.agent.action.u
        self.agent.state.force = u[:, :2]
        self.agent.state.torque = u[:, 2:3]
