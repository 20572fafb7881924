"""Differential drive: u = (forward velocity, angular velocity) (ref dynamics/diff_drive.py:14-82)."""
import torch

from ._kinematic import KinematicDynamics


class DiffDrive(KinematicDynamics):
    def __init__(self, world, integration: str = "rk4"):
        super().__init__(world, integration)

    def f(self, state, u_command, ang_vel_command):
        heading = state[:, 2]
        return torch.stack(
            (u_command * torch.cos(heading), u_command * torch.sin(heading), ang_vel_command), dim=-1
        )

    @property
    def needed_action_size(self) -> int:
        return 2

    def process_action(self):
        u = self.agent.action.u
        pose = torch.cat((self.agent.state.pos, self.agent.state.rot), dim=1)
        delta = self._delta(pose, u[:, 0], u[:, 1])
        self._drive(delta[:, 0], delta[:, 1], delta[:, 2])
