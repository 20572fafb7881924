"""Kinematic bicycle: u = (speed, steering angle) (ref dynamics/kinematic_bicycle.py:14-111;
model of Polack et al., IV 2017, eq. 2)."""
import torch

from ._kinematic import KinematicDynamics


class KinematicBicycle(KinematicDynamics):
    def __init__(self, world, width: float, l_f: float, l_r: float, max_steering_angle: float,
                 integration: str = "rk4"):
        super().__init__(world, integration)
        self.width = width
        self.l_f = l_f  # front axle to centre of gravity
        self.l_r = l_r  # rear axle to centre of gravity
        self.max_steering_angle = max_steering_angle

    def f(self, state, steering_command, v_command):
        yaw = state[:, 2]
        wheelbase = self.l_f + self.l_r
        slip = torch.atan2(
            torch.tan(steering_command) * self.l_r / wheelbase, torch.tensor(1, device=self.world.device)
        )
        dx = v_command * torch.cos(yaw + slip)
        dy = v_command * torch.sin(yaw + slip)
        dyaw = v_command / wheelbase * torch.cos(slip) * torch.tan(steering_command)
        return torch.stack((dx, dy, dyaw), dim=1)

    @property
    def needed_action_size(self) -> int:
        return 2

    def process_action(self):
        u = self.agent.action.u
        steering = torch.clamp(u[:, 1], -self.max_steering_angle, self.max_steering_angle)
        pose = torch.cat((self.agent.state.pos, self.agent.state.rot), dim=1)
        delta = self._delta(pose, steering, u[:, 0])
        self._drive(delta[:, 0], delta[:, 1], delta[:, 2])
