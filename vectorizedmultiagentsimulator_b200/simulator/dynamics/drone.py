"""Quadrotor: u = (thrust, torque x/y/z); a 12-state rigid body is integrated per env and its
planar motion + yaw drive the 2-D agent (ref dynamics/drone.py:17-166)."""
from typing import Union

import torch
from torch import Tensor

from ._kinematic import KinematicDynamics


class Drone(KinematicDynamics):
    def __init__(self, world, I_xx: float = 8.1e-3, I_yy: float = 8.1e-3, I_zz: float = 14.2e-3,
                 integration: str = "rk4"):
        super().__init__(world, integration)
        self.I_xx, self.I_yy, self.I_zz = I_xx, I_yy, I_zz
        self.g = 9.81
        self.reset()

    def reset(self, index: Union[Tensor, int] = None):
        # state: roll, pitch, yaw | body rates p, q, r | velocity x, y, z | position x, y, z
        if index is None:
            self.drone_state = torch.zeros(self.world.batch_dim, 12, device=self.world.device)
        else:
            self.drone_state = self.drone_state.clone()
            self.drone_state[index] = 0.0

    def zero_grad(self):
        self.drone_state = self.drone_state.detach()

    def f(self, state, thrust_command, torque_command):
        roll, pitch, yaw = state[:, 0], state[:, 1], state[:, 2]
        p, q, r = state[:, 3], state[:, 4], state[:, 5]
        c_r, s_r = torch.cos(roll), torch.sin(roll)
        c_p, s_p = torch.cos(pitch), torch.sin(pitch)
        c_y, s_y = torch.cos(yaw), torch.sin(yaw)
        mass = self.agent.mass
        x_ddot = (c_r * s_p * c_y + s_r * s_y) * thrust_command / mass
        y_ddot = (c_r * s_p * s_y - s_r * c_y) * thrust_command / mass
        z_ddot = (c_r * c_p) * thrust_command / mass - self.g
        p_dot = (torque_command[:, 0] - (self.I_yy - self.I_zz) * q * r) / self.I_xx
        q_dot = (torque_command[:, 1] - (self.I_zz - self.I_xx) * p * r) / self.I_yy
        r_dot = (torque_command[:, 2] - (self.I_xx - self.I_yy) * p * q) / self.I_zz
        return torch.stack(
            [p, q, r, p_dot, q_dot, r_dot, x_ddot, y_ddot, z_ddot, state[:, 6], state[:, 7], state[:, 8]], dim=-1
        )

    def needs_reset(self) -> Tensor:
        """Envs whose roll or pitch left +-30 degrees."""
        return torch.any(self.drone_state[:, :2].abs() > 30 * (torch.pi / 180), dim=-1)

    @property
    def needed_action_size(self) -> int:
        return 4

    def process_action(self):
        u = self.agent.action.u
        thrust = u[:, 0]
        thrust += self.agent.mass * self.g  # hover feed-forward (in place on the action, as shipped)
        torque = u[:, 1:4]
        self.drone_state[:, 9] = self.agent.state.pos[:, 0]
        self.drone_state[:, 10] = self.agent.state.pos[:, 1]
        self.drone_state[:, 2] = self.agent.state.rot[:, 0]
        delta = self._delta(self.drone_state, thrust, torque)
        self.drone_state = self.drone_state + delta
        self._drive(delta[:, 6], delta[:, 7], delta[:, 5])
