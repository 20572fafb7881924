"""Shared machinery of the kinematic action models (diff drive, bicycle, drone).

Each model integrates a small ODE over ``dt`` (Euler or classic RK4) to get the pose change the
command asks for, then back-solves the force and torque that produce exactly that change under
the world's semi-implicit Euler step: ``a = (delta - v*dt) / dt^2``, ``F = m*a``, ``tau = I*a``.
Host-side torch ops that run just before ``World.step``; not on the CUDA hot path.
"""
from __future__ import annotations

from .common import Dynamics


class KinematicDynamics(Dynamics):
    def __init__(self, world, integration: str = "rk4"):
        super().__init__()
        assert integration in ("rk4", "euler"), "Integration method must be 'euler' or 'rk4'."
        self.world = world
        self.dt = world.dt
        self.integration = integration

    def f(self, state, *commands):
        raise NotImplementedError

    def euler(self, state, *commands):
        return self.dt * self.f(state, *commands)

    def runge_kutta(self, state, *commands):
        dt = self.dt
        k1 = self.f(state, *commands)
        k2 = self.f(state + dt * k1 / 2, *commands)
        k3 = self.f(state + dt * k2 / 2, *commands)
        k4 = self.f(state + dt * k3, *commands)
        return (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4)

    def _delta(self, state, *commands):
        step = self.euler if self.integration == "euler" else self.runge_kutta
        return step(state, *commands)

    def _drive(self, dx, dy, dyaw):
        """Writes the force / torque that realise the planar pose change (dx, dy, dyaw)."""
        agent, dt = self.agent, self.dt
        vel, ang_vel = agent.state.vel, agent.state.ang_vel
        ax = (dx - vel[:, 0] * dt) / dt**2
        ay = (dy - vel[:, 1] * dt) / dt**2
        a_yaw = (dyaw - ang_vel[:, 0] * dt) / dt**2
        fx = agent.mass * ax
        fy = agent.mass * ay
        torque = agent.moment_of_inertia * a_yaw
        agent.state.force[:, 0] = fx
        agent.state.force[:, 1] = fy
        agent.state.torque = torque.unsqueeze(-1)
