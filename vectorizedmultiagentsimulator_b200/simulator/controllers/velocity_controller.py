"""PID velocity controller: turns a velocity target in ``agent.action.u`` into a force
(API and arithmetic of ref vmas/simulator/controllers/velocity_controller.py:16-125).

``ctrl_params`` is ``[gain, integral time, derivative time]`` in ``"standard"`` form or
``[kP, kI, kD]`` in ``"parallel"`` form (``Ti = kP / kI``, ``Td = kD / kP``).  Runs as host-side
torch ops right before ``World.step``; it is not part of the CUDA hot path.
"""
from __future__ import annotations

import warnings
from typing import Optional

import torch


class VelocityController:
    def __init__(self, agent, world, ctrl_params=(1, 0, 0), pid_form: str = "standard"):
        self.agent = agent
        self.world = world
        self.dt = world.dt
        self.ctrl_gain = ctrl_params[0]
        if pid_form == "standard":
            self.integralTs, self.derivativeTs = ctrl_params[1], ctrl_params[2]
        elif pid_form == "parallel":
            self.integralTs = 0.0 if ctrl_params[1] == 0 else self.ctrl_gain / ctrl_params[1]
            self.derivativeTs = ctrl_params[2] / self.ctrl_gain
        else:
            raise Exception("PID form is either standard or parallel.")

        self.use_integrator = self.integralTs != 0
        if self.use_integrator:
            # anti-windup at half of the tighter force limit
            limits = [x for x in (agent.max_f, agent.f_range) if x is not None]
            if limits:
                fmax = min(limits)
                self.integrator_windup_cutoff = 0.5 * fmax * self.integralTs / (self.dt * self.ctrl_gain)
            else:
                self.integrator_windup_cutoff = None
                warnings.warn("Force limits not specified. Integrator can wind up!")
        self.reset()

    def reset(self, index: Optional[int] = None):
        if index is None:
            shape = (self.world.batch_dim, self.world.dim_p)
            self.accum_errs = torch.zeros(shape, device=self.world.device)
            self.prev_err = torch.zeros(shape, device=self.world.device)
        else:
            self.accum_errs = self.accum_errs.clone()
            self.prev_err = self.prev_err.clone()
            self.accum_errs[index] = 0.0
            self.prev_err[index] = 0.0

    def integralError(self, err):
        if not self.use_integrator:
            return 0
        self.accum_errs += self.dt * err
        if self.integrator_windup_cutoff is not None:
            self.accum_errs = self.accum_errs.clamp(
                -self.integrator_windup_cutoff, self.integrator_windup_cutoff
            )
        return (1.0 / self.integralTs) * self.accum_errs

    def rateError(self, err):
        rate = self.derivativeTs * (err - self.prev_err) / self.dt
        self.prev_err = err
        return rate

    def process_force(self):
        self.accum_errs = self.accum_errs.to(self.world.device)
        self.prev_err = self.prev_err.to(self.world.device)
        err = self.agent.action.u - self.agent.state.vel
        u = self.ctrl_gain * (err + self.integralError(err) + self.rateError(err))
        u *= self.agent.mass
        self.agent.action.u = u
