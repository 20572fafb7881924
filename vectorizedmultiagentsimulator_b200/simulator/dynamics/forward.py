"""One action component pushing along the agent's heading (ref dynamics/forward.py)."""
import torch

from ..utils import TorchUtils
from .common import Dynamics


class Forward(Dynamics):
    @property
    def needed_action_size(self) -> int:
        return 1

    def process_action(self):
        agent = self.agent
        body_force = torch.zeros(agent.batch_dim, 2, device=agent.device, dtype=torch.float32)
        body_force[:, 0] = agent.action.u[:, 0]
        agent.state.force = TorchUtils.rotate_vector(body_force, agent.state.rot)
