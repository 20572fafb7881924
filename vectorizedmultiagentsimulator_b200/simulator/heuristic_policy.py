"""Scripted-policy base classes scenarios subclass for their ``HeuristicPolicy``
(API of ref vmas/simulator/heuristic_policy.py:10-22)."""
import abc

import torch
from torch import Tensor


class BaseHeuristicPolicy(abc.ABC):
    """Maps one agent's observation batch to an action batch."""

    def __init__(self, continuous_action: bool):
        self.continuous_actions = continuous_action

    @abc.abstractmethod
    def compute_action(self, observation: Tensor, u_range: float) -> Tensor:
        ...


class RandomPolicy(BaseHeuristicPolicy):
    """Gaussian 2-D actions clipped to the agent's range."""

    def compute_action(self, observation: Tensor, u_range: float) -> Tensor:
        sample = torch.randn(observation.shape[0], 2, device=observation.device)
        return sample.clamp_(-u_range, u_range)
