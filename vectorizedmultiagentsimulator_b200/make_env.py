"""``make_env``: scenario name (or instance) → ``Environment`` (ref vmas/make_env.py:14-101)."""
from __future__ import annotations

from typing import Optional, Union

from . import scenarios
from .simulator.environment import Environment, Wrapper
from .simulator.scenario import BaseScenario
from .simulator.utils import DEVICE_TYPING


def make_env(
    scenario: Union[str, BaseScenario],
    num_envs: int,
    device: DEVICE_TYPING = "cuda",
    continuous_actions: bool = True,
    wrapper: Optional[Union[Wrapper, str]] = None,
    max_steps: Optional[int] = None,
    seed: Optional[int] = None,
    dict_spaces: bool = False,
    multidiscrete_actions: bool = False,
    clamp_actions: bool = False,
    grad_enabled: bool = False,
    terminated_truncated: bool = False,
    wrapper_kwargs: Optional[dict] = None,
    cuda_graph: bool = False,
    action_checks: Optional[str] = None,
    **kwargs,
):
    """Create a vectorised environment.

    Arguments are the reference's.  ``scenario`` may be a scenario file name (looked up first
    among this package's scenarios, then in ``$VMAS_SCENARIO_PATH`` directories — e.g. an
    unmodified reference checkout's ``vmas/scenarios``), a path to a scenario file, or a
    ``BaseScenario`` instance.  ``device`` defaults to ``"cuda"``: the physics only runs there.
    ``cuda_graph=True`` replays one captured CUDA graph per ``step`` (graph-safe scenarios only),
    ``action_checks`` selects ``"sync"`` / ``"deferred"`` / ``"off"`` validation of the input
    actions (see ``Environment``).  Remaining ``kwargs`` go to ``Scenario.make_world``.
    """
    if isinstance(scenario, str):
        if not scenario.endswith(".py"):
            scenario += ".py"
        scenario = scenarios.load(scenario).Scenario()

    env = Environment(
        scenario,
        num_envs=num_envs,
        device=device,
        continuous_actions=continuous_actions,
        max_steps=max_steps,
        seed=seed,
        dict_spaces=dict_spaces,
        multidiscrete_actions=multidiscrete_actions,
        clamp_actions=clamp_actions,
        grad_enabled=grad_enabled,
        terminated_truncated=terminated_truncated,
        cuda_graph=cuda_graph,
        action_checks=action_checks,
        **kwargs,
    )
    if wrapper is not None and isinstance(wrapper, str):
        wrapper = Wrapper[wrapper.upper()]
    if wrapper_kwargs is None:
        wrapper_kwargs = {}
    return wrapper.get_env(env, **wrapper_kwargs) if wrapper is not None else env
