"""vectorizedmultiagentsimulator_b200 — a B200-native drop-in for VMAS's physics hot path.

``World.step`` (batched 2-D rigid-body substep: forces, Sphere/Box/Line contacts, joints,
semi-implicit Euler) and the LIDAR ray cast are hand-written sm_100a CUDA kernels behind the
reference's own Python API (``make_env`` / ``Environment.step`` / ``BaseScenario``).
"""
from .make_env import make_env
from .simulator.environment import Environment, Wrapper

__version__ = "0.1.0"

__all__ = ["make_env", "Environment", "Wrapper", "scenarios"]

#: scenarios shipped with this build (re-written on the public API; same names as the reference)
scenarios_list = ["balance", "flocking", "navigation", "transport"]
