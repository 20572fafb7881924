"""An agent that ignores its action (ref dynamics/static.py)."""
from .common import Dynamics


class Static(Dynamics):
    @property
    def needed_action_size(self) -> int:
        return 0

    def process_action(self):
        pass
