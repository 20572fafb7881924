"""Action/observation space descriptors.

The reference imports ``gym.spaces`` (ref environment.py:14).  ``gym`` / ``gymnasium`` are
optional here: when neither is installed these minimal stand-ins carry the same attributes
(`shape`, `dtype`, `low`, `high`, `n`, `nvec`, `spaces`) so ``Environment`` keeps its surface.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the host image
    from gym import spaces as _spaces  # type: ignore

    Box, Discrete, MultiDiscrete, Tuple, Dict = (
        _spaces.Box,
        _spaces.Discrete,
        _spaces.MultiDiscrete,
        _spaces.Tuple,
        _spaces.Dict,
    )
    BACKEND = "gym"
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gymnasium import spaces as _spaces  # type: ignore

        Box, Discrete, MultiDiscrete, Tuple, Dict = (
            _spaces.Box,
            _spaces.Discrete,
            _spaces.MultiDiscrete,
            _spaces.Tuple,
            _spaces.Dict,
        )
        BACKEND = "gymnasium"
    except Exception:  # noqa: BLE001
        BACKEND = "builtin"

        class Space:
            def __init__(self, shape=None, dtype=None):
                self.shape = None if shape is None else tuple(shape)
                self.dtype = None if dtype is None else np.dtype(dtype)

            def __repr__(self):
                return f"{type(self).__name__}(shape={self.shape}, dtype={self.dtype})"

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                if shape is None:
                    shape = np.asarray(low).shape
                super().__init__(shape, dtype)
                self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        class Discrete(Space):
            def __init__(self, n):
                super().__init__((), np.int64)
                self.n = int(n)

            def contains(self, x):
                return 0 <= int(x) < self.n

        class MultiDiscrete(Space):
            def __init__(self, nvec):
                self.nvec = np.asarray(nvec, dtype=np.int64)
                super().__init__(self.nvec.shape, np.int64)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        class Tuple(Space):
            def __init__(self, spaces):
                super().__init__(None, None)
                self.spaces = tuple(spaces)

            def __getitem__(self, i):
                return self.spaces[i]

            def __len__(self):
                return len(self.spaces)

            def __iter__(self):
                return iter(self.spaces)

        class Dict(Space):
            def __init__(self, spaces):
                super().__init__(None, None)
                self.spaces = dict(spaces)

            def __getitem__(self, k):
                return self.spaces[k]

            def keys(self):
                return self.spaces.keys()

            def items(self):
                return self.spaces.items()

            def __len__(self):
                return len(self.spaces)
