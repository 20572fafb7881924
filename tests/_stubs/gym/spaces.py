import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = dtype


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__(shape if shape is not None else np.asarray(low).shape, dtype)
        self.low, self.high = low, high


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = n


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)
        super().__init__(self.nvec.shape, np.int64)


class Tuple(Space):
    def __init__(self, spaces):
        super().__init__()
        self.spaces = tuple(spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)


class Dict(Space):
    def __init__(self, spaces):
        super().__init__()
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]
