import numpy as np
from . import Space


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.shape(low)
        super().__init__(tuple(shape), dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape)


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = start


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype)


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        self.n = n
        super().__init__((n,) if np.isscalar(n) else tuple(n), np.int8)


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kw):
        super().__init__(None, None)
        self.spaces = dict(spaces or {}, **kw)

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def __getitem__(self, k):
        return self.spaces[k]


class Tuple(Space):
    def __init__(self, spaces=(), seed=None):
        super().__init__(None, None)
        self.spaces = tuple(spaces)
