"""Minimal observation/action space types + ``space2shape`` (reference: xuance/environment/utils/shapes.py:5-18).

The reference takes gymnasium spaces; gymnasium is not a dependency of the hot path, so any object exposing
``.shape`` (and ``.n`` for discrete spaces, ``.dtype`` optionally) is accepted - gymnasium's own ``Box`` /
``Discrete`` instances work unchanged.  ``Box`` / ``Discrete`` here are light stand-ins for synthetic workloads."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        shape = np.shape(low) if shape is None else shape
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self):
        return np.random.randint(self.n)


def is_discrete(space):
    return hasattr(space, "n") and tuple(getattr(space, "shape", ()) or ()) == ()


def space2shape(space):
    """shapes.py:5-18: dict-of-spaces -> dict of shapes; tuple passes through; else ``space.shape``."""
    if isinstance(space, dict) or (hasattr(space, "keys") and hasattr(space, "__getitem__") and not hasattr(space, "shape")):
        return {k: space[k].shape for k in space.keys()}
    if isinstance(space, tuple):
        return space
    if hasattr(space, "spaces") and isinstance(getattr(space, "spaces"), dict):
        return {k: v.shape for k, v in space.spaces.items()}
    return space.shape


def combined_shape(length, shape=None):
    """shapes.py:21-45."""
    if shape is None:
        return (length,)
    return (length, shape) if np.isscalar(shape) else (length, *shape)
