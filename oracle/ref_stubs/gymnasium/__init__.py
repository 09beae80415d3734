"""Import stub: just enough of `gymnasium` for the reference (agi-brain/xuance) to import in a
container that has no gymnasium.  TEST INFRASTRUCTURE ONLY - used by tests/golden/make_golden.py and
the `live reference` CPU tests; never imported by the product package."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)

    @property
    def shape(self):
        return self._shape

    def __class_getitem__(cls, item):
        return cls


class Env:
    metadata = {}
    observation_space = None
    action_space = None

    def __class_getitem__(cls, item):
        return cls


class Wrapper(Env):
    def __init__(self, env=None):
        self.env = env


class ObservationWrapper(Wrapper):
    pass


class ActionWrapper(Wrapper):
    pass


class RewardWrapper(Wrapper):
    pass


def make(*a, **k):
    raise RuntimeError("gymnasium stub: no simulators in this container")


def register_envs(*a, **k):
    return None


from . import spaces  # noqa: E402,F401
