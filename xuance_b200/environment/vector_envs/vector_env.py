"""The vectorised-environment contract the agents program against.

Surface kept from the reference (xuance/environment/vector_envs/vector_env.py:5-78): ``num_envs``, the two spaces,
``reset``, ``step`` = ``step_async`` then ``step_wait``, ``render``, ``close`` (idempotent; subclasses release their
processes / shared memory in ``close_extras``) and the two exceptions raised when the async protocol is misused.
"""
from abc import ABC, abstractmethod


class _StepProtocolError(Exception):
    """Raised when step_async / step_wait are called out of order."""
    message = "vector env step protocol violated"

    def __init__(self):
        super().__init__(self.message)


class AlreadySteppingError(_StepProtocolError):
    message = "already running an async step"


class NotSteppingError(_StepProtocolError):
    message = "not running an async step"


class VecEnv(ABC):
    closed = False

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = int(num_envs)
        self.observation_space = observation_space
        self.action_space = action_space

    # ---- what a backend implements
    @abstractmethod
    def reset(self):
        """-> (observations [num_envs, ...], list of info dicts)."""

    @abstractmethod
    def step_async(self, actions):
        """Hand one action per env to the backend; raises AlreadySteppingError if a step is still pending."""

    @abstractmethod
    def step_wait(self):
        """-> (obs, rewards, terminated, truncated, infos) of the pending step; NotSteppingError if there is none."""

    def close_extras(self):
        """Release backend resources (worker processes, shared memory).  Called once by ``close``."""

    # ---- provided
    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def render(self, mode):
        raise NotImplementedError("%s does not render" % type(self).__name__)

    def close(self):
        if self.closed:
            return
        self.close_extras()
        self.closed = True
