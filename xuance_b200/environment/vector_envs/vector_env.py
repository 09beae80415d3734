"""Vectorised-environment contract (reference: xuance/environment/vector_envs/vector_env.py:5-78)."""
from abc import ABC, abstractmethod


class AlreadySteppingError(Exception):
    def __init__(self):
        Exception.__init__(self, 'already running an async step')


class NotSteppingError(Exception):
    def __init__(self):
        Exception.__init__(self, 'not running an async step')


class VecEnv(ABC):
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space, self.action_space = observation_space, action_space
        self.closed = False

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def render(self, mode):
        raise NotImplementedError

    def close_extras(self):
        pass

    def close(self):
        if not self.closed:
            self.close_extras()
        self.closed = True
