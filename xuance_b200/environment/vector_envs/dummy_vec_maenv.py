"""Sequential vector wrapper for multi-agent environments - contract of
xuance/environment/vector_envs/dummy/dummy_vec_maenv.py:6-83: observations / rewards / terminations are per-env dicts
keyed by agent, ``buf_state`` / ``buf_avail_actions`` mirror the latest ``info``, and an env whose agents have all
terminated (or that was truncated) is reset inside ``step_wait`` with the fresh episode's first observation, state and
availability mask handed back through ``info['reset_obs' | 'reset_state' | 'reset_avail_actions']``."""
import numpy as np

from ...common.spaces import space2shape
from .vector_env import VecEnv, AlreadySteppingError, NotSteppingError


class DummyVecMultiAgentEnv(VecEnv):
    def __init__(self, env_fns, env_seed=1):
        self.waiting = False
        self.envs = [self._make(fn, env_seed + i) for i, fn in enumerate(env_fns)]
        env = self.envs[0]
        super().__init__(len(self.envs), env.observation_space, env.action_space)
        self.env_info, self.groups_info = env.env_info, env.groups_info
        self.agents, self.num_agents = env.agents, env.num_agents
        self.state_space = env.state_space
        self.max_episode_steps = env.max_episode_steps
        self.buf_state = [np.zeros(space2shape(self.state_space)) for _ in range(self.num_envs)]
        self.buf_obs = [{} for _ in range(self.num_envs)]
        self.buf_avail_actions = [{} for _ in range(self.num_envs)]
        self.buf_info = [{} for _ in range(self.num_envs)]
        self.actions = None

    @staticmethod
    def _make(fn, seed):
        try:
            return fn(env_seed=seed)
        except TypeError:
            return fn()

    def _latch(self, e):
        self.buf_state[e] = self.buf_info[e]['state']
        self.buf_avail_actions[e] = self.buf_info[e]['avail_actions']

    def reset(self):
        for e, env in enumerate(self.envs):
            self.buf_obs[e], self.buf_info[e] = env.reset()
            self._latch(e)
        return self.buf_obs.copy(), self.buf_info.copy()

    def step_async(self, actions):
        if self.waiting:
            raise AlreadySteppingError
        if isinstance(actions, dict):                      # a single env's action dict
            assert self.num_envs == 1, "one action dict cannot be matched to %d environments" % self.num_envs
            actions = [actions]
        assert len(actions) == self.num_envs
        self.actions = actions
        self.waiting = True

    def step_wait(self):
        if not self.waiting:
            raise NotSteppingError
        rewards, terminated, truncated = [None] * self.num_envs, [None] * self.num_envs, [False] * self.num_envs
        for e, env in enumerate(self.envs):
            self.buf_obs[e], rewards[e], terminated[e], truncated[e], self.buf_info[e] = env.step(self.actions[e])
            self._latch(e)
            if all(terminated[e].values()) or truncated[e]:
                obs0, info0 = env.reset()
                self.buf_info[e]["reset_obs"] = obs0
                self.buf_info[e]["reset_avail_actions"] = info0['avail_actions']
                self.buf_info[e]["reset_state"] = info0['state']
        self.waiting = False
        return self.buf_obs.copy(), rewards, terminated, truncated, self.buf_info.copy()

    def close_extras(self):
        for env in self.envs:
            try:
                env.close()
            except Exception:
                pass

    def render(self, mode):
        return [env.render(mode) for env in self.envs]
