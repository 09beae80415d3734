"""In-process vector env (reference: xuance/environment/vector_envs/dummy/dummy_vec_env.py:17-103): sequential
step of every env, auto-reset with the first observation of the new episode stashed in ``info["reset_obs"]``,
copies of the batched buffers returned."""
import numpy as np

from ...common.spaces import space2shape, combined_shape
from .vector_env import VecEnv, AlreadySteppingError, NotSteppingError


class DummyVecEnv(VecEnv):
    obs_dtype = np.float32

    def __init__(self, env_fns, env_seed=None):
        self.waiting = False
        self.envs = [fn() for fn in env_fns]
        env = self.envs[0]
        super().__init__(len(env_fns), env.observation_space, env.action_space)
        self.obs_shape = space2shape(self.observation_space)
        self.buf_obs = np.zeros(combined_shape(self.num_envs, self.obs_shape), dtype=self.obs_dtype)
        self.buf_terminated = np.zeros((self.num_envs,), dtype=np.bool_)
        self.buf_truncated = np.zeros((self.num_envs,), dtype=np.bool_)
        self.buf_rewards = np.zeros((self.num_envs,), dtype=np.float32)
        self.buf_info = [{} for _ in range(self.num_envs)]
        self.actions = None
        self.max_episode_steps = env.max_episode_steps
        self.env_seed = env_seed

    def reset(self):
        for e in range(self.num_envs):
            kw = {} if self.env_seed is None else {"seed": self.env_seed + e}
            obs, info = self.envs[e].reset(**kw)
            self.buf_obs[e] = obs
            self.buf_info[e] = info
        self.env_seed = None
        self.buf_terminated[:] = False
        self.buf_truncated[:] = False
        self.buf_rewards[:] = 0
        return self.buf_obs.copy(), list(self.buf_info)

    def step_async(self, actions):
        if self.waiting:
            raise AlreadySteppingError
        try:
            ok = len(actions) == self.num_envs
        except TypeError:
            ok = False
        if not ok:
            assert self.num_envs == 1, "actions must provide one entry per environment"
            actions = [actions]
        self.actions = actions
        self.waiting = True

    def step_wait(self):
        if not self.waiting:
            raise NotSteppingError
        for e in range(self.num_envs):
            obs, self.buf_rewards[e], self.buf_terminated[e], self.buf_truncated[e], self.buf_info[e] = \
                self.envs[e].step(self.actions[e])
            if self.buf_terminated[e] or self.buf_truncated[e]:
                obs_reset, _ = self.envs[e].reset()
                self.buf_info[e]["reset_obs"] = obs_reset
            self.buf_obs[e] = obs
        self.waiting = False
        return (self.buf_obs.copy(), self.buf_rewards.copy(), self.buf_terminated.copy(), self.buf_truncated.copy(),
                list(self.buf_info))

    def close_extras(self):
        for env in self.envs:
            try:
                env.close()
            except Exception:
                pass

    def render(self, mode):
        return [env.render(mode) for env in self.envs]


class DummyVecEnv_Atari(DummyVecEnv):
    obs_dtype = np.uint8
