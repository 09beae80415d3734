"""Multi-process vector env (reference: xuance/environment/vector_envs/subprocess/subproc_vec_env.py:8-152): one
worker process per ``in_series`` envs, commands over Pipes, thunks shipped with cloudpickle, daemon workers that
swallow KeyboardInterrupt; same reset/step contract and auto-reset semantics as DummyVecEnv."""
import multiprocessing as mp

import numpy as np

from ...common.spaces import space2shape, combined_shape
from .vector_env import VecEnv, AlreadySteppingError, NotSteppingError


def _dumps(x):
    import cloudpickle
    return cloudpickle.dumps(x)


def _worker(remote, parent_remote, payload):
    import pickle
    parent_remote.close()
    envs = [fn() for fn in pickle.loads(payload)]

    def step_env(env, action):
        obs, rew, term, trunc, info = env.step(action)
        if term or trunc:
            obs_reset, _ = env.reset()
            info["reset_obs"] = obs_reset
        return obs, rew, term, trunc, info

    try:
        while True:
            cmd, data = remote.recv()
            if cmd == 'step':
                remote.send([step_env(env, a) for env, a in zip(envs, data)])
            elif cmd == 'reset':
                remote.send([env.reset(**({} if s is None else {"seed": s})) for env, s in zip(envs, data)])
            elif cmd == 'render':
                remote.send([env.render(data) for env in envs])
            elif cmd == 'get_spaces':
                remote.send((envs[0].observation_space, envs[0].action_space, envs[0].max_episode_steps))
            elif cmd == 'close':
                remote.close()
                break
            else:
                raise NotImplementedError(cmd)
    except KeyboardInterrupt:
        pass
    finally:
        for env in envs:
            try:
                env.close()
            except Exception:
                pass


class SubprocVecEnv(VecEnv):
    obs_dtype = np.float32

    def __init__(self, env_fns, env_seed=None, context='fork', in_series=1):
        self.waiting, self.closed = False, False
        n = len(env_fns)
        assert n % in_series == 0, "Number of envs must be divisible by number of envs to run in series"
        self.in_series, self.n_remotes = in_series, n // in_series
        groups = np.array_split(np.arange(n), self.n_remotes)
        ctx = mp.get_context(context)
        self.remotes, self.work_remotes = zip(*[ctx.Pipe() for _ in range(self.n_remotes)])
        self.ps = [ctx.Process(target=_worker, args=(wr, r, _dumps([env_fns[i] for i in g])), daemon=True)
                   for wr, r, g in zip(self.work_remotes, self.remotes, groups)]
        for p in self.ps:
            p.start()
        for wr in self.work_remotes:
            wr.close()
        self.remotes[0].send(('get_spaces', None))
        obs_space, act_space, self.max_episode_steps = self.remotes[0].recv()
        super().__init__(n, obs_space, act_space)
        self.obs_shape = space2shape(obs_space)
        self.buf_obs = np.zeros(combined_shape(n, self.obs_shape), dtype=self.obs_dtype)
        self.buf_info = [{} for _ in range(n)]
        self.env_seed = env_seed

    def reset(self):
        seeds = [None] * self.num_envs if self.env_seed is None else [self.env_seed + e for e in range(self.num_envs)]
        self.env_seed = None
        for r, s in zip(self.remotes, np.array_split(np.array(seeds, dtype=object), self.n_remotes)):
            r.send(('reset', list(s)))
        res = [x for r in self.remotes for x in r.recv()]
        for e, (obs, info) in enumerate(res):
            self.buf_obs[e], self.buf_info[e] = obs, info
        return self.buf_obs.copy(), list(self.buf_info)

    def step_async(self, actions):
        if self.waiting:
            raise AlreadySteppingError
        for r, a in zip(self.remotes, np.array_split(np.asarray(actions), self.n_remotes)):
            r.send(('step', a))
        self.waiting = True

    def step_wait(self):
        if not self.waiting:
            raise NotSteppingError
        res = [x for r in self.remotes for x in r.recv()]
        self.waiting = False
        obs, rews, terms, truncs, infos = zip(*res)
        self.buf_obs[...] = np.array(obs)
        self.buf_info = list(infos)
        return (self.buf_obs.copy(), np.array(rews, dtype=np.float32), np.array(terms, dtype=np.bool_),
                np.array(truncs, dtype=np.bool_), list(infos))

    def close_extras(self):
        if self.waiting:
            for r in self.remotes:
                r.recv()
        for r in self.remotes:
            try:
                r.send(('close', None))
            except Exception:
                pass
        for p in self.ps:
            p.join(timeout=2)

    def render(self, mode):
        for r in self.remotes:
            r.send(('render', mode))
        return [x for r in self.remotes for x in r.recv()]


class SubprocVecEnv_Atari(SubprocVecEnv):
    obs_dtype = np.uint8
