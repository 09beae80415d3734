"""Multi-process vector env for multi-agent environments - the QMIX caller side with the simulators in worker processes
(reference: xuance/environment/vector_envs/subprocess/subproc_vec_maenv.py:8-145): ``in_series`` environments per worker,
commands over Pipes, thunks shipped with cloudpickle, env ``i`` seeded ``env_seed + i``; same ``reset / step_async /
step_wait`` contract, auto-reset semantics and ``buf_state / buf_avail_actions`` latches as ``DummyVecMultiAgentEnv``
(``reset_obs / reset_avail_actions / reset_state`` in the info of a finished episode)."""
import multiprocessing as mp

import numpy as np

from ...common.spaces import space2shape
from .vector_env import VecEnv, AlreadySteppingError, NotSteppingError


def _ma_worker(remote, parent_remote, payload, first_seed):
    import pickle
    parent_remote.close()
    fns = pickle.loads(payload)
    envs = []
    for i, fn in enumerate(fns):
        try:
            envs.append(fn(env_seed=None if first_seed is None else first_seed + i))
        except TypeError:
            envs.append(fn())

    def step_env(env, action):
        obs, rew, term, trunc, info = env.step(action)
        if all(term.values()) or trunc:
            obs0, info0 = env.reset()
            info["reset_obs"], info["reset_avail_actions"], info["reset_state"] = obs0, info0['avail_actions'], info0['state']
        return obs, rew, term, trunc, info

    try:
        while True:
            cmd, data = remote.recv()
            if cmd == 'step':
                remote.send([step_env(env, a) for env, a in zip(envs, data)])
            elif cmd == 'reset':
                remote.send([env.reset() for env in envs])
            elif cmd == 'render':
                remote.send([env.render(data) for env in envs])
            elif cmd == 'describe':
                e = envs[0]
                remote.send(dict(env_info=e.env_info, groups_info=e.groups_info, observation_space=e.observation_space,
                                 action_space=e.action_space, agents=e.agents, num_agents=e.num_agents,
                                 state_space=e.state_space, max_episode_steps=e.max_episode_steps))
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


class SubprocVecMultiAgentEnv(VecEnv):
    def __init__(self, env_fns, env_seed=1, context='spawn', in_series=1):
        import cloudpickle
        self.waiting, self.closed = False, False
        n = len(env_fns)
        assert n % in_series == 0, "the number of envs must be divisible by the number of envs run in series"
        self.in_series, self.n_remotes = in_series, n // in_series
        ctx = mp.get_context(context)
        pipes = [ctx.Pipe() for _ in range(self.n_remotes)]
        self.remotes, work_remotes = [p[0] for p in pipes], [p[1] for p in pipes]
        self.ps = []
        for r in range(self.n_remotes):
            chunk = list(env_fns[r * in_series:(r + 1) * in_series])
            seed = None if env_seed is None else env_seed + r * in_series
            p = ctx.Process(target=_ma_worker, args=(work_remotes[r], self.remotes[r], cloudpickle.dumps(chunk), seed), daemon=True)
            p.start()
            self.ps.append(p)
        for w in work_remotes:
            w.close()
        self.remotes[0].send(('describe', None))
        d = self.remotes[0].recv()
        super().__init__(n, d["observation_space"], d["action_space"])
        self.env_info, self.groups_info = d["env_info"], d["groups_info"]
        self.agents, self.num_agents = d["agents"], d["num_agents"]
        self.state_space, self.max_episode_steps = d["state_space"], d["max_episode_steps"]
        self.buf_state = [np.zeros(space2shape(self.state_space)) for _ in range(n)]
        self.buf_obs = [{} for _ in range(n)]
        self.buf_avail_actions = [{} for _ in range(n)]
        self.buf_info = [{} for _ in range(n)]

    def _gather(self):
        out = []
        for remote in self.remotes:
            out.extend(remote.recv())
        return out

    def _latch(self, infos):
        self.buf_info = list(infos)
        self.buf_state = [i['state'] for i in infos]
        self.buf_avail_actions = [i['avail_actions'] for i in infos]

    def reset(self):
        assert not self.closed, "operating on a closed vector env"
        for remote in self.remotes:
            remote.send(('reset', None))
        obs, infos = zip(*self._gather())
        self.buf_obs = list(obs)
        self._latch(infos)
        return list(obs), list(infos)

    def step_async(self, actions):
        assert not self.closed, "operating on a closed vector env"
        if self.waiting:
            raise AlreadySteppingError
        if isinstance(actions, dict):
            assert self.num_envs == 1, "one action dict cannot be matched to %d environments" % self.num_envs
            actions = [actions]
        assert len(actions) == self.num_envs
        for r, remote in enumerate(self.remotes):
            remote.send(('step', list(actions[r * self.in_series:(r + 1) * self.in_series])))
        self.waiting = True

    def step_wait(self):
        if not self.waiting:
            raise NotSteppingError
        obs, rewards, terminated, truncated, infos = zip(*self._gather())
        self.waiting = False
        self.buf_obs = list(obs)
        self._latch(infos)
        return list(obs), list(rewards), list(terminated), list(truncated), list(infos)

    def render(self, mode):
        for remote in self.remotes:
            remote.send(('render', mode))
        return self._gather()

    def close_extras(self):
        if self.waiting:
            self._gather()
            self.waiting = False
        for remote in self.remotes:
            try:
                remote.send(('close', None))
            except (BrokenPipeError, OSError):
                pass
        for p in self.ps:
            p.join(timeout=5)
        self.closed = True
