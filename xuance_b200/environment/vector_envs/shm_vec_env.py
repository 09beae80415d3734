"""Shared-memory multi-process vector env (the step on the other side of the boundary).

Same contract as ``SubprocVecEnv`` (reference: xuance/environment/vector_envs/subprocess/subproc_vec_env.py:8-152), but
the observations never travel through a pickle + pipe: every worker writes its envs' observations straight into one
``multiprocessing.shared_memory`` block laid out ``[num_envs, *obs_shape]`` (7.2 MB per step for 256 Atari frames),
only (reward, terminated, truncated, info) go over the pipe.  When CUDA is present the block is page-locked with
cudaHostRegister, so ``buf_obs`` is directly DMA-able: ``memory.store(envs.buf_obs_pinned, ...)`` is ONE async H2D copy
with no staging memcpy (the buffers detect pinned NumPy inputs)."""
import multiprocessing as mp
from multiprocessing import shared_memory

import numpy as np

from ...common.spaces import space2shape, combined_shape
from .vector_env import VecEnv, AlreadySteppingError, NotSteppingError
from .subproc_vec_env import _dumps


def _shm_worker(remote, parent_remote, payload, shm_name, shape, dtype, lo):
    import pickle
    parent_remote.close()
    envs = [fn() for fn in pickle.loads(payload)]
    shm = shared_memory.SharedMemory(name=shm_name)
    obs_all = np.ndarray(shape, dtype=dtype, buffer=shm.buf)
    try:
        while True:
            cmd, data = remote.recv()
            if cmd == 'step':
                out = []
                for j, (env, a) in enumerate(zip(envs, data)):
                    obs, rew, term, trunc, info = env.step(a)
                    if term or trunc:
                        info["reset_obs"], _ = env.reset()
                    obs_all[lo + j] = obs
                    out.append((rew, term, trunc, info))
                remote.send(out)
            elif cmd == 'reset':
                infos = []
                for j, (env, s) in enumerate(zip(envs, data)):
                    obs, info = env.reset(**({} if s is None else {"seed": s}))
                    obs_all[lo + j] = obs
                    infos.append(info)
                remote.send(infos)
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
        del obs_all
        shm.close()
        for env in envs:
            try:
                env.close()
            except Exception:
                pass


class ShmSubprocVecEnv(VecEnv):
    obs_dtype = np.float32

    def __init__(self, env_fns, env_seed=None, context='fork', in_series=1):
        self.waiting, self.closed = False, False
        n = len(env_fns)
        assert n % in_series == 0, "Number of envs must be divisible by number of envs to run in series"
        self.in_series, self.n_remotes = in_series, n // in_series
        probe = env_fns[0]()
        obs_space, act_space, self.max_episode_steps = probe.observation_space, probe.action_space, probe.max_episode_steps
        probe.close()
        super().__init__(n, obs_space, act_space)
        self.obs_shape = space2shape(obs_space)
        shape = combined_shape(n, self.obs_shape)
        nbytes = int(np.prod(shape)) * np.dtype(self.obs_dtype).itemsize
        self._shm = shared_memory.SharedMemory(create=True, size=max(nbytes, 1))
        self.buf_obs_pinned = np.ndarray(shape, dtype=self.obs_dtype, buffer=self._shm.buf)
        self.buf_obs_pinned[...] = 0
        self._registered = False
        try:   # page-lock the block so the device can DMA straight from it
            import torch
            if torch.cuda.is_available():
                ptr = self.buf_obs_pinned.ctypes.data
                self._registered = int(torch.cuda.cudart().cudaHostRegister(ptr, nbytes, 0)) == 0
        except Exception:
            self._registered = False
        ctx = mp.get_context(context)
        groups = np.array_split(np.arange(n), self.n_remotes)
        self.remotes, self.work_remotes = zip(*[ctx.Pipe() for _ in range(self.n_remotes)])
        self.ps = [ctx.Process(target=_shm_worker, daemon=True,
                               args=(wr, r, _dumps([env_fns[i] for i in g]), self._shm.name, shape, self.obs_dtype, int(g[0])))
                   for wr, r, g in zip(self.work_remotes, self.remotes, groups)]
        for p in self.ps:
            p.start()
        for wr in self.work_remotes:
            wr.close()
        self.buf_info = [{} for _ in range(n)]
        self.env_seed = env_seed

    @property
    def buf_obs(self):
        return self.buf_obs_pinned

    def reset(self):
        seeds = [None] * self.num_envs if self.env_seed is None else [self.env_seed + e for e in range(self.num_envs)]
        self.env_seed = None
        for r, s in zip(self.remotes, np.array_split(np.array(seeds, dtype=object), self.n_remotes)):
            r.send(('reset', list(s)))
        self.buf_info = [x for r in self.remotes for x in r.recv()]
        return self.buf_obs_pinned.copy(), list(self.buf_info)

    def step_async(self, actions):
        if self.waiting:
            raise AlreadySteppingError
        for r, a in zip(self.remotes, np.array_split(np.asarray(actions), self.n_remotes)):
            r.send(('step', a))
        self.waiting = True

    def step_wait(self, copy=True):
        """``copy=False`` returns the shared (pinned) observation block itself instead of a copy - valid until the next
        step; this is what the zero-staging H2D path uses."""
        if not self.waiting:
            raise NotSteppingError
        res = [x for r in self.remotes for x in r.recv()]
        self.waiting = False
        rews, terms, truncs, infos = zip(*res)
        self.buf_info = list(infos)
        obs = self.buf_obs_pinned.copy() if copy else self.buf_obs_pinned
        return (obs, np.array(rews, dtype=np.float32), np.array(terms, dtype=np.bool_),
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
        try:
            if self._registered:
                import torch
                torch.cuda.cudart().cudaHostUnregister(self.buf_obs_pinned.ctypes.data)
        except Exception:
            pass
        del self.buf_obs_pinned
        self._shm.close()
        self._shm.unlink()

    def render(self, mode):
        return []


class ShmSubprocVecEnv_Atari(ShmSubprocVecEnv):
    obs_dtype = np.uint8
