"""Episode replay for recurrent off-policy MARL (QMIX) - mirror of MARL_OffPolicyBuffer_RNN
(xuance/common/memory_tools_marl.py:770-996).

Same ``store(**step_data) / finish_path(i_env, **terminal_data) / sample()`` surface and sample-dict keys.  The
per-env episode scratch (``episode_data``) stays on the host - it is written a few hundred bytes at a time by the
vector env loop - while the replay ring (``data``) lives in HBM, agent-stacked ``[capacity, n_agents, T(+1), dim]``
so that one K1 row-store per field ingests a finished episode and one K3 row-gather per field builds a batch
already in the ``[B, n_agents, T(+1), dim]`` layout the learner consumes (the reference re-stacks per-agent
dicts on every update, marl_learner.py:336-394).  ``sample()`` returns per-agent views of those stacked tensors
under the reference's keys, plus the stacked tensors themselves under ``'_stacked'``."""
import numpy as np
import torch

from .. import _lib
from .spaces import space2shape


def _pad4(n):
    return (n + 3) // 4 * 4


class MARL_OffPolicyBuffer_RNN:
    def __init__(self, agent_keys, state_space=None, obs_space=None, act_space=None, n_envs=1, buffer_size=1,
                 batch_size=1, max_episode_steps=1, device="cuda:0", **kwargs):
        assert buffer_size % n_envs == 0, "buffer_size must be divisible by the number of envs (parallels)"
        self.agent_keys = list(agent_keys)
        self.n_agents = len(self.agent_keys)
        self.n_envs, self.buffer_size, self.batch_size = n_envs, buffer_size, batch_size
        self.max_eps_len = max_episode_steps
        self.state_space = state_space
        self.store_global_state = state_space is not None
        self.use_actions_mask = kwargs.get("use_actions_mask", False)
        self.avail_actions_shape = kwargs.get("avail_actions_shape", None)
        self.obs_shape = {k: tuple(space2shape(obs_space[k])) for k in self.agent_keys}
        self.act_shape = {k: tuple(space2shape(act_space[k])) for k in self.agent_keys}
        k0 = self.agent_keys[0]
        if any(self.obs_shape[k] != self.obs_shape[k0] or self.act_shape[k] != self.act_shape[k0] for k in self.agent_keys):
            raise NotImplementedError("heterogeneous agents are outside the hot-path scope (one parameter-sharing group)")
        if self.act_shape[k0] != ():
            raise NotImplementedError("QMIX uses discrete actions")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("xuance_b200 buffers are device-resident: device must be CUDA (no CPU fallback)")
        _lib.load()
        _lib.use_device(self.device)
        self._obs_dim = int(np.prod(self.obs_shape[k0]))
        self._state_dim = int(np.prod(space2shape(state_space))) if self.store_global_state else 0
        self._n_act = int(self.avail_actions_shape[k0][0]) if self.use_actions_mask else 0
        self.clear()
        self.clear_episodes()

    # ---- field table: name -> (per-episode shape, dtype, per-agent?)
    def _fields(self):
        n, T, Tp = self.n_agents, self.max_eps_len, _pad4(self.max_eps_len)
        f = {'obs': ((n, T + 1, self._obs_dim), torch.float32), 'actions': ((n, T), torch.float32),
             'rewards': ((n, T), torch.float32), 'terminals': ((n, Tp), torch.uint8),
             'agent_mask': ((n, Tp), torch.uint8), 'filled': ((Tp,), torch.uint8)}
        if self.store_global_state:
            f['state'] = ((T + 1, self._state_dim), torch.float32)
        if self.use_actions_mask:
            f['avail_actions'] = ((n, T + 1, _pad4(self._n_act)), torch.uint8)
        return f

    @property
    def data_keys(self):
        return list(self._fields().keys())

    def clear(self):
        self._dev = {k: torch.zeros((self.buffer_size,) + s, dtype=dt, device=self.device)
                     for k, (s, dt) in self._fields().items()}
        self.ptr, self.size = 0, 0

    def clear_episodes(self):
        self.episode_data = {k: torch.zeros((self.n_envs,) + s, dtype=dt).pin_memory()
                             for k, (s, dt) in self._fields().items()}
        self._ep_np = {k: v.numpy() for k, v in self.episode_data.items()}

    # ---- checkpoint (the reference does not save its replay)
    def state_dict(self):
        return {"class": type(self).__name__, "buffer_size": self.buffer_size, "n_envs": self.n_envs,
                "max_eps_len": self.max_eps_len, "ptr": int(self.ptr), "size": int(self.size),
                "data": {k: v.cpu() for k, v in self._dev.items()},
                "episode_data": {k: v.clone() for k, v in self.episode_data.items()}}

    def load_state_dict(self, sd):
        if (sd.get("class"), sd.get("buffer_size"), sd.get("n_envs"), sd.get("max_eps_len")) != \
                (type(self).__name__, self.buffer_size, self.n_envs, self.max_eps_len):
            raise ValueError("episode-replay checkpoint does not match this buffer's capacity / envs / episode length")
        for k, v in self._dev.items():
            if tuple(sd["data"][k].shape) != tuple(v.shape) or sd["data"][k].dtype != v.dtype:
                raise ValueError("episode-replay checkpoint field %s has shape %s" % (k, tuple(sd["data"][k].shape)))
            v.copy_(sd["data"][k])
        for k, v in self.episode_data.items():
            v.copy_(sd["episode_data"][k])
        self.ptr, self.size = int(sd["ptr"]), int(sd["size"])

    @property
    def full(self):
        return self.size >= self.buffer_size

    def can_sample(self):
        return self.size >= self.batch_size

    # ---- reference :912-929
    def store(self, **step_data):
        steps = step_data['episode_steps']
        envs = np.arange(self.n_envs)
        ep = self._ep_np
        ep['filled'][envs, steps] = 1
        for key, val in step_data.items():
            if key not in ep or key == 'filled':
                continue
            if key == 'state':
                ep['state'][envs, steps] = val
                continue
            for i, a in enumerate(self.agent_keys):
                if key == 'avail_actions':
                    ep[key][envs, i, steps, :self._n_act] = val[a]
                else:
                    ep[key][envs, i, steps] = val[a]

    # ---- reference :952-969 (+ store_episodes :931-950): terminal obs/state, then the episode goes to HBM
    def finish_path(self, i_env, **terminal_data):
        t = terminal_data['episode_step']
        ep = self._ep_np
        if self.store_global_state:
            ep['state'][i_env, t] = terminal_data['state']
        for i, a in enumerate(self.agent_keys):
            ep['obs'][i_env, i, t] = terminal_data['obs'][a]
            if self.use_actions_mask:
                ep['avail_actions'][i_env, i, t, :self._n_act] = terminal_data['avail_actions'][a]
        for key, dst in self._dev.items():
            row = self.episode_data[key][i_env:i_env + 1].to(self.device, non_blocking=True)
            flat = row.reshape(1, -1)
            row_bytes = flat.shape[1] * flat.element_size()
            # K1 rows path with N=1 "env", T=capacity slots: dst[0, ptr] = row
            _lib.call("xb_rollout_store", _lib.ptr(dst), _lib.ptr(flat), row_bytes, None, None, 0, 1,
                      self.buffer_size, int(self.ptr))
        torch.cuda.current_stream().synchronize()   # the pinned scratch row is reused right away
        self.ptr = (self.ptr + 1) % self.buffer_size
        self.size = min(self.size + 1, self.buffer_size)
        ep['filled'][i_env] = 0

    def _gather(self, key, idx_t):
        src = self._dev[key]
        row_shape = tuple(src.shape[1:])
        row_bytes = int(np.prod(row_shape)) * src.element_size()
        out = torch.empty((idx_t.numel(),) + row_shape, dtype=src.dtype, device=self.device)
        _lib.call("xb_gather_rows", _lib.ptr(src), _lib.ptr(idx_t), idx_t.numel(), row_bytes, _lib.ptr(out))
        return out

    # ---- reference :971-996
    def sample(self, batch_size=None, episode_choices=None):
        assert self.size > 0, "You need to first store experience data into the buffer!"
        if batch_size is None:
            batch_size = self.batch_size
        if episode_choices is None:
            episode_choices = np.random.choice(self.size, batch_size)
        idx_t = torch.from_numpy(np.asarray(episode_choices, dtype=np.int64)).to(self.device)
        T, A = self.max_eps_len, self._n_act
        st = {k: self._gather(k, idx_t) for k in self._dev}
        st['terminals'] = st['terminals'][:, :, :T].bool()
        st['agent_mask'] = st['agent_mask'][:, :, :T].bool()
        st['filled'] = st['filled'][:, :T].bool()
        if self.use_actions_mask:
            st['avail_actions'] = st['avail_actions'][..., :A].bool()
        out = {}
        for key, v in st.items():
            out[key] = v if key in ('filled', 'state') else {a: v[:, i] for i, a in enumerate(self.agent_keys)}
        out['batch_size'] = batch_size
        out['sequence_length'] = T
        out['_stacked'] = st
        return out
