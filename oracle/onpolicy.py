"""On-policy rollout buffer + GAE, restated from the reference in NumPy.  TEST INFRASTRUCTURE (see oracle/__init__).

Follows xuance/common/memory_tools.py:
  * slot layout ``[n_envs, horizon, *shape]`` and the float32 default dtype ........ :12-40, :221-230
  * ``store`` (column write at ptr, ring advance) .................................... :232-240
  * ``finish_path`` (GAE recurrence, or discounted returns when use_gae=False) ...... :242-265
  * ``sample`` (divmod slot addressing, per-minibatch population-std adv norm) ...... :267-287
  * uint8 observation storage of the Atari variant ................................... :319-328
The arithmetic is kept in the reference's order and dtypes (float32 arrays, Python-float gamma/lambda under
NumPy>=2 weak-scalar promotion) so that results are bit-identical to the reference run in this container.
"""
import numpy as np


def gae_segment(rewards, values, dones, bootstrap, gamma, lam):
    """One path segment. memory_tools.py:246-256.  All inputs float32 1-D; ``bootstrap`` scalar.

    delta_t = r_t + (1-d_t)*gamma*V_{t+1} - V_t ;  A_t = delta_t + (1-d_t)*gamma*lam*A_{t+1}
    returns = A + V.  Evaluated sequentially from the back, in float32, in the reference's operand order."""
    rewards = np.asarray(rewards, np.float32)
    dones = np.asarray(dones, np.float32)
    vs = np.append(np.asarray(values, np.float32), [bootstrap], axis=0)  # float32 stays float32 (weak scalar)
    if vs.dtype != np.float32:  # a float64 bootstrap array element promotes, as in the reference
        pass
    adv = np.zeros_like(rewards)
    carry = 0
    for t in range(len(rewards) - 1, -1, -1):
        nd = 1 - dones[t]
        delta = rewards[t] + nd * gamma * vs[t + 1] - vs[t]
        carry = delta + nd * gamma * lam * carry
        adv[t] = carry
    ret = adv + vs[:-1]
    return adv, ret


def discounted_returns(x, gamma):
    """common_tools.py:160-174.  scipy.signal.lfilter([1], [1, -gamma], x[::-1])[::-1] runs its direct-form
    recurrence y[n] = x[n] + gamma*y[n-1] in float64 whatever the input dtype (b, a are float64); restated
    as that backward recurrence in float64."""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros(len(x), dtype=np.float64)
    acc = np.float64(0.0)
    g = np.float64(gamma)
    for t in range(len(x) - 1, -1, -1):
        acc = x[t] + g * acc
        out[t] = acc
    return out


class OnPolicyBufferOracle:
    """Restatement of DummyOnPolicyBuffer / DummyOnPolicyBuffer_Atari (memory_tools.py:182-328)."""

    def __init__(self, obs_shape, act_shape, aux_shapes, n_envs, horizon, use_gae=True, use_advnorm=True,
                 gamma=0.99, gae_lam=0.95, obs_dtype=np.float32):
        self.obs_shape, self.act_shape = tuple(obs_shape), tuple(act_shape)
        self.aux_shapes = dict(aux_shapes or {})
        self.n_envs, self.n_size = n_envs, horizon
        self.use_gae, self.use_advnorm = use_gae, use_advnorm
        self.gamma, self.gae_lam = gamma, gae_lam
        self.obs_dtype = obs_dtype
        self.start_ids = np.zeros(n_envs, np.int64)
        self.clear()

    @property
    def full(self):
        return self.size >= self.n_size

    def _zeros(self, shape, dtype=np.float32):
        return np.zeros((self.n_envs, self.n_size) + tuple(shape), dtype)

    def clear(self):  # :221-230 (note: start_ids is NOT reset by the reference's clear)
        self.ptr = self.size = 0
        self.observations = self._zeros(self.obs_shape, self.obs_dtype)
        self.actions = self._zeros(self.act_shape)
        self.rewards, self.returns, self.values = self._zeros(()), self._zeros(()), self._zeros(())
        self.terminals, self.advantages = self._zeros(()), self._zeros(())
        self.aux = {k: self._zeros(s) for k, s in self.aux_shapes.items()}

    def store(self, obs, acts, rews, value, terminals, aux_info=None):  # :232-240
        p = self.ptr
        self.observations[:, p] = obs
        self.actions[:, p] = acts
        self.rewards[:, p] = rews
        self.values[:, p] = value
        self.terminals[:, p] = terminals
        for k, v in (aux_info or {}).items():
            self.aux[k][:, p] = v
        self.ptr = (p + 1) % self.n_size
        self.size = min(self.size + 1, self.n_size)

    def finish_path(self, val, i):  # :242-265
        stop = self.n_size if self.full else self.ptr
        sl = np.arange(self.start_ids[i], stop).astype(np.int32)
        if self.use_gae:
            adv, ret = gae_segment(self.rewards[i, sl], self.values[i, sl], self.terminals[i, sl], val,
                                   self.gamma, self.gae_lam)
        else:
            vs = np.append(np.array(self.values[i, sl]), [val], axis=0)
            rw = np.append(np.array(self.rewards[i, sl]), [val], axis=0)
            ret = discounted_returns(rw, self.gamma)[:-1]
            adv = rw[:-1] + self.gamma * vs[1:] - vs[:-1]
        self.returns[i, sl] = ret
        self.advantages[i, sl] = adv
        self.start_ids[i] = self.ptr

    def sample(self, indexes):  # :267-287
        assert self.full
        env, step = divmod(np.asarray(indexes), self.n_size)
        adv = self.advantages[env, step]
        if self.use_advnorm:
            adv = (adv - np.mean(adv)) / (np.std(adv) + 1e-8)
        return {
            'obs': self.observations[env, step],
            'actions': self.actions[env, step],
            'returns': self.returns[env, step],
            'values': self.values[env, step],
            'aux_batch': {k: v[env, step] for k, v in self.aux.items()},
            'batch_size': len(indexes),
            'advantages': adv,
        }
