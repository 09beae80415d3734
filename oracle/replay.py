"""Replay buffers (uniform + prioritized) and the segment trees, restated in NumPy.  TEST INFRASTRUCTURE.

Follows
  * xuance/common/segtree_tool.py:4-220   (array-heap trees, end-exclusive reduce, find_prefixsum_idx)
  * xuance/common/memory_tools.py:331-387 (DummyOffPolicyBuffer store / independent env,step sampling)
  * xuance/common/memory_tools.py:471-598 (PerOffPolicyBuffer: one (sum,min) tree pair PER ENV, stratified
    sampling with p_total = sum(0, size-1), weights p/p_min, update_priorities with 0 -> 1e-8)

CANONICAL ARITHMETIC (DESIGN.md "PER numerics"): all leaves and all internal nodes are IEEE float32; a
parent is fl32(left + right) / min(left, right).  The reference reproduces exactly this when every leaf is
np.float32 (operator.add on two np.float32 gives np.float32).  Out of the box the reference mixes float64
(fresh leaves come from the float64 ``_max_priority`` array) with float32 (updated leaves, NumPy>=2) - an
accident of NumPy scalar promotion, SURVEY.md appendix B #10.  The live-reference test therefore swaps the
reference's ``_max_priority`` array for a float32 one (a test-side state change, no source change); with
that the reference and this oracle agree bit for bit on leaves, sums, sampled indices and weights.
"""
import random
import numpy as np

F32 = np.float32


class TreeF32:
    """Array-heap segment tree with float32 nodes.  segtree_tool.py:24-122."""

    def __init__(self, capacity, kind):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.cap = capacity
        self.kind = kind
        self.v = np.full(2 * capacity, 0.0 if kind == 'sum' else np.inf, dtype=F32)

    def _op(self, a, b):
        return F32(a + b) if self.kind == 'sum' else (a if a <= b else b)  # Python min(a,b): b only if b < a

    def set(self, idx, val):  # __setitem__ :87-104
        n = idx + self.cap
        self.v[n] = F32(val)
        n //= 2
        while n >= 1:
            self.v[n] = self._op(self.v[2 * n], self.v[2 * n + 1])
            n //= 2

    def get(self, idx):
        return self.v[self.cap + idx]

    def reduce(self, start=0, end=None):  # :41-85 ; END-EXCLUSIVE, recursion order preserved
        if end is None:
            end = self.cap
        if end < 0:
            end += self.cap
        end -= 1
        return self._red(start, end, 1, 0, self.cap - 1)

    def _red(self, s, e, node, ns, ne):
        if s == ns and e == ne:
            return self.v[node]
        mid = (ns + ne) // 2
        if e <= mid:
            return self._red(s, e, 2 * node, ns, mid)
        if mid + 1 <= s:
            return self._red(s, e, 2 * node + 1, mid + 1, ne)
        return self._op(self._red(s, mid, 2 * node, ns, mid), self._red(mid + 1, e, 2 * node + 1, mid + 1, ne))

    def find_prefixsum_idx(self, mass):  # :161-182 (float32 compare / subtract)
        mass = F32(mass)
        n = 1
        while n < self.cap:
            left = self.v[2 * n]
            if left > mass:
                n = 2 * n
            else:
                mass = F32(mass - left)
                n = 2 * n + 1
        return n - self.cap


def next_pow2(n):
    c = 1
    while c < n:
        c *= 2
    return c


class UniformReplayOracle:
    """DummyOffPolicyBuffer(+_Atari).  memory_tools.py:331-387, :601-630."""

    def __init__(self, obs_shape, act_shape, n_envs, buffer_size, batch_size, obs_dtype=np.float32):
        assert buffer_size % n_envs == 0
        self.n_envs, self.n_size, self.batch_size = n_envs, buffer_size // n_envs, batch_size
        z = lambda shape, dt=np.float32: np.zeros((n_envs, self.n_size) + tuple(shape), dt)
        self.observations, self.next_observations = z(obs_shape, obs_dtype), z(obs_shape, obs_dtype)
        self.actions, self.rewards, self.terminals = z(act_shape), z(()), z(())
        self.ptr = self.size = 0

    def store(self, obs, acts, rews, terminals, next_obs):
        p = self.ptr
        self.observations[:, p], self.actions[:, p] = obs, acts
        self.rewards[:, p], self.terminals[:, p] = rews, terminals
        self.next_observations[:, p] = next_obs
        self.ptr = (p + 1) % self.n_size
        self.size = min(self.size + 1, self.n_size)

    def gather(self, env, step):
        return {'obs': self.observations[env, step], 'actions': self.actions[env, step],
                'obs_next': self.next_observations[env, step], 'rewards': self.rewards[env, step],
                'terminals': self.terminals[env, step]}

    def sample(self, batch_size=None):  # :374-387: two independent np.random.choice draws, with replacement
        bs = self.batch_size if batch_size is None else batch_size
        env = np.random.choice(self.n_envs, bs)
        step = np.random.choice(self.size, bs)
        d = self.gather(env, step)
        d['batch_size'] = bs
        return d


class PerReplayOracle(UniformReplayOracle):
    """PerOffPolicyBuffer under the canonical float32 rule.  memory_tools.py:471-598."""

    def __init__(self, obs_shape, act_shape, n_envs, buffer_size, batch_size, alpha=0.6, obs_dtype=np.float32):
        super().__init__(obs_shape, act_shape, n_envs, buffer_size, batch_size, obs_dtype)
        self.alpha = alpha
        cap = next_pow2(self.n_size)
        self.sum = [TreeF32(cap, 'sum') for _ in range(n_envs)]
        self.min = [TreeF32(cap, 'min') for _ in range(n_envs)]
        self.max_priority = np.ones(n_envs, F32)

    def store(self, obs, acts, rews, terminals, next_obs):  # :537-550
        p = self.ptr
        for i in range(self.n_envs):
            leaf = self.max_priority[i] ** self.alpha  # float32 ** python float -> float32 (powf)
            self.sum[i].set(p, leaf)
            self.min[i].set(p, leaf)
        super().store(obs, acts, rews, terminals, next_obs)

    def sample_proportional(self, i, k, uniforms):  # :518-527
        p_total = self.sum[i].reduce(0, self.size - 1)  # excludes the newest valid slot (end-exclusive)
        seg = p_total / k                                 # float32 / python int -> float32
        out = []
        for j in range(k):
            mass = uniforms[j] * seg + j * seg            # python float * f32 -> f32 ; int * f32 -> f32 ; f32+f32
            out.append(int(self.sum[i].find_prefixsum_idx(mass)))
        return out

    def sample(self, beta, uniforms=None):  # :552-586
        """``uniforms``: [n_envs, B/n_envs] python floats in [0,1); default draws random.random() in the
        reference's order (env-major) so that random.seed(s) reproduces the reference's choices."""
        assert beta > 0
        k = int(self.batch_size / self.n_envs)
        env = np.arange(self.n_envs).repeat(k)
        steps = np.zeros((self.n_envs, k))
        weights = np.zeros((self.n_envs, k))
        for i in range(self.n_envs):
            u = [random.random() for _ in range(k)] if uniforms is None else [float(x) for x in uniforms[i]]
            idx = self.sample_proportional(i, k, u)
            total = self.sum[i].reduce()
            p_min = self.min[i].reduce() / total
            max_w = p_min * self.size ** (-beta)
            w = []
            for j in idx:
                p_s = self.sum[i].get(j) / total
                w.append((p_s * self.size ** (-beta)) / max_w)
            steps[i] = idx
            weights[i] = np.array(w)
        steps = steps.astype(np.int64)
        d = self.gather(env, steps.flatten())
        d.update(weights=weights, step_choices=steps, batch_size=self.batch_size)
        return d

    def update_priorities(self, idxes, priorities):  # :588-598
        k = int(self.batch_size / self.n_envs)
        priorities = np.asarray(priorities).reshape(self.n_envs, k)
        for i in range(self.n_envs):
            for j, p in zip(idxes[i], priorities[i]):
                p = F32(p)
                if p == 0:
                    p = F32(p + 1e-8)
                assert 0 <= j < self.size
                leaf = p ** self.alpha
                self.sum[i].set(int(j), leaf)
                self.min[i].set(int(j), leaf)
                self.max_priority[i] = max(self.max_priority[i], p)
