"""On-policy agent core - mirror of xuance/torch/agents/core/on_policy.py:16-397.

``train_epochs`` is the measured function of the headline metric (on_policy.py:182-205): per epoch a NumPy
shuffle of the slot indices (same RNG stream as the reference, so seeded runs pick identical minibatches), then
``n_minibatch`` x (``memory.sample`` + ``learner.update``).  Here the whole epoch's index permutation is uploaded
once, each minibatch is gathered and converted on the device (K3, fused u8->float in the network's layout), and
only the last minibatch's info is synchronised - the reference returns only that one (appendix B #4)."""
from copy import deepcopy

import numpy as np
import torch

from ... import _lib
from ...common import DummyOnPolicyBuffer, DummyOnPolicyBuffer_Atari
from ..rl_models import ActionOutput
from ..utils import stratified_minibatches
from .agent import Agent


class OnPolicyAgent(Agent):
    def __init__(self, config, envs=None, observation_space=None, action_space=None, callback=None):
        super().__init__(config, envs, observation_space, action_space, callback)
        self.continuous_control = hasattr(self.action_space, "low")
        self.horizon_size = config.horizon_size
        self.n_minibatch = config.n_minibatch
        self.gae_lam = config.gae_lambda
        self.memory = None
        self._shard_rng = np.random.default_rng(getattr(config, "seed", 0) * 1000 + self.rank)

    def _build_memory(self, auxiliary_info_shape=None):
        """on_policy.py:65-104 - the buffer holds THIS rank's envs (n_envs is the rank-local count)."""
        self.atari = getattr(self.config, "env_name", None) == "Atari"
        Buffer = DummyOnPolicyBuffer_Atari if self.atari else DummyOnPolicyBuffer
        self.buffer_size = self.n_envs * self.horizon_size
        self.batch_size = self.buffer_size // self.n_minibatch
        return Buffer(observation_space=self.observation_space, action_space=self.action_space,
                      auxiliary_shape=auxiliary_info_shape, n_envs=self.n_envs, horizon_size=self.horizon_size,
                      use_gae=self.config.use_gae, use_advnorm=self.config.use_advnorm, gamma=self.gamma,
                      gae_lam=self.gae_lam, device=self.device)

    def get_terminated_values(self, observations_next, rewards=None):
        return self.get_actions(self._process_observation(observations_next)).values

    @torch.no_grad()
    def get_actions(self, observations, deterministic=False, return_dists=False, return_logpi=False):
        """on_policy.py:128-169."""
        if isinstance(observations, np.ndarray):
            observations = torch.from_numpy(observations).to(self.device)
        out = self.model(observations)
        dists, values = out.distributions, out.values
        actions = dists.deterministic_sample() if deterministic else dists.stochastic_sample()
        log_pi = dists.log_prob(actions).cpu().numpy() if return_logpi else None
        values = 0 if values is None else values.cpu().numpy()
        return ActionOutput(env_actions=actions.cpu().numpy(), values=values,
                            distributions=dists if return_dists else None, log_probs=log_pi)

    def get_aux_info(self, policy_output=None):
        return {}

    def _obs_format(self):
        rep = getattr(self.model, "representation", None)
        if rep is not None and hasattr(rep, "preferred_obs_format"):
            return rep.preferred_obs_format()
        return _lib.OBS_U8

    def _sample_and_update(self, idx, adv_stats=None):
        """Device half of one minibatch: K3 gathers (+ adv-norm) -> network -> K4 -> backward -> [all-reduce] -> K7.
        ``adv_stats`` (sharded runs): the minibatch's global advantage mean / std, all-reduced once per epoch."""
        s = self.memory.sample_prepared(idx, self._obs_format(), adv_stats) if getattr(self.config, "fused_sample", True) \
            else self.memory.sample(idx, adv_stats)
        lrn = self.learner
        old_logp = s['aux_batch'].get('old_logp') if lrn.loss_kind == 0 else None
        lrn._device_update(s['obs'], s['actions'], s['returns'], s[lrn.adv_key], old_logp)

    def _graphed_minibatch(self, idx, adv_stats=None):
        """One CUDA-graph replay per minibatch (config.use_cuda_graph): the host only uploads 16 bytes of Adam
        hyper-parameters and the minibatch's slot indices.  Matters once the per-GPU minibatch is small (8 GPUs: 1024
        rows per rank) and Python launch overhead would otherwise bound the step."""
        from ..utils import CapturedStep
        key = int(idx.numel())
        args = [idx] if adv_stats is None else [idx, adv_stats]
        if key not in self._graphs:
            self._graphs[key] = CapturedStep(self._sample_and_update, args, self.learner.optimizer.snapshot,
                                             self.learner.optimizer.restore)
        self._graphs[key](*args)

    def train_epochs(self, n_epochs=1):
        """on_policy.py:182-205."""
        train_info = {}
        use_graph = getattr(self.config, "use_cuda_graph", False) and hasattr(self.learner, "_device_update")
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        # pending finish_path segments -> one K2 launch, OUTSIDE any captured region: a graph replay never runs the
        # buffer's host-side dirty check, so the scan must already have happened for every rollout
        self.memory._ensure_gae()
        total = n_epochs * -(-self.buffer_size // self.batch_size)     # a ragged last minibatch counts (ceil)
        done = 0
        indexes = np.arange(self.buffer_size)      # shuffled IN PLACE every epoch, as the reference does (on_policy.py:196-199):
        for _ in range(n_epochs):                  # epoch e's order is the composition of e shuffles, same RNG stream
            if self.world_size > 1:
                batches = stratified_minibatches(self.buffer_size, self.buffer_size // self.batch_size, self._shard_rng)
                perm = np.concatenate(batches)
            else:
                np.random.shuffle(indexes)
                perm = indexes
            perm_d = torch.from_numpy(perm).to(self.device, non_blocking=True)   # one H2D per epoch
            adv_all = None
            if self.world_size > 1 and getattr(self.memory, "use_advnorm", False):
                # ONE small all-reduce per epoch yields the global advantage statistics of all its minibatches
                adv_all = self.memory.global_adv_stats(perm_d, self.buffer_size // self.batch_size)
            for mb, start in enumerate(range(0, self.buffer_size, self.batch_size)):
                idx = perm_d[start:start + self.batch_size]
                stats = adv_all[mb] if adv_all is not None else None
                done += 1
                if use_graph:
                    self.learner.host_pre_step()
                    self._graphed_minibatch(idx, stats)
                    self.learner.host_post_step()
                    if done == total:
                        train_info = self.learner.materialize_info()
                else:
                    samples = self.memory.sample_prepared(idx, self._obs_format(), stats) \
                        if getattr(self.config, "fused_sample", True) else self.memory.sample(idx, stats)
                    train_info = self.learner.update(sync=(done == total), **samples)
        return train_info

    def test(self, test_episodes=1, test_envs=None, close_envs=True):
        envs = test_envs or self.train_envs
        obs, _ = envs.reset()
        scores, episode_score = [], np.zeros(envs.num_envs, np.float32)
        while len(scores) < test_episodes:
            obs = self._process_observation(obs)
            acts = self.get_actions(obs, deterministic=getattr(self.config, "deterministic_test", False)).env_actions
            obs, rew, term, trunc, infos = envs.step(acts)
            episode_score += rew
            for i in range(envs.num_envs):
                if term[i] or trunc[i]:
                    scores.append(float(infos[i].get("episode_score", episode_score[i])))
                    episode_score[i] = 0
                    if "reset_obs" in infos[i]:
                        obs[i] = infos[i]["reset_obs"]
        if close_envs and test_envs is not None:
            envs.close()
        return scores[:test_episodes]
