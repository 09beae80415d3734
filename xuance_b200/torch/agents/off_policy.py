"""Off-policy agent core - mirror of xuance/torch/agents/core/off_policy.py (exploration schedule, replay memory,
``train_epochs`` = n_epochs x (memory.sample + learner.update), step-based ``train`` loop)."""
from copy import deepcopy

import numpy as np
import torch

from ...common import DummyOffPolicyBuffer, DummyOffPolicyBuffer_Atari
from ..rl_models import ActionOutput
from .agent import Agent


class OffPolicyAgent(Agent):
    def __init__(self, config, envs=None, observation_space=None, action_space=None, callback=None):
        super().__init__(config, envs, observation_space, action_space, callback)
        self.start_greedy = getattr(config, "start_greedy", None)
        self.end_greedy = getattr(config, "end_greedy", None)
        self.e_greedy = self.start_greedy
        self.delta_egreedy = None
        if self.start_greedy is not None:
            self.delta_egreedy = (self.start_greedy - self.end_greedy) / getattr(config, "decay_step_greedy", 1)
        self.start_noise = getattr(config, "start_noise", None)
        self.end_noise = getattr(config, "end_noise", None)
        self.noise_scale = self.start_noise
        if self.start_noise is not None:
            self.delta_noise = (self.start_noise - self.end_noise) / getattr(config, "decay_step_noise", 1)
        self.buffer_size = getattr(config, "buffer_size", None)
        self.batch_size = getattr(config, "batch_size", None)
        if self.world_size > 1 and self.buffer_size is not None:   # global sizes -> this rank's shard
            self.buffer_size //= self.world_size
            self.batch_size //= self.world_size

    def _build_memory(self, auxiliary_info_shape=None):
        self.atari = getattr(self.config, "env_name", None) == "Atari"
        Buffer = DummyOffPolicyBuffer_Atari if self.atari else DummyOffPolicyBuffer
        return Buffer(observation_space=self.observation_space, action_space=self.action_space,
                      auxiliary_shape=auxiliary_info_shape, n_envs=self.n_envs, buffer_size=self.buffer_size,
                      batch_size=self.batch_size, device=self.device)

    def _update_explore_factor(self):
        if self.e_greedy is not None:
            if self.e_greedy > self.end_greedy:
                self.e_greedy = self.start_greedy - self.current_step * self.delta_egreedy
        elif self.noise_scale is not None:
            if self.noise_scale >= self.end_noise:
                self.noise_scale = self.start_noise - self.current_step * self.delta_noise

    def exploration(self, pi_actions):
        if self.e_greedy is not None:
            mask = torch.rand(self.n_envs, device=pi_actions.device) < self.e_greedy
            rand = torch.randint(0, self.action_space.n, size=(self.n_envs,), device=pi_actions.device)
            return torch.where(mask, rand, pi_actions)
        if self.noise_scale is not None:
            noisy = pi_actions + torch.randn_like(pi_actions) * self.noise_scale
            lo = torch.as_tensor(self.action_space.low, device=pi_actions.device)
            hi = torch.as_tensor(self.action_space.high, device=pi_actions.device)
            return torch.clamp(noisy, lo, hi)
        return pi_actions

    @torch.no_grad()
    def get_actions(self, observations, test_mode=False):
        if isinstance(observations, np.ndarray):
            observations = torch.from_numpy(observations).to(self.device)
        out = self.model(observations)
        actions = out.actions
        if not test_mode:
            actions = self.exploration(actions)
        return ActionOutput(env_actions=actions.detach().cpu().numpy())

    def train_epochs(self, n_epochs=1):
        train_info = {}
        for e in range(n_epochs):
            samples = self.memory.sample()
            train_info = self.learner.update(sync=(e == n_epochs - 1), **samples)
        train_info["epsilon-greedy"] = self.e_greedy
        train_info["noise_scale"] = self.noise_scale
        return train_info

    def _after_update(self, train_steps):
        """Hook run after every train_epochs call inside ``train`` (PER anneals beta here, perdqn_agent.py:72)."""

    def train(self, train_steps):
        train_info = {}
        # a copy: the vector envs write the next observation into buf_obs in place, which would otherwise alias the
        # first stored transition's obs to its next_obs (the reference has exactly that aliasing on the first step)
        obs = np.array(self.train_envs.buf_obs, copy=True)
        for _ in range(train_steps):
            self.obs_rms.update(obs)
            obs = self._process_observation(obs)
            policy_out = self.get_actions(obs, test_mode=False)
            actions = policy_out.env_actions
            next_obs, rewards, terminals, truncations, infos = self.train_envs.step(actions)
            self.callback.on_train_step(self.current_step, envs=self.train_envs, model=self.model, obs=obs,
                                        policy_out=policy_out, next_obs=next_obs, rewards=rewards, terminals=terminals,
                                        truncations=truncations, infos=infos, train_steps=train_steps)
            self.memory.store(obs, actions, self._process_reward(rewards), terminals,
                              self._process_observation(next_obs))
            if self.current_step > self.start_training and self.current_step % self.training_frequency == 0:
                update_info = self.train_epochs(n_epochs=self.n_epochs)
                self.log_infos(update_info, self.current_step)
                train_info.update(update_info)
                self._after_update(train_steps)
                self.callback.on_train_epochs_end(self.current_step, model=self.model, memory=self.memory,
                                                  current_episode=self.current_episode, train_steps=train_steps,
                                                  update_info=update_info)
            self.returns = self.gamma * self.returns + rewards
            obs = deepcopy(next_obs)
            for i in range(self.n_envs):
                if terminals[i] or truncations[i]:
                    if getattr(self, "atari", False) and (not truncations[i]):
                        continue
                    obs[i] = infos[i]["reset_obs"]
                    self.train_envs.buf_obs[i] = obs[i]
                    self.ret_rms.update(self.returns[i:i + 1])
                    self.returns[i] = 0.0
                    self.current_episode[i] += 1
                    episode_info = {
                        f"Episode-Steps/rank_{self.rank}": {f"env-{i}": infos[i]["episode_step"]},
                        f"Train-Episode-Rewards/rank_{self.rank}": {f"env-{i}": infos[i]["episode_score"]}}
                    self.log_infos(episode_info, self.current_step)
                    train_info.update(episode_info)
            self.current_step += self.n_envs
            self._update_explore_factor()
            self.callback.on_train_step_end(self.current_step, envs=self.train_envs, model=self.model,
                                            train_steps=train_steps, train_info=train_info)
        return train_info

    def test(self, test_episodes=1, test_envs=None, close_envs=True):
        envs = test_envs or self.train_envs
        obs, _ = envs.reset()
        scores, running = [], np.zeros(envs.num_envs, np.float32)
        while len(scores) < test_episodes:
            acts = self.get_actions(self._process_observation(obs), test_mode=True).env_actions
            obs, rew, term, trunc, infos = envs.step(acts)
            running += rew
            for i in range(envs.num_envs):
                if term[i] or trunc[i]:
                    scores.append(float(infos[i].get("episode_score", running[i])))
                    running[i] = 0
                    if "reset_obs" in infos[i]:
                        obs[i] = infos[i]["reset_obs"]
        if close_envs and test_envs is not None:
            envs.close()
        return scores[:test_episodes]
