"""PPO-Clip agent - mirror of xuance/torch/agents/policy_gradient/ppo_agent.py:16-181."""
from copy import deepcopy

import numpy as np
import torch

from ...common.spaces import is_discrete
from ..rl_models import CategoricalActorHead, ValueHead, SharedActorCritic
from .on_policy import OnPolicyAgent


class PPO_Agent(OnPolicyAgent):
    def __init__(self, config, envs=None, observation_space=None, action_space=None, callback=None):
        super().__init__(config, envs, observation_space, action_space, callback)
        self.model = self._build_model()
        self.memory = self._build_memory(self.auxiliary_info_shape)
        self.learner = self._build_learner(self.config, self.model, self.callback)

    def _build_model(self):
        """ppo_agent.py:40-93 (shared representation + categorical actor head + value head)."""
        if not getattr(self.config, "shared_representation", True):
            raise NotImplementedError("separate actor/critic representations are outside the hot path scope")
        rep = self._build_representation(self.config.representation, self.observation_space, self.config)
        if not is_discrete(self.action_space):
            raise NotImplementedError("PPO with a Box action space (Gaussian head) is listed as 'next' in DESIGN.md; "
                                      "the fused K4 loss covers the categorical policy of BASELINE configs 1-2")
        kw = dict(normalizer=self.normalize_fn, initializer=self.initializer, activation=self.activation,
                  device=self.device)
        actor = CategoricalActorHead(feature_dim=rep.output_shapes['state'][0], hidden_size=self.config.actor_hidden_size,
                                     action_dim=self.action_space.n, **kw)
        critic = ValueHead(feature_dim=rep.output_shapes['state'][0], hidden_size=self.config.critic_hidden_size, **kw)
        return SharedActorCritic(representation=rep, actor=actor, critic=critic).to(self.device)

    @property
    def auxiliary_info_shape(self):
        return {"old_logp": ()}

    def get_aux_info(self, policy_output=None):
        return {"old_logp": policy_output.log_probs}

    def train(self, train_steps):
        """ppo_agent.py:111-181: one iteration = one vector-env step; update when the horizon is full."""
        train_info = {}
        obs = self.train_envs.buf_obs
        for _ in range(train_steps):
            self.obs_rms.update(obs)
            obs = self._process_observation(obs)
            policy_out = self.get_actions(obs, return_dists=False, return_logpi=True)
            acts, value = policy_out.env_actions, policy_out.values
            next_obs, rewards, terminals, truncations, infos = self.train_envs.step(acts)
            aux_info = self.get_aux_info(policy_out)
            self.callback.on_train_step(self.current_step, envs=self.train_envs, policy=self.model, obs=obs,
                                        policy_out=policy_out, acts=acts, vals=value, next_obs=next_obs,
                                        rewards=rewards, terminals=terminals, truncations=truncations, infos=infos,
                                        aux_info=aux_info, train_steps=train_steps)
            self.memory.store(obs, acts, self._process_reward(rewards), value, terminals, aux_info)
            if self.memory.full:
                vals = self.get_terminated_values(next_obs)
                for i in range(self.n_envs):
                    self.memory.finish_path(0.0 if terminals[i] else vals[i], i)
                update_info = self.train_epochs(self.n_epochs)
                self.log_infos(update_info, self.current_step)
                train_info.update(update_info)
                self.callback.on_train_epochs_end(self.current_step, policy=self.model, memory=self.memory,
                                                  current_episode=self.current_episode, train_steps=train_steps,
                                                  update_info=update_info)
                self.memory.clear()
            self.returns = self.gamma * self.returns + rewards
            obs = deepcopy(next_obs)
            for i in range(self.n_envs):
                if terminals[i] or truncations[i]:
                    self.ret_rms.update(self.returns[i:i + 1])
                    self.returns[i] = 0.0
                    if self.atari and (not truncations[i]):
                        continue
                    if terminals[i]:
                        self.memory.finish_path(0.0, i)
                    else:
                        vals = self.get_terminated_values(next_obs)
                        self.memory.finish_path(vals[i], i)
                    obs[i] = infos[i]["reset_obs"]
                    self.train_envs.buf_obs[i] = obs[i]
                    self.current_episode[i] += 1
                    episode_info = {
                        f"Episode-Steps/rank_{self.rank}": {f"env-{i}": infos[i]["episode_step"]},
                        f"Train-Episode-Rewards/rank_{self.rank}": {f"env-{i}": infos[i]["episode_score"]}}
                    self.log_infos(episode_info, self.current_step)
                    train_info.update(episode_info)
                    self.callback.on_train_episode_info(envs=self.train_envs, policy=self.model, env_id=i, infos=infos,
                                                        rank=self.rank, use_wandb=self.use_wandb,
                                                        current_step=self.current_step,
                                                        current_episode=self.current_episode, train_steps=train_steps)
            self.current_step += self.n_envs
            self.callback.on_train_step_end(self.current_step, envs=self.train_envs, policy=self.model,
                                            train_steps=train_steps, train_info=train_info)
        return train_info
